"""Mixed-model batches (cfg3): chunking, static LPT plan, and the dynamic store-backed work queue on 2 gloo ranks."""
import os
import subprocess
import sys

import numpy as np

from mj_grasp_sim_b200.mixed import Bucket, make_chunks, plan_static, run_mixed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_bucket(name, n, weight=1.0):
    truth = (np.arange(n) * 7 + len(name)) % 3 == 0
    return Bucket(name, n, lambda lo, hi: (truth[lo:hi], 10 * (hi - lo), 1e-3 * (hi - lo)), weight), truth


def test_chunks_cover_every_candidate_once():
    bks = [fake_bucket("a", 10)[0], fake_bucket("b", 0)[0], fake_bucket("c", 4097)[0], fake_bucket("d", 2048)[0]]
    ch = make_chunks(bks, 2048)
    for b, bk in enumerate(bks):
        seg = sorted((lo, hi) for bb, lo, hi in ch if bb == b)
        assert sum(hi - lo for lo, hi in seg) == bk.n
        assert all(seg[i][1] == seg[i + 1][0] for i in range(len(seg) - 1))
    assert [(lo, hi) for b, lo, hi in ch if b == 2] == [(0, 2048), (2048, 4097)]  # ragged tail merged


def test_static_plan_balances_two_models():
    # cfg3 shape: 8 objects x {panda, vx300} x 4096, vx300 ~1.3x the cost of panda
    bks = [fake_bucket(f"{g}{k}", 4096, w)[0] for k in range(8) for g, w in (("panda", 1.0), ("vx300", 1.3))]
    for world in (1, 2, 4, 8):
        plan, chunks = plan_static(bks, 4096, world)
        assert sorted(i for p in plan for i in p) == list(range(len(chunks)))
        load = [sum(bks[chunks[i][0]].cost_per_candidate * (chunks[i][2] - chunks[i][1]) for i in p) for p in plan]
        assert max(load) / (sum(load) / world) < 1.05, (world, load)


def test_single_process_run_matches_truth():
    pairs = [fake_bucket("x", 100), fake_bucket("yy", 37)]
    labels, st = run_mixed([p[0] for p in pairs], 32)
    for (bk, truth), got in zip(pairs, labels):
        assert (got == truth).all()
    assert st["candidates"] == 137 and st["env_steps"] == 1370


WORKER = r"""
import os, sys, time
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from mj_grasp_sim_b200.mixed import Bucket, run_mixed
dist.init_process_group("gloo")
rank = dist.get_rank()
def mk(name, n):
    truth = (np.arange(n) * 7 + len(name)) % 3 == 0
    def run(lo, hi):
        time.sleep(0.02 if rank == 0 else 0.002)   # rank 0 is slow: the dynamic queue must give it fewer chunks
        return truth[lo:hi], 10 * (hi - lo), 0.001
    return Bucket(name, n, run), truth
pairs = [mk("a", 640), mk("bb", 333), mk("ccc", 64)]
for dynamic in (True, False, True):
    labels, st = run_mixed([p[0] for p in pairs], 32, dynamic=dynamic)
    for (bk, truth), got in zip(pairs, labels):
        assert (got == truth).all(), (dynamic, bk.name)
    tot = [None, None]
    dist.all_gather_object(tot, st["chunks"])
    assert sum(tot) == 20 + 10 + 2, tot
    if dynamic:
        assert tot[1] > tot[0], tot
if rank == 0:
    print("MIXED_OK", tot)
dist.destroy_process_group()
"""


def test_two_rank_dynamic_queue(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29537", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "MIXED_OK" in out.stdout, out.stdout + out.stderr

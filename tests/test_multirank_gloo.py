"""World-size-2 gloo run of the sharded evaluation path (host logic only; compute = lane-1 harness)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from mj_grasp_sim_b200 import scenes, shard
from mj_grasp_sim_b200.lib import MgsRolloutCfg
from hostsim import lane1
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m, info, pose7, joints = scenes.workload("panda", "cube", 0, 7)   # ragged: 7 candidates over 2 ranks
lo, hi = shard.shard_range(len(pose7), rank, world)
L = lane1.sim(m)
sched = MgsRolloutCfg(150, 60, 10, 0, 0.01, 0.01)
lab, _ = L.stability(pose7[lo:hi], joints[lo:hi], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], sched)
full = shard.gather_labels(lab, len(pose7))
ref, _ = L.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], sched)
assert full.shape == ref.shape and (full == ref).all(), (full, ref)
# the env-level path: rounds of world x chunk with the enough_stable early stop (shard.evaluate_sharded)
evaluated = []
def run_range(lo, hi):
    evaluated.append(hi - lo)
    return L.stability(pose7[lo:hi], joints[lo:hi], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], sched)[0]
for enough in (None, 1, 2, 100):
    del evaluated[:]
    got = shard.evaluate_sharded(len(pose7), run_range, enough, chunk=2)
    assert (got == shard.apply_enough_stable(ref, enough)).all(), (enough, got, ref)
    if enough == 1 and ref[:4].any():
        assert sum(evaluated) <= 2  # first round (2 ranks x 2 candidates) already reached the target: no second round
if rank == 0:
    print("GLOO_OK", full.astype(int))
dist.destroy_process_group()
"""


def test_two_rank_shard_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "GLOO_OK" in out.stdout, out.stdout + out.stderr

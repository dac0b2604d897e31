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
if rank == 0:
    ref, _ = L.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], sched)
    assert full.shape == ref.shape and (full == ref).all(), (full, ref)
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

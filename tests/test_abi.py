"""The C-ABI library: builds for sm_100a, exports every symbol include/mgs_b200.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mj_grasp_sim_b200 import lib as mlib
from mj_grasp_sim_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    return mlib.build()


def test_exports_match_header(so):
    hdr = open(os.path.join(ROOT, "include", "mgs_b200.h")).read()
    names = set(re.findall(r"\b(mgs_[a-z_]+)\s*\(", hdr))
    assert {"mgs_model_create", "mgs_grasp_stability", "mgs_grasp_collision_mask", "mgs_rollout_device", "mgs_step_device",
            "mgs_step_host", "mgs_last_error", "mgs_model_destroy", "mgs_model_info", "mgs_launch_count", "mgs_last_aux",
            "mgs_set_qvel_clip", "mgs_build_stamp", "mgs_overflow_count"} <= names
    L = C.CDLL(so)
    for n in names:
        assert hasattr(L, n), n


def test_build_stamp_matches_the_checked_out_sources(so):
    """The prebuilt .so files are git-ignored artefacts: the stamp compiled into the library must be the hash of the sources
    in this tree, so a stale library cannot pass for HEAD."""
    L = C.CDLL(so)
    L.mgs_build_stamp.restype = C.c_char_p
    assert L.mgs_build_stamp().decode() == mlib.source_stamp()
    so64 = mlib.build(f64=True)
    L64 = C.CDLL(so64)
    L64.mgs_build_stamp.restype = C.c_char_p
    assert L64.mgs_build_stamp().decode() == mlib.source_stamp()


def test_model_desc_header_in_sync():
    from mj_grasp_sim_b200.model_desc import c_declaration
    assert open(os.path.join(ROOT, "include", "mgs_model_desc.h")).read() == c_declaration()


def test_no_cpu_fallback(panda_cube):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mlib.MgsError, match="no CUDA device|CUDA"):
        mlib.BatchSim(panda_cube[0])


def test_sass_is_sm100(so):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_shard_and_enough_stable():
    assert [shard.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard.shard_range(2, 3, 4) == (2, 2)
    lab = np.array([1, 0, 1, 1, 0, 1], dtype=bool)
    assert shard.apply_enough_stable(lab, 2).tolist() == [True, False, True, False, False, False]
    assert shard.apply_enough_stable(lab, None).tolist() == lab.tolist()

"""GPU test (-m gpu) of the library's first-pass capacity selection (csrc/mgs_b200.cu mgs_model_create_ex).  In its own file, last in
collection order: it asserts capacities measured in round 2 (profiles/bench_r2d_allegro.json, bench_r2e_leap.json), nothing else
depends on it."""
import os

import numpy as np  # noqa: F401
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    from mj_grasp_sim_b200 import lib as mlib
    from oracle import oracle as orc
    mlib.load()
    return mlib, orc


def test_first_pass_capacity_selection(libs, allegro_hull, leap_hull, panda_hull):
    """With the capacities left to the library (ncon_max = nefc_max = 0) a single-object scene whose default capacity would fall to the
    environment-per-CTA variant gets the largest smaller one that fits four warp-environments per SM: the fp64 Allegro hand 16
    contacts / 86 rows, the fp64 LEAP hand 12 / 92 (profiles/bench_r2d_allegro.json, bench_r2e_leap.json); explicit capacities and
    models that fit anyway are left alone."""
    mlib, _ = libs
    if not os.path.exists(mlib.SO_PATH_F64):
        pytest.fail("the fp64 build is missing")
    for fixture, caps in ((allegro_hull, (16, 86)), (leap_hull, (12, 92))):
        G = mlib.BatchSim(fixture[0], f64=True)
        assert (G.info.ncon_max, G.info.nefc_max) == caps and G.info.lanes_per_env == 32 and G.info.warps_per_block * G.info.blocks_per_sm == 4
        G.close()
        E = mlib.BatchSim(fixture[0], f64=True, ncon_max=32)  # explicit capacity: taken as given (here: the environment-per-CTA variant)
        assert E.info.ncon_max == 32 and E.info.lanes_per_env == 256
        E.close()
    P = mlib.BatchSim(panda_hull[0])
    assert P.info.lanes_per_env == 32 and P.info.ncon_max >= 24 and P.info.warps_per_block * P.info.blocks_per_sm >= 8
    P.close()

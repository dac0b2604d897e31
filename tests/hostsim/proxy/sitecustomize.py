"""CPU PROXY for the GPU parity module (test harness only, never part of the product).

  PYTHONPATH=tests/hostsim/proxy python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
      --deselect tests/test_gpu_parity.py::test_device_pointer_entry_with_torch

routes `BatchSim` to the 1-lane HOST builds of the kernel source (tests/hostsim/lane1.cpp, fp32 and fp64), so that every
oracle-facing assertion of tests/test_gpu_parity.py can be exercised without a GPU when GPU minutes are short (same source,
same arithmetic up to reduction order; capacities are the host build's defaults, overflow counts read 0).  It says nothing
about the CUDA build itself - the real `-m gpu` run on the B200 box remains the gate.
"""
import ctypes as C
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for _p in (_ROOT, os.path.join(_ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

if os.environ.get("MGS_PROXY", "1") == "1":
    from hostsim import lane1
    from mj_grasp_sim_b200 import lib as mlib
    import mj_grasp_sim_b200.mgs.env.clutter_table as _ct
    import mj_grasp_sim_b200.mgs.env.gravityless_object_grasping as _gog
    import mj_grasp_sim_b200.mgs.core.simualtion as _simmod

    _orig = mlib.BatchSim
    _libs = {f: mlib.bind(C.CDLL(lane1.build(f)), prefix="l1_") for f in (False, True)}

    class ProxySim(_orig):
        def __init__(self, model, device=0, f64=False, lib=None, prefix="mgs_", ncon_max=0, nefc_max=0, ground_name="geom:ground"):
            super().__init__(model, lib=_libs[bool(f64)], prefix="l1_", ground_name=ground_name)

        def overflow_count(self):
            return 0

    mlib.BatchSim = _gog.BatchSim = _ct.BatchSim = _simmod.BatchSim = ProxySim
    mlib.load = lambda f64=False: None

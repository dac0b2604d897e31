"""Host logic of the capacity ladder (mgs.env EscalatingSim) on fake sims: first pass, the default capacity of a single-object scene
when the first pass was cut below it, then the largest capacity that fits; every candidate ends on the first rung that holds its
contacts, and what is still over at the top is counted and warned about."""
import types
import warnings

import numpy as np
import pytest

from mj_grasp_sim_b200.mgs.env.gravityless_object_grasping import EscalatingSim


class FakeSim:
    """labels = the capacity the candidate ran on; a candidate overflows when it needs more contacts than that"""
    created = []

    def __init__(self, cap, needs):
        self.info = types.SimpleNamespace(ncon_max=cap)
        self.needs, self.last, self.closed = np.asarray(needs), None, False
        FakeSim.created.append(cap)

    def run(self, idx):
        self.last = np.arange(len(self.needs)) if idx is None else np.asarray(idx)
        return np.full(len(self.last), self.info.ncon_max), self.last.copy()

    def last_aux(self, n):
        assert n == len(self.last)
        over = self.needs[self.last] > self.info.ncon_max
        return dict(overflow=over, bad=np.zeros(n, bool), pos_drift=np.full(n, float(self.info.ncon_max)), rot_drift_deg=np.zeros(n))

    def close(self):
        self.closed = True


def _ladder(first_cap, needs, fits=lambda cap: True):
    FakeSim.created = []

    def make(caps):
        cap = first_cap if caps is None else caps[0]
        if caps is not None and not fits(cap):
            raise RuntimeError("does not fit")
        return FakeSim(cap, needs)
    return EscalatingSim(make)


def test_every_candidate_ends_on_the_first_rung_that_holds_it():
    needs = [3, 10, 40, 12, 200, 4]
    E = _ladder(4, needs)
    (cap_used, idx), aux = E.run(lambda sim, i: sim.run(i), len(needs))
    assert cap_used.tolist() == [4, 32, 256, 32, 256, 4] and idx.tolist() == list(range(6))
    assert E.last_overflow == dict(first_pass=4, after_escalation=0) and not aux["overflow"].any()
    assert aux["pos_drift"].tolist() == [4, 32, 256, 32, 256, 4]  # per-candidate outputs follow the rung that produced the label
    assert FakeSim.created == [4, 32, 256]
    E.close()


def test_ladder_skips_the_mid_rung_when_the_first_pass_already_has_it_and_stops_when_nothing_is_over():
    E = _ladder(32, [3, 40, 10])
    (cap_used, _), _ = E.run(lambda sim, i: sim.run(i), 3)
    assert cap_used.tolist() == [32, 256, 32] and FakeSim.created == [32, 256]
    E2 = _ladder(12, [3, 10, 11])
    (cap_used, _), _ = E2.run(lambda sim, i: sim.run(i), 3)
    assert cap_used.tolist() == [12, 12, 12] and FakeSim.created == [12] and E2.last_overflow == dict(first_pass=0, after_escalation=0)
    E3 = _ladder(16, [3, 20, 30])  # the fp64 Allegro case: over 16, under 32 -> the big instance is never built
    (cap_used, _), _ = E3.run(lambda sim, i: sim.run(i), 3)
    assert cap_used.tolist() == [16, 32, 32] and FakeSim.created == [16, 32]


def test_largest_capacity_that_fits_and_a_warning_for_what_is_left():
    needs = [3, 100, 500, 70]
    E = _ladder(4, needs, fits=lambda cap: cap <= 96)  # 256 and 128 do not fit one CTA's shared memory
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        (cap_used, _), aux = E.run(lambda sim, i: sim.run(i), len(needs))
    assert cap_used.tolist() == [4, 96, 96, 96] and FakeSim.created == [4, 32, 96]
    assert E.last_overflow == dict(first_pass=3, after_escalation=2) and aux["overflow"].tolist() == [False, True, True, False]
    assert len(w) == 1 and "2 of 4 environments needed more than 96 contacts" in str(w[0].message)
    # nothing above the first pass fits: labels of the first pass, counted and warned about
    E = _ladder(40, [3, 100], fits=lambda cap: False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        (cap_used, _), aux = E.run(lambda sim, i: sim.run(i), 2)
    assert cap_used.tolist() == [40, 40] and aux["overflow"].tolist() == [False, True] and len(w) == 1


def test_the_label_tool_climbs_the_same_ladder():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import label_agreement as la
    needs = [3, 10, 40, 12, 200, 4, 1000]
    FakeSim.created = []

    def make(nc):
        if nc > 128:
            raise RuntimeError("does not fit")
        return FakeSim(nc, needs)
    (cap_used, idx), first, left = la.run_escalated(FakeSim(4, needs), make, lambda sim, i: sim.run(i))
    assert cap_used.tolist() == [4, 32, 128, 32, 128, 4, 128] and idx.tolist() == list(range(7))
    assert (first, left) == (5, 2) and FakeSim.created == [4, 32, 128]  # 200 and 1000 contacts do not fit 128 either
    (cap_used, _), first, left = la.run_escalated(FakeSim(32, [3, 40]), make, lambda sim, i: sim.run(i))  # first pass at the default: no mid rung
    assert cap_used.tolist() == [32, 128] and (first, left) == (1, 0)

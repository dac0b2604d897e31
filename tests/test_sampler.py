"""Antipodal sampler mirror (reference mgs/sampler/antipodal.py) and the gen_grasp_candidates CLI: geometric
invariants of the batched restatement (CPU; the ray casting is a torch expression that also runs on the GPU)."""
import numpy as np
import pytest

from mj_grasp_sim_b200.mgs.cli import gen_grasp_candidates
from mj_grasp_sim_b200.mgs.obj.selector import get_object
from mj_grasp_sim_b200.mgs.sampler.antipodal import AntipodalGraspGenerator, _vmf3


def _contacts(H, w):
    c, x = H[:, :3, 3], H[:, :3, 0]
    return c - 0.5 * w[:, None] * x, c + 0.5 * w[:, None] * x


def test_vmf_concentrates_around_mu():
    rng = np.random.default_rng(0)
    mu = rng.normal(size=(20000, 3))
    mu /= np.linalg.norm(mu, axis=1, keepdims=True)
    d = _vmf3(mu, 10.0, rng)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0)
    assert abs((d * mu).sum(axis=1).mean() - (1.0 / np.tanh(10.0) - 0.1)) < 5e-3  # E[mu . x] = coth(k) - 1/k


def test_cube_grasps_touch_the_surface_and_frames_are_right_handed():
    g = AntipodalGraspGenerator(get_object("cube"), device="cpu", seed=3)
    H, aux = g.generate_grasps(3000)
    w = aux["width"]
    assert H.shape == (3000, 4, 4) and w.shape == (3000,) and H.dtype == np.float64
    R = H[:, :3, :3]
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-12 and np.linalg.det(R).min() > 0.999999
    ok = ~aux["fallback"]
    assert ok.mean() > 0.99  # a closed convex body: the inward ray always leaves through another face
    p1, p2 = _contacts(H[ok], w[ok])
    for p in (p1, p2):  # both contacts on the boundary of the 4 cm cube
        assert np.abs(np.abs(p).max(axis=1) - 0.02).max() < 1e-9
    assert (w[ok] > 0).all() and w.max() <= 0.04 * np.sqrt(3) + 1e-9
    # antipodal: the closing axis is within ~60 degrees of the inward normal at the first contact (kappa = 10)
    n1 = -np.sign(p1) * (np.abs(np.abs(p1) - 0.02) < 1e-9)
    cosang = (H[ok][:, :3, 0] * n1).sum(axis=1) / np.maximum(np.linalg.norm(n1, axis=1), 1e-12)
    assert np.median(cosang) > 0.85


def test_hull_grasps_lie_on_the_hull_and_are_reproducible():
    obj = get_object("hull:3")
    v, t = obj.mesh()
    H, aux = AntipodalGraspGenerator(obj, device="cpu", seed=7).generate_grasps(1500)
    H2, aux2 = AntipodalGraspGenerator(obj, device="cpu", seed=7).generate_grasps(1500)
    H3, _ = AntipodalGraspGenerator(obj, device="cpu", seed=8).generate_grasps(1500)
    assert np.array_equal(H, H2) and np.array_equal(aux["width"], aux2["width"]) and not np.array_equal(H, H3)
    T = v[t]
    fn = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    d0 = (fn * T[:, 0]).sum(axis=1)
    ok = ~aux["fallback"]
    p1, p2 = _contacts(H[ok], aux["width"][ok])
    # the reference maps normalised points back with the UNSCALED centroid as offset (docstring): undo that known shift
    g = AntipodalGraspGenerator(obj, device="cpu")
    g.normalize_load()
    shift = (-g.offset) * (g.scale - 1.0)
    for p in (p1, p2):
        sd = (p - shift) @ fn.T - d0[None]  # signed distance to every face plane
        assert sd.max() < 1e-9 and np.abs(sd.max(axis=1)).max() < 1e-9  # inside all planes and on at least one


def test_gen_grasp_candidates_cli(tmp_path):
    from mj_grasp_sim_b200.mgs.env.gravityless_object_grasping import GravitylessObjectGrasping
    from mj_grasp_sim_b200.mgs.gripper.selector import get_gripper
    from mj_grasp_sim_b200.mgs.util.geo.transforms import SE3Pose
    H, joints = gen_grasp_candidates.run("PandaGripper", "hull:0", 256, str(tmp_path), seed=0)
    f = np.load(tmp_path / "PandaGripper" / "hull:0" / "candidates.npz")
    assert f["pose"].shape == (256, 4, 4) and f["joints"].shape == (256, 2)
    assert (f["joints"][:, 0] >= 0).all() and (f["joints"][:, 0] <= 0.04).all() and (f["joints"][:, 1] <= 0).all() and (f["joints"][:, 1] >= -0.04).all()
    # the file feeds the filter entry points unchanged: same loading path as filter_to_stable (from_mat + _process)
    env = GravitylessObjectGrasping(get_gripper("PandaGripper"), get_object("hull:0"))
    pose7, j32, jadr = env._process(SE3Pose.from_mat(f["pose"], type="wxyz"), f["joints"])
    assert pose7.shape == (256, 7) and j32.shape == (256, 2) and np.isfinite(pose7).all()
    with pytest.raises(NotImplementedError):  # Allegro has neither an antipodal nor a contact-based producer in the reference
        gen_grasp_candidates.run("AllegroGripper", "hull:0", 8, str(tmp_path))

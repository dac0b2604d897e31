"""Contact-based differentiable sampler (reference mgs/sampler/contact.py + kin/*.py) in PyTorch: the hand
kinematics derived from the compiled model against the oracle's kinematics, and the optimisation's invariants (CPU)."""
import numpy as np
import pytest

from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.compiler.mjcf import compile_mjcf
from mj_grasp_sim_b200.mgs.obj.selector import get_object
from mj_grasp_sim_b200.mgs.sampler.contact import ContactBasedDiff, farthest_point_sampling, matrix_to_rotation_6d, rotation_6d_to_matrix
from mj_grasp_sim_b200.mgs.sampler.kin import HandKinematics
from oracle.oracle import OracleSim


@pytest.mark.parametrize("name", ["leap", "shadow"])
def test_hand_fk_matches_the_oracle_kinematics(name):
    """link frames of the torch FK (relative to the base body) = the oracle's body poses of the same compiled model at the same joints"""
    import torch
    kin = HandKinematics(name)
    g = scenes.GRIPPERS[name]
    gx, ga = scenes.gripper_fragment(name)
    m = compile_mjcf("<mujoco><compiler angle='radian' autolimits='true'/><option integrator='implicitfast'/>" + gx + "</mujoco>", ga)
    rng = np.random.default_rng(0)
    theta = rng.uniform(kin.joint_ranges[:, 0], kin.joint_ranges[:, 1]).astype(np.float64)
    s = OracleSim(m)
    s.reset()
    for k, j in enumerate(g["joints"]):
        s.qpos[int(m.jnt_qposadr[m.names["joint"][j]])] = theta[k]
    s.kinematics()
    R, p = kin.fk(torch.as_tensor(theta[None], dtype=torch.float64))
    base = int(m.jnt_bodyid[m.names["joint"][g["freejoint"]]])
    xpos, xmat = np.asarray(s.xpos).reshape(-1, 3), np.asarray(s.xmat).reshape(-1, 3, 3)
    for lb, b in enumerate(kin.bodies):
        assert np.abs(xmat[base].T @ (xpos[b] - xpos[base]) - p[0, lb].numpy()).max() < 1e-6
        assert np.abs(xmat[base].T @ xmat[b] - R[0, lb].numpy()).max() < 1e-6
    assert kin.num_dofs == len(g["joints"]) and len(kin.fingertip_body) == (4 if name == "leap" else 5)
    # pad points sit on the pad side of the distal link: on the face opposite to the stored (negative) normal
    for f in range(len(kin.fingertip_body)):
        assert (kin.local_fingertip_contact_positions[f] @ kin.neg_normal[f] <= 1e-3).all()


def test_rotation_6d_round_trip_and_fps():
    import torch
    from scipy.spatial.transform import Rotation
    R = torch.as_tensor(Rotation.random(16, random_state=1).as_matrix(), dtype=torch.float64)
    assert (rotation_6d_to_matrix(matrix_to_rotation_6d(R)) - R).abs().max() < 1e-12
    x = np.random.default_rng(2).normal(size=(500, 3))
    idx = farthest_point_sampling(x, 20)
    assert idx[0] == 0 and len(set(idx.tolist())) == 20
    d = np.linalg.norm(x[idx][:, None] - x[idx][None], axis=-1) + np.eye(20) * 1e9
    assert d.min() > 0.5 * np.median(np.linalg.norm(x[:, None] - x[None], axis=-1))  # well spread


@pytest.mark.parametrize("name", ["leap", "shadow"])
def test_contact_sampler_pulls_the_fingertips_onto_their_targets(name):
    obj = get_object("hull:3")
    kin = HandKinematics(name)
    gen = ContactBasedDiff(obj, device="cpu", seed=5)
    import mj_grasp_sim_b200.mgs.sampler.contact as C
    first = None
    for steps in (1, 150):
        C.N_STEPS = steps
        g2 = ContactBasedDiff(obj, device="cpu", seed=5)
        H, aux = g2.generate_grasps(24, kin)
        if first is None:
            first = g2.last_losses.copy()
    C.N_STEPS = 150
    assert H.shape == (24, 4, 4) and aux["joints"].shape == (24, kin.num_dofs) and H.dtype == np.float64
    R = H[:, :3, :3]
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-5 and np.linalg.det(R).min() > 0.999
    assert (aux["joints"] >= kin.joint_ranges[:, 0] - 1e-6).all() and (aux["joints"] <= kin.joint_ranges[:, 1] + 1e-6).all()
    assert np.median(g2.last_losses) < 0.25 * np.median(first)  # 150 steps reduce the median loss at least four times
    # same seed, same result
    H2, aux2 = ContactBasedDiff(obj, device="cpu", seed=5).generate_grasps(24, kin)
    assert np.array_equal(H, H2) and np.array_equal(aux["joints"], aux2["joints"])


def test_gen_grasp_candidates_cli_for_a_hand(tmp_path):
    from mj_grasp_sim_b200.mgs.cli import gen_grasp_candidates
    H, joints = gen_grasp_candidates.run("ShadowHand", "hull:0", 8, str(tmp_path), seed=1)
    f = np.load(tmp_path / "ShadowHand" / "hull:0" / "candidates.npz")
    assert f["pose"].shape == (8, 4, 4) and f["joints"].shape == (8, 22)

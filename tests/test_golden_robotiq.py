"""Known-answer test against the only MuJoCo-produced numbers the reference ships on this path:
the Robotiq 2F-85 fully-closed mjSTATE_INTEGRATION vector
(/root/reference/mgs/cli/config/gripper/robotiq_2f_85.yaml:11, extracted by tools/extract_golden.py).

A lone Robotiq in zero gravity with ctrl = 255 and the mocap at (0, 0, -0.15) settles into the 4-bar
linkage equilibrium MuJoCo 3.2.2 recorded.  Reproducing it exercises hinge kinematics, the two connect
equalities and the joint equality, the tendon-driven general actuator, the spring-link joint spring, the
pad-pad contact and the soft-constraint parameters.  Tolerance 5e-4 rad (the recorded state comes from a
different scene file - camera body, lights - and MuJoCo's own box-box manifold)."""
import json
import os

import numpy as np
import pytest

from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.compiler.mjcf import compile_mjcf
from oracle.oracle import OracleSim

HERE = os.path.dirname(os.path.abspath(__file__))

SCAN_XML = """<mujoco><compiler angle="radian" autolimits="true"/>
<option integrator="implicitfast" timestep="0.001" cone="elliptic" impratio="3" noslip_iterations="2"
        noslip_tolerance="1e-8" tolerance="1e-8" gravity="0 0 0"><flag multiccd="enable"/></option>
{gripper}
<worldbody><body name="center" pos="0 0 0"><geom name="geom:center" size="0.000001"/></body>
<body name="body:camera" pos="0 0 .4"><freejoint name="camera:joint"/><geom name="geom:camera" size="0.01"/></body>
</worldbody></mujoco>"""


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "robotiq_2f85_state_close.json")))


@pytest.fixture(scope="module")
def scan_model():
    gx, ga = scenes.gripper_fragment("robotiq2f85")
    return compile_mjcf(SCAN_XML.format(gripper=gx), ga)


def test_state_vector_layout(golden, scan_model):
    m = scan_model
    st = golden["state"]
    # mjSTATE_INTEGRATION: time, qpos, qvel, act, qacc_warmstart, ctrl, qfrc_applied, xfrc_applied, eq_active, mocap_pos, mocap_quat
    n = 1 + m.nq + m.nv + 0 + m.nv + m.nu + m.nv + 6 * m.nbody + int(m.arr["neq"]) + 3 + 4
    assert (m.nq, m.nv, m.nu, m.nbody, int(m.arr["neq"])) == (22, 20, 1, 18, 4)
    assert len(st) == n == 203
    assert st[1 + m.nq + 2 * m.nv] == 255.0  # ctrl
    assert st[-11:-7] == [1, 1, 1, 1] and np.allclose(st[-7:-4], [0, 0, -0.15], atol=1e-7) and st[-4:] == [1, 0, 0, 0]


def test_oracle_reproduces_mujoco_closed_state(golden, scan_model):
    m = scan_model
    g = np.array(golden["state"])
    q_gold = g[1:1 + m.nq]
    s = OracleSim(m)
    s.reset()
    s.qpos[0:3] = [0, 0, -0.15]
    s.mocap_pos[0] = [0, 0, -0.15]
    s.ctrl[:] = 255.0
    s.step(2500)
    assert s.bad == 0
    jn = m.names["joint"]
    for name in ("right_driver_joint", "right_coupler_joint", "right_spring_link_joint", "right_follower_joint",
                 "left_driver_joint", "left_coupler_joint", "left_spring_link_joint", "left_follower_joint"):
        a = m.jnt_qposadr[jn[name]]
        assert abs(s.qpos[a] - q_gold[a]) < 5e-4, (name, s.qpos[a], q_gold[a])
    assert np.abs(s.qpos[:3] - q_gold[:3]).max() < 1e-5 and np.abs(s.qvel).max() < 1e-3  # at rest at the mocap target

"""Drop-in mirror of the reference's `mgs` package for the grasp-evaluation hot path.

Same class / method names and argument meaning as /root/reference/mgs (env, gripper, obj, core, util),
but `GravitylessObjectGrasping.grasp_collision_mask` / `.grasp_stability_evaluation_from_joints` run
all candidates in one batched launch of the B200 kernels instead of a Python loop over
`mujoco.mj_step`.  Use `from mj_grasp_sim_b200 import mgs` (or put `mj_grasp_sim_b200/` on sys.path and
`import mgs`).
"""

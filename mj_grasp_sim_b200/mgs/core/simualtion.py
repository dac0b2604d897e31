"""Sim handle (file name keeps the reference's spelling, /root/reference/mgs/core/simualtion.py).

The reference's MjSimulation wraps one MjModel + MjData.  Here `model` is the compiled flat model and
the batched state lives on the GPU inside the C-ABI library, so the handle exposes the model-side
queries the hot path uses (`get_joint_idxs`) with the same semantics - including the reference's
unknown-name quirk: `mj_name2id` returns -1 and `jnt_qposadr[-1]` is the LAST joint's address
(simualtion.py:37-43; hit by Robotiq's two misnamed joints)."""
from __future__ import annotations

from typing import List

import numpy as np


class MjSimulation:
    model = None

    def get_joint_idxs(self, joint_list: List[str]) -> List[int]:
        out = []
        for j in joint_list:
            jid = self.model.names["joint"].get(j, -1)
            out.append(int(self.model.jnt_qposadr[jid]))
        return out

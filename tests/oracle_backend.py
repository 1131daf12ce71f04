"""Test-only backend for roboschool_b200.cpp_household: drives the shim from the CPU oracle so the
drop-in boundary can be exercised (against the reference's own Python env code) without a GPU."""
import numpy as np

from oracle.oracle import OracleWorld


class OracleBackend:
    def __init__(self, model, n_envs=1, device=0):
        assert n_envs == 1
        self.model = model
        self.o = OracleWorld(model, 1)
        self.S = self.o.sdim
        s = np.zeros(self.S)
        if not model.fixed_base:
            s[3:7] = [0, 0, 0, 1]
        self.o.set_state(0, s)

    def set_state(self, state, clear_contacts=True):
        self.o.set_state(0, np.asarray(state, dtype=np.float64).reshape(self.S), clear_contacts)

    def get_state(self):
        return self.o.get_state(0).reshape(1, self.S).astype(np.float64)

    def substeps(self, n, tau=None):
        if tau is not None:
            self.o.set_tau(0, np.asarray(tau, dtype=np.float64).reshape(-1))
        self.o.substeps(0, int(n))

    def set_joint_motor(self, dof, kp, kd, target_pos, target_vel, max_impulse, env=-1):
        self.o.set_joint_motor(0, int(dof), float(kp), float(kd), float(target_pos), float(target_vel), float(max_impulse))

    def link_poses(self):
        return self.o.link_poses(0)[None]

    def contacts(self, env, max_pts=64):
        c = self.o.contacts(0, max_pts)
        return c[:, 0].astype(np.int32), c[:, 2:12]

    def close(self):
        pass

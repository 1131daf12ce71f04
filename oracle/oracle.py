"""ctypes wrapper over oracle/liboracle.so (CPU restatement; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  See rs_oracle.c for the citations and the
"parity unpinned" statement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rs_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "rs_model.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        U = ctypes.c_uint32
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_create.restype = P
        L.oracle_create.argtypes = [P, I]
        L.oracle_destroy.argtypes = [P]
        L.oracle_state_dim.argtypes = [P]
        L.oracle_set_state.argtypes = [P, I, dp, I]
        L.oracle_get_state.argtypes = [P, I, dp]
        L.oracle_set_episode.argtypes = [P, I, D, D, dp, I]
        L.oracle_get_episode.argtypes = [P, I, dp, dp, dp, ip]
        L.oracle_reset.argtypes = [P, I, U, U]
        L.oracle_reset_all.argtypes = [P, U, U]
        L.oracle_set_tau.argtypes = [P, I, dp]
        L.oracle_substeps.argtypes = [P, I, I]
        L.oracle_step.argtypes = [P, dp, I, U, U, I]
        L.oracle_get_obs.argtypes = [P, dp, dp, ip]
        L.oracle_get_rewards5.argtypes = [P, I, dp]
        L.oracle_link_poses.argtypes = [P, I, dp]
        L.oracle_get_contacts.argtypes = [P, I, dp, I]
        L.oracle_get_contacts.restype = I
        L.oracle_counts.argtypes = [P, I, ip, ip]
        L.oracle_get_contact_cache.argtypes = [P, I, ip, dp, I]
        L.oracle_get_contact_cache.restype = I
        L.oracle_set_knob.argtypes = [ctypes.c_char_p, D]
        L.oracle_accel_deltas.argtypes = [P, I, dp, dp]
        L.oracle_energy.argtypes = [P, I]
        L.oracle_energy.restype = D
        L.oracle_foot_flags.argtypes = [P, I, ip, ip]
        L.oracle_set_flag.argtypes = [P, I, D, D, I, U]
        L.oracle_set_joint_motor.argtypes = [P, I, I, D, D, D, D, D]
        L.oracle_get_flag.argtypes = [P, I, dp, dp, ip, ctypes.POINTER(U)]
        L.oracle_get_task_state.argtypes = [P, I, dp]
        L.oracle_set_task_state.argtypes = [P, I, dp]
        L.oracle_hash.argtypes = [U, U, U, U]
        L.oracle_hash.restype = U
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class OracleWorld:
    """N independent envs stepped by the double-precision CPU restatement."""

    def __init__(self, model, n_envs=1):
        self.L = lib()
        self.model = model
        self.n = n_envs
        self.h = self.L.oracle_create(ctypes.addressof(model), n_envs)
        self.sdim = self.L.oracle_state_dim(self.h)

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def set_state(self, env, s, clear_contacts=True):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.sdim,)
        self.L.oracle_set_state(self.h, env, _dp(s), int(clear_contacts))

    def get_state(self, env=None):
        if env is None:
            return np.stack([self.get_state(i) for i in range(self.n)])
        s = np.zeros(self.sdim)
        self.L.oracle_get_state(self.h, env, _dp(s))
        return s

    def set_episode(self, env, potential, initial_z, feet_contact, frame=0):
        fc = np.zeros(8)
        fc[:len(feet_contact)] = feet_contact
        self.L.oracle_set_episode(self.h, env, potential, initial_z, _dp(fc), frame)

    def get_episode(self, env):
        pot, iz = ctypes.c_double(), ctypes.c_double()
        fr = ctypes.c_int()
        fc = np.zeros(8)
        self.L.oracle_get_episode(self.h, env, ctypes.byref(pot), ctypes.byref(iz), _dp(fc), ctypes.byref(fr))
        return pot.value, iz.value, fc[:self.model.n_feet].copy(), fr.value

    def reset(self, seed=0, env_id0=0):
        self.seed, self.env_id0 = seed, env_id0     # the flag RNG stream of later steps continues from these
        self.L.oracle_reset_all(self.h, seed, env_id0)

    def set_joint_motor(self, env, dof, kp, kd, target_pos, target_vel, max_impulse):
        """btMultiBodyJointMotor of one dof (env < 0: every env); max_impulse per sub-step, 0 = off."""
        self.L.oracle_set_joint_motor(self.h, env, dof, kp, kd, target_pos, target_vel, max_impulse)

    def set_flag(self, env, tx, ty, timeout, draws):
        self.L.oracle_set_flag(self.h, env, tx, ty, int(timeout), int(draws))

    def get_task_state(self, env):
        s = np.zeros(8); self.L.oracle_get_task_state(self.h, env, _dp(s)); return s

    def set_task_state(self, env, s):
        s = np.ascontiguousarray(s, dtype=np.float64); self.L.oracle_set_task_state(self.h, env, _dp(s))

    def get_flag(self, env):
        tx, ty = ctypes.c_double(), ctypes.c_double()
        to, dr = ctypes.c_int(), ctypes.c_uint32()
        self.L.oracle_get_flag(self.h, env, ctypes.byref(tx), ctypes.byref(ty), ctypes.byref(to), ctypes.byref(dr))
        return tx.value, ty.value, to.value, dr.value

    def set_tau(self, env, tau):
        t = np.zeros(self.model.n_dof)
        t[:] = tau
        self.L.oracle_set_tau(self.h, env, _dp(t))

    def substeps(self, env, n=1):
        self.L.oracle_substeps(self.h, env, n)

    def step(self, actions, auto_reset=False, seed=None, env_id0=None, threads=0):
        seed = getattr(self, "seed", 0) if seed is None else seed
        env_id0 = getattr(self, "env_id0", 0) if env_id0 is None else env_id0
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, self.model.n_act)
        self.L.oracle_step(self.h, _dp(a), int(auto_reset), seed, env_id0, threads)
        return self.obs()

    def obs(self):
        O = self.model.obs_dim
        obs = np.zeros((self.n, O))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.int32)
        self.L.oracle_get_obs(self.h, _dp(obs), _dp(rew), _ip(done))
        return obs, rew, done

    def rewards5(self, env):
        r = np.zeros(5)
        self.L.oracle_get_rewards5(self.h, env, _dp(r))
        return r

    def link_poses(self, env):
        """(1+L, 10): pos3, quat4 (xyzw), linear velocity3 of every body COM frame, [0] = base."""
        out = np.zeros((self.model.n_links + 1, 10))
        self.L.oracle_link_poses(self.h, env, _dp(out))
        return out

    def contacts(self, env, max_pts=128):
        out = np.zeros((max_pts, 12))
        c = self.L.oracle_get_contacts(self.h, env, _dp(out), max_pts)
        return out[:c]

    def contact_cache(self, max_pts=16):
        """Persistent manifolds of every env as the CUDA path stores them: counts [n], keys [n,P], data [n,P,10]."""
        cnt = np.zeros(self.n, dtype=np.int32)
        keys = np.zeros((self.n, max_pts), dtype=np.int32)
        data = np.zeros((self.n, max_pts, 10))
        for i in range(self.n):
            cnt[i] = self.L.oracle_get_contact_cache(self.h, i, _ip(keys[i]), _dp(data[i]), max_pts)
        return cnt, keys, data

    def accel_deltas(self, env, force):
        """M^-1 force of the env's current configuration (btMultiBody::calcAccelerationDeltasMultiDof)."""
        f = np.ascontiguousarray(force, dtype=np.float64)
        out = np.zeros_like(f)
        self.L.oracle_accel_deltas(self.h, env, _dp(f), _dp(out))
        return out

    def counts(self, env):
        a, b = ctypes.c_int(), ctypes.c_int()
        self.L.oracle_counts(self.h, env, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def foot_flags(self, env):
        g = np.zeros(8, dtype=np.int32)
        o = np.zeros(8, dtype=np.int32)
        self.L.oracle_foot_flags(self.h, env, _ip(g), _ip(o))
        return g[:self.model.n_feet].copy(), o[:self.model.n_feet].copy()

    def energy(self, env):
        return self.L.oracle_energy(self.h, env)


def mlp_forward(weights, x):
    """agent_zoo v1 policy: relu(relu(x@W1+b1)@W2+b2)@W3+b3
    (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34), numpy float64."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(w, dtype=np.float64) for w in weights]
    h = np.maximum(x @ W1 + b1, 0)
    h = np.maximum(h @ W2 + b2, 0)
    return h @ W3 + b3

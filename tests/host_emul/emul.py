"""ctypes wrapper for tests/host_emul/libemul.so: the CUDA per-env code (rs_core.cuh, fp32)
compiled for the host.  Test infrastructure only (GPU-less container debugging)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libemul.so")
    srcs = [os.path.join(_HERE, "emul.cpp"),
            os.path.join(_HERE, "..", "..", "roboschool_b200", "csrc", "rs_core.cuh"),
            os.path.join(_HERE, "..", "..", "roboschool_b200", "csrc", "rs_devmodel.h"),
            os.path.join(_HERE, "..", "..", "roboschool_b200", "csrc", "rs_static.cuh"),
            os.path.join(_HERE, "..", "..", "roboschool_b200", "csrc", "rs_pooled.cuh"),
            os.path.join(_HERE, "..", "..", "roboschool_b200", "csrc", "rs_topo_gen.h"),
            os.path.join(_HERE, "..", "..", "include", "rs_model.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                               "-o", so, srcs[0]])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        P, I, D, U = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint32
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
        L.emul_create.restype = P; L.emul_create.argtypes = [P, I]
        L.emul_destroy.argtypes = [P]
        L.emul_use_static.argtypes = [P, I]
        L.emul_state_dim.argtypes = [P]
        L.emul_set_state.argtypes = [P, I, dp, I]
        L.emul_get_state.argtypes = [P, I, dp]
        L.emul_set_tau.argtypes = [P, I, dp]
        L.emul_set_joint_motor.argtypes = [P, I, I, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.emul_substeps.argtypes = [P, I, I]
        L.emul_reset_all.argtypes = [P, U, U]
        L.emul_set_episode.argtypes = [P, I, D, D, dp, I]
        L.emul_step.argtypes = [P, dp, I, U, U]
        L.emul_get_obs.argtypes = [P, dp, dp, ip]
        L.emul_get_rewards5.argtypes = [P, I, dp]
        L.emul_get_contacts.argtypes = [P, I, dp, I]; L.emul_get_contacts.restype = I
        L.emul_counts.argtypes = [P, I, ip, ip]
        L.emul_set_contacts.argtypes = [P, I, I, ip, dp]
        L.emul_link_poses.argtypes = [P, I, dp]
        L.emul_set_flag.argtypes = [P, I, D, D, I, U]
        L.emul_get_flag.argtypes = [P, I, dp]
        L.emul_set_task_state.argtypes = [P, I, dp]
        L.emul_get_task_state.argtypes = [P, I, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class EmulWorld:
    def __init__(self, model, n_envs=1):
        self.L = lib(); self.model = model; self.n = n_envs
        self.h = self.L.emul_create(ctypes.addressof(model), n_envs)
        assert self.h
        self.sdim = self.L.emul_state_dim(self.h)
        self.topo = 0

    def use_static(self, on=True):
        """Switch to the static-topology sweeps (rs_static.cuh); returns the matched topology id (0 = none)."""
        self.topo = self.L.emul_use_static(self.h, int(on))
        return self.topo

    def __del__(self):
        try: self.L.emul_destroy(self.h)
        except Exception: pass

    def set_state(self, env, s, clear_contacts=True):
        s = np.ascontiguousarray(s, dtype=np.float64); self.L.emul_set_state(self.h, env, _dp(s), int(clear_contacts))

    def get_state(self, env):
        s = np.zeros(self.sdim); self.L.emul_get_state(self.h, env, _dp(s)); return s

    def set_tau(self, env, tau):
        t = np.zeros(self.model.n_dof); t[:] = tau; self.L.emul_set_tau(self.h, env, _dp(t))

    def set_joint_motor(self, env, dof, kp, kd, target_pos, target_vel, max_impulse):
        self.L.emul_set_joint_motor(self.h, env, dof, kp, kd, target_pos, target_vel, max_impulse)

    def substeps(self, env, n=1): self.L.emul_substeps(self.h, env, n)

    def reset(self, seed=0, env_id0=0):
        self.seed, self.env_id0 = seed, env_id0
        self.L.emul_reset_all(self.h, seed, env_id0)

    def set_flag(self, env, tx, ty, timeout, draws): self.L.emul_set_flag(self.h, env, tx, ty, int(timeout), int(draws))

    def set_task_state(self, env, s):
        s = np.ascontiguousarray(s, dtype=np.float64); self.L.emul_set_task_state(self.h, env, _dp(s))

    def get_task_state(self, env):
        s = np.zeros(8); self.L.emul_get_task_state(self.h, env, _dp(s)); return s

    def get_flag(self, env):
        o = np.zeros(4); self.L.emul_get_flag(self.h, env, _dp(o)); return o[0], o[1], int(o[2]), int(o[3])

    def set_episode(self, env, potential, initial_z, feet_contact, frame=0):
        fc = np.zeros(8); fc[:len(feet_contact)] = feet_contact
        self.L.emul_set_episode(self.h, env, potential, initial_z, _dp(fc), frame)

    def step(self, actions, auto_reset=False, seed=None, env_id0=None):
        seed = getattr(self, "seed", 0) if seed is None else seed
        env_id0 = getattr(self, "env_id0", 0) if env_id0 is None else env_id0
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, self.model.n_act)
        self.L.emul_step(self.h, _dp(a), int(auto_reset), seed, env_id0)
        return self.obs()

    def obs(self):
        O = self.model.obs_dim
        obs = np.zeros((self.n, O)); rew = np.zeros(self.n); done = np.zeros(self.n, dtype=np.int32)
        self.L.emul_get_obs(self.h, _dp(obs), _dp(rew), done.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        return obs, rew, done

    def rewards5(self, env):
        r = np.zeros(5); self.L.emul_get_rewards5(self.h, env, _dp(r)); return r

    def contacts(self, env, max_pts=128):
        out = np.zeros((max_pts, 12)); c = self.L.emul_get_contacts(self.h, env, _dp(out), max_pts); return out[:c]

    def set_contacts(self, env, count, keys, data):
        k = np.ascontiguousarray(keys, dtype=np.int32); d = np.ascontiguousarray(data, dtype=np.float64)
        self.L.emul_set_contacts(self.h, env, int(count), k.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(d))

    def counts(self, env):
        a, b = ctypes.c_int(), ctypes.c_int(); self.L.emul_counts(self.h, env, ctypes.byref(a), ctypes.byref(b)); return a.value, b.value

    def link_poses(self, env):
        out = np.zeros((self.model.n_links + 1, 10)); self.L.emul_link_poses(self.h, env, _dp(out)); return out

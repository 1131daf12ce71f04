"""BatchedWorld: thin Python handle over the C ABI (one compiled model, N independent envs on one GPU).

Host-side mirror of Household::World for the batched path
(roboschool/cpp-household/household.h:138-205, physics-bullet.cpp:358-412).
numpy buffers are HOST buffers; torch CUDA tensors are passed by data_ptr().
"""
import ctypes

import numpy as np

from . import _lib
from .model import ModelDesc


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()   # torch tensor


class BatchedWorld:
    def __init__(self, model: ModelDesc, n_envs: int, device: int = 0):
        self.L = _lib.lib()
        self.model = model
        self.n = int(n_envs)
        self.device = device
        h = ctypes.c_void_p()
        _lib.check(self.L.rs_world_create(ctypes.addressof(model), self.n, device, ctypes.byref(h)))
        self.h = h
        self.S = self.L.rs_world_state_dim(h)
        self.O = model.obs_dim
        self.A = model.n_act

    def close(self):
        if getattr(self, "h", None):
            self.L.rs_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-tensor path (torch) ----
    def reset(self, seed=0, env_id0=0, mask=None, obs=None, stream=None):
        _lib.check(self.L.rs_world_reset(self.h, seed, env_id0, _ptr(mask), _ptr(obs), stream))

    def step(self, actions, obs, reward, done, auto_reset=True, stream=None):
        _lib.check(self.L.rs_world_step(self.h, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), int(auto_reset), stream))

    # ---- host-buffer path (numpy): the reference's calling convention ----
    def step_host(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, self.A)
        obs = np.empty((self.n, self.O), dtype=np.float32)
        rew = np.empty(self.n, dtype=np.float32)
        done = np.empty(self.n, dtype=np.uint8)
        _lib.check(self.L.rs_world_step_host(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, int(auto_reset)))
        return obs, rew, done

    def pinned_host_buffers(self):
        """(obs [n,O] f32, reward [n] f32, done [n] u8) numpy views of page-locked host memory: passed as `out=` (and as
        the next call's obs_in) they are copied to/from the device directly, without the library's staging memcpy."""
        import torch
        t = (torch.empty((self.n, self.O), dtype=torch.float32).pin_memory(), torch.empty(self.n, dtype=torch.float32).pin_memory(),
             torch.empty(self.n, dtype=torch.uint8).pin_memory())
        self._pinned = getattr(self, "_pinned", []) + [t]     # keep the torch owners alive
        return tuple(x.numpy() for x in t)

    def policy_step_host(self, policy, obs_in, out=None):
        """obs (host) -> policy (tensor cores) -> step (auto-reset) -> obs, reward, done (host)."""
        o = None if obs_in is None else np.ascontiguousarray(obs_in, dtype=np.float32).reshape(self.n, self.O)
        if out is None:
            out = (np.empty((self.n, self.O), dtype=np.float32), np.empty(self.n, dtype=np.float32), np.empty(self.n, dtype=np.uint8))
        obs, rew, done = out
        _lib.check(self.L.rs_world_policy_step_host(self.h, policy.h, _ptr(o), obs.ctypes.data, rew.ctypes.data, done.ctypes.data))
        return obs, rew, done

    def substeps(self, n, tau=None):
        t = None
        if tau is not None:
            t = np.ascontiguousarray(tau, dtype=np.float32).reshape(self.n, self.model.n_dof)
        _lib.check(self.L.rs_world_substeps(self.h, int(n), _ptr(t)))

    def set_joint_motor(self, dof, kp, kd, target_pos, target_vel, max_impulse, env=-1):
        """btMultiBodyJointMotor of one joint dof (env -1: every env); max_impulse per sub-step, <= 0 switches it off."""
        _lib.check(self.L.rs_world_set_joint_motor(self.h, int(env), int(dof), float(kp), float(kd), float(target_pos), float(target_vel), float(max_impulse)))

    def set_state(self, state, clear_contacts=True):
        s = np.ascontiguousarray(state, dtype=np.float32).reshape(self.n, self.S)
        _lib.check(self.L.rs_world_set_state(self.h, s.ctypes.data, int(clear_contacts)))

    def get_state(self):
        s = np.empty((self.n, self.S), dtype=np.float32)
        _lib.check(self.L.rs_world_get_state(self.h, s.ctypes.data))
        return s

    def set_episode(self, potential=None, initial_z=None, feet_contact=None, frame=None):
        p = None if potential is None else np.ascontiguousarray(potential, dtype=np.float64).reshape(self.n)
        z = None if initial_z is None else np.ascontiguousarray(initial_z, dtype=np.float32).reshape(self.n)
        f = None if feet_contact is None else np.ascontiguousarray(feet_contact, dtype=np.float32).reshape(self.n, self.model.n_feet)
        fr = None if frame is None else np.ascontiguousarray(frame, dtype=np.int32).reshape(self.n)
        _lib.check(self.L.rs_world_set_episode(self.h, _ptr(p), _ptr(z), _ptr(f), _ptr(fr)))

    def get_episode(self):
        p = np.empty(self.n, dtype=np.float64)
        z = np.empty(self.n, dtype=np.float32)
        f = np.empty((self.n, max(self.model.n_feet, 1)), dtype=np.float32)
        fr = np.empty(self.n, dtype=np.int32)
        _lib.check(self.L.rs_world_get_episode(self.h, p.ctypes.data, z.ctypes.data, f.ctypes.data, fr.ctypes.data))
        return p, z, f[:, :self.model.n_feet], fr

    def set_flags(self, target_xy=None, timeout=None, draws=None):
        """Flagrun: per-env flag position [n,2] (float64), steps until it moves, RNG draws consumed."""
        t = None if target_xy is None else np.ascontiguousarray(target_xy, dtype=np.float64).reshape(self.n, 2)
        to = None if timeout is None else np.ascontiguousarray(timeout, dtype=np.int32).reshape(self.n)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.uint32).reshape(self.n)
        _lib.check(self.L.rs_world_set_flags(self.h, _ptr(t), _ptr(to), _ptr(d)))

    def get_flags(self):
        t = np.empty((self.n, 2), dtype=np.float64)
        to = np.empty(self.n, dtype=np.int32)
        d = np.empty(self.n, dtype=np.uint32)
        _lib.check(self.L.rs_world_get_flags(self.h, t.ctypes.data, to.ctypes.data, d.ctypes.data))
        return t, to, d

    def set_task_state(self, state):
        """FlagrunHarder task state [n,8] float64 (see rs_world_set_task_state)."""
        s = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, 8)
        _lib.check(self.L.rs_world_set_task_state(self.h, s.ctypes.data))

    def get_task_state(self):
        s = np.empty((self.n, 8), dtype=np.float64)
        _lib.check(self.L.rs_world_get_task_state(self.h, s.ctypes.data))
        return s

    def link_poses(self):
        out = np.empty((self.n, self.model.n_links + 1, 10), dtype=np.float32)
        _lib.check(self.L.rs_world_link_poses(self.h, out.ctypes.data))
        return out

    def contacts(self, env, max_pts=64):
        keys = np.zeros(max_pts, dtype=np.int32)
        data = np.zeros((max_pts, 10), dtype=np.float32)
        cnt = ctypes.c_int()
        _lib.check(self.L.rs_world_get_contacts(self.h, env, keys.ctypes.data, data.ctypes.data, max_pts, ctypes.addressof(cnt)))
        return keys[:cnt.value], data[:cnt.value]

    def set_contacts(self, counts, keys, data):
        """Overwrite every env's persistent contact manifolds: counts [n], keys [n,P], data [n,P,10] (solver order)."""
        c = np.ascontiguousarray(counts, dtype=np.int32).reshape(self.n)
        k = np.ascontiguousarray(keys, dtype=np.int32).reshape(self.n, -1)
        d = np.ascontiguousarray(data, dtype=np.float32).reshape(self.n, k.shape[1], 10)
        _lib.check(self.L.rs_world_set_contacts(self.h, c.ctypes.data, k.ctypes.data, d.ctypes.data, int(k.shape[1])))

    def counts(self):
        r = np.empty(self.n, dtype=np.int32)
        c = np.empty(self.n, dtype=np.int32)
        _lib.check(self.L.rs_world_get_counts(self.h, r.ctypes.data, c.ctypes.data))
        return r, c

    def rewards5(self):
        out = np.empty((self.n, 5), dtype=np.float32)
        _lib.check(self.L.rs_world_get_rewards5(self.h, out.ctypes.data))
        return out

    # ---- fused on-device rollout ----
    def rollout_reset(self, seed=0, env_id0=0, stream=None):
        _lib.check(self.L.rs_world_rollout_reset(self.h, seed, env_id0, stream))

    def rollout(self, policy, n_steps, seed=None, env_id0=None, stats=None, stream=None):
        """n_steps fused policy + env steps.  Auto-resets / random actions use the seed and env_id0 of the last
        reset / rollout_reset (the `seed` / `env_id0` arguments are accepted for source compatibility and ignored)."""
        ph = policy.h if policy is not None else None
        _lib.check(self.L.rs_world_rollout(self.h, ph, int(n_steps), _ptr(stats), stream))

    def rollout_record(self, policy, n_steps, obs_T, act_T, rew_T, done_T, obs0=None, stats=None, stream=None):
        """T steps into time-major device tensors (raw pointers or torch tensors); see rs_world_rollout_record."""
        ph = policy.h if policy is not None else None
        q = lambda t: t if (t is None or isinstance(t, int)) else t.data_ptr()
        _lib.check(self.L.rs_world_rollout_record(self.h, ph, int(n_steps), q(obs0), q(obs_T), q(act_T), q(rew_T), q(done_T), _ptr(stats), stream))

    def rollout_buffers(self):
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        _lib.check(self.L.rs_world_rollout_buffers(self.h, *[ctypes.byref(p) for p in ptrs]))
        return [p.value for p in ptrs]

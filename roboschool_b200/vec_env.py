"""VecEnv: the batched vector-env entry point added on top of the reference's gym.Env contract.

    env = VecEnv("RoboschoolHumanoid-v1", n_envs=32768, device=0)
    obs = env.reset(seed=0)                       # torch.float32 [N, O] on the GPU
    obs, rew, done = env.step(actions)            # actions torch.float32 [N, A] on the GPU

Semantics per env are RoboschoolForwardWalker.reset/step (gym_mujoco_xml_env.py:39-78,
gym_forward_walker.py:89-133); finished envs (done or max_episode_steps=1000,
roboschool/__init__.py:71-77) are reset on the device in the same launch.
PyTorch is only the tensor hand-off: allocation, streams, policies.
"""
import torch

from .envspec import load_env_model
from .world import BatchedWorld


class VecEnv:
    def __init__(self, env_id, n_envs, device=0, env_id0=0, model=None, auto_reset=True):
        if not torch.cuda.is_available():
            raise RuntimeError("roboschool_b200.VecEnv needs a CUDA device; there is no CPU fallback")
        self.env_id = env_id
        self.model = model if model is not None else load_env_model(env_id)
        self.n = int(n_envs)
        self.device = torch.device("cuda", device)
        self.env_id0 = env_id0
        self.auto_reset = auto_reset
        torch.cuda.set_device(self.device)
        self.world = BatchedWorld(self.model, self.n, device)
        self.obs = torch.empty((self.n, self.model.obs_dim), dtype=torch.float32, device=self.device)
        self.rew = torch.empty((self.n,), dtype=torch.float32, device=self.device)
        self.done = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        self.seed_value = 0

    @property
    def obs_dim(self):
        return self.model.obs_dim

    @property
    def act_dim(self):
        return self.model.n_act

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, seed=0, mask=None):
        self.seed_value = int(seed)
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self.world.reset(self.seed_value, self.env_id0, m, self.obs, self._stream())
        return self.obs

    def step(self, actions):
        a = actions
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        assert a.shape == (self.n, self.model.n_act)
        self.world.step(a, self.obs, self.rew, self.done, self.auto_reset, self._stream())
        return self.obs, self.rew, self.done

    def rollout(self, policy, n_steps, actions=None):
        """Collect n_steps transitions on the device for a trainer: returns time-major tensors
        (obs [T+1,N,O], act [T,N,A], rew [T,N], done [T,N] uint8).  `policy` is a ZooPolicy-like object (MLP on the
        tensor cores); with policy=None the given `actions` [T,N,A] are replayed.  Continues from the current state
        (call reset() first); finished envs auto-reset inside the rollout."""
        T, N, O, A = int(n_steps), self.n, self.model.obs_dim, self.model.n_act
        obs_T = torch.empty((T + 1, N, O), dtype=torch.float32, device=self.device)
        rew_T = torch.empty((T, N), dtype=torch.float32, device=self.device)
        done_T = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        if policy is None:
            assert actions is not None and tuple(actions.shape) == (T, N, A)
            act_T = actions.to(device=self.device, dtype=torch.float32).contiguous()
        else:
            act_T = torch.empty((T, N, A), dtype=torch.float32, device=self.device)
        self.world.rollout_record(policy, T, obs_T, act_T, rew_T, done_T, obs0=self.obs, stream=self._stream())
        self.obs.copy_(obs_T[T])
        return obs_T, act_T, rew_T, done_T

    def close(self):
        self.world.close()

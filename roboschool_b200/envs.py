"""Single-env gym.Env facade over the batched stepper (N=1), keeping the reference's contract:

    env = roboschool_b200.make("RoboschoolHumanoid-v1")
    obs = env.reset()                       # np.float32[O]            (gym_mujoco_xml_env.py:39,78)
    obs, reward, done, info = env.step(a)   # float reward, bool done  (gym_forward_walker.py:133)

action_space Box(-1,1,[A]) float32, observation_space Box(-inf,inf,[O]) float32
(gym_mujoco_xml_env.py:26-29); TimeLimit of max_episode_steps=1000 from the registration
(roboschool/__init__.py:11-110) is applied here since there is no gym registry wrapper.
For thousands of envs use roboschool_b200.VecEnv instead; this class exists for drop-in use and tests.
"""
import numpy as np

from . import gym_shim
from .envspec import ENV_SPECS, load_env_model


class RoboschoolB200Env(gym_shim.Env):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 60}

    def __init__(self, env_id, device=0, model=None):
        from .world import BatchedWorld
        self.env_id = env_id
        self.model = model if model is not None else load_env_model(env_id)
        m = self.model
        high = np.ones([m.n_act])
        self.action_space = gym_shim.Box(-high, high, dtype=np.float32)
        high = np.inf * np.ones([m.obs_dim])
        self.observation_space = gym_shim.Box(-high, high, dtype=np.float32)
        self.spec = gym_shim.EnvSpec(env_id, m.max_episode_steps, m.meta.get("reward_threshold"))
        self.world = BatchedWorld(m, 1, device)
        self._episode = 0
        self.frame = 0
        self.done = 0
        self.reward = 0.0
        self.rewards = [0.0] * 5
        self.seed()

    def seed(self, seed=None):
        self.np_random, seed = gym_shim.np_random(seed)
        self._seed = 0 if seed is None else int(seed) & 0xFFFFFFFF
        return [seed]

    def reset(self):
        import torch
        obs = torch.empty((1, self.model.obs_dim), dtype=torch.float32, device="cuda:%d" % self.world.device)
        # one episode = one draw of the counter-based RNG stream (seed, env 0, episode counter)
        self.world.reset(seed=self._seed, env_id0=self._episode, obs=obs)
        self._episode += 1
        self.frame, self.done, self.reward = 0, 0, 0.0
        return obs.cpu().numpy()[0]

    def step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done = self.world.step_host(a.reshape(1, -1), auto_reset=False)
        self.rewards = self.world.rewards5()[0].tolist()
        self.frame += 1
        d = bool(done[0])
        self.done += d
        self.reward += float(rew[0])
        # gym's TimeLimit wrapper from the registration (roboschool/__init__.py: max_episode_steps=1000, Reacher 150):
        # done at the limit, flagged as a truncation the way later gym versions do
        info = {}
        if self.frame >= self.spec.max_episode_steps:
            info["TimeLimit.truncated"] = not d
            d = True
        return obs[0], float(rew[0]), d, info

    def render(self, mode="human"):
        return None   # renderer out of scope

    def close(self):
        self.world.close()


def make(env_id, **kw):
    if env_id not in ENV_SPECS:
        raise KeyError("unknown env id %r (have: %s)" % (env_id, ", ".join(sorted(ENV_SPECS))))
    return RoboschoolB200Env(env_id, **kw)

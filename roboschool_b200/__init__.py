"""roboschool_b200 -- B200-native batched replacement for roboschool's cpp_household World::step path.

    VecEnv(env_id, n_envs)          batched vector env (torch CUDA tensors in/out)
    make(env_id)                    single-env gym.Env facade (N=1)
    cpp_household                   drop-in module: World/Robot/Joint/Thingy/Pose over the C ABI
    ZooPolicy / load_zoo_weights    agent_zoo v1 MLP policies on the tensor cores
"""
from .envspec import ENV_SPECS, load_env_model, compile_env  # noqa: F401


def __getattr__(name):   # lazy: importing the package must not need torch or the CUDA library
    if name == "VecEnv":
        from .vec_env import VecEnv
        return VecEnv
    if name in ("make", "RoboschoolB200Env"):
        from . import envs
        return getattr(envs, name)
    if name in ("ZooPolicy", "load_zoo_weights"):
        from . import policy
        return getattr(policy, name)
    if name == "BatchedWorld":
        from .world import BatchedWorld
        return BatchedWorld
    raise AttributeError(name)

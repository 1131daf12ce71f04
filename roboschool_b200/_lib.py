"""ctypes binding of libroboschool_b200.so (include/roboschool_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is visible,
every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RS_B200_LIB") or os.path.join(_HERE, "libroboschool_b200.so")     # RS_B200_LIB: A/B builds (tools/_ab)
_LIB = None


class RoboschoolB200Error(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RoboschoolB200Error(
            "CUDA extension %s is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). roboschool_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    P, I, U = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
    PP = ctypes.POINTER(ctypes.c_void_p)
    L.rs_last_error.restype = ctypes.c_char_p
    L.rs_device_count.restype = I
    L.rs_world_create.argtypes = [P, I, I, PP]
    L.rs_model_compile.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_double, I, I, P, P]
    L.rs_world_destroy.argtypes = [P]
    for f in ("rs_world_kernel_variant", "rs_world_cta_warps", "rs_world_n_envs", "rs_world_state_dim", "rs_world_obs_dim", "rs_world_act_dim"):
        getattr(L, f).argtypes = [P]
    L.rs_world_reset.argtypes = [P, U, U, P, P, P]
    L.rs_world_step.argtypes = [P, P, P, P, P, I, P]
    L.rs_world_step_stats.argtypes = [P, P, P, P, P, I, P, P]
    L.rs_world_step_profile.argtypes = [P, P, P, P, P, P, P]
    L.rs_world_step_host.argtypes = [P, P, P, P, P, I]
    L.rs_world_policy_step_host.argtypes = [P, P, P, P, P, P]
    L.rs_world_substeps.argtypes = [P, I, P]
    L.rs_world_set_joint_motor.argtypes = [P, I, I, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    L.rs_world_set_state.argtypes = [P, P, I]
    L.rs_world_get_state.argtypes = [P, P]
    L.rs_world_set_episode.argtypes = [P, P, P, P, P]
    L.rs_world_get_episode.argtypes = [P, P, P, P, P]
    L.rs_world_set_flags.argtypes = [P, P, P, P]
    L.rs_world_get_flags.argtypes = [P, P, P, P]
    L.rs_world_set_task_state.argtypes = [P, P]
    L.rs_world_get_task_state.argtypes = [P, P]
    L.rs_world_link_poses.argtypes = [P, P]
    L.rs_world_get_contacts.argtypes = [P, I, P, P, I, P]
    L.rs_world_get_counts.argtypes = [P, P, P]
    L.rs_world_set_contacts.argtypes = [P, P, P, P, I]
    L.rs_world_get_rewards5.argtypes = [P, P]
    L.rs_world_state_dev.argtypes = [P, PP]
    L.rs_policy_create.argtypes = [I, I, I, I, P, P, P, P, P, P, I, PP]
    L.rs_policy_destroy.argtypes = [P]
    L.rs_policy_forward.argtypes = [P, P, P, I, P]
    L.rs_world_rollout.argtypes = [P, P, I, P, P]
    L.rs_world_rollout_reset.argtypes = [P, U, U, P]
    L.rs_world_rollout_buffers.argtypes = [P, PP, PP, PP, PP]
    L.rs_world_rollout_record.argtypes = [P, P, I, P, P, P, P, P, P, P]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise RoboschoolB200Error(lib().rs_last_error().decode("utf-8", "replace"))


EXPORTED_SYMBOLS = [
    "rs_last_error", "rs_device_count", "rs_model_compile", "rs_world_create", "rs_world_destroy", "rs_world_n_envs", "rs_world_kernel_variant", "rs_world_cta_warps",
    "rs_world_state_dim", "rs_world_obs_dim", "rs_world_act_dim", "rs_world_reset", "rs_world_step",
    "rs_world_step_stats", "rs_world_step_profile", "rs_world_step_host", "rs_world_policy_step_host", "rs_world_substeps", "rs_world_set_joint_motor", "rs_world_set_state", "rs_world_get_state",
    "rs_world_set_episode", "rs_world_get_episode", "rs_world_set_flags", "rs_world_get_flags", "rs_world_set_task_state", "rs_world_get_task_state", "rs_world_link_poses", "rs_world_get_contacts", "rs_world_set_contacts",
    "rs_world_get_counts", "rs_world_get_rewards5", "rs_world_state_dev", "rs_policy_create",
    "rs_policy_destroy", "rs_policy_forward", "rs_world_rollout", "rs_world_rollout_reset",
    "rs_world_rollout_buffers", "rs_world_rollout_record",
]

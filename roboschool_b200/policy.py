"""agent_zoo v1 MLP policies on the tensor cores (agent_zoo/*_v1_2017jul.py:21-34,63-68)."""
import ctypes
import os

import numpy as np

from . import _lib

ZOO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zoo")


def load_zoo_weights(env_id, version="v1"):
    """Weights of the reference's pretrained policy for `env_id` (converted by tools/convert_zoo_weights.py).
    version "v1": agent_zoo/*_v1_2017jul.weights; "v0": the independently trained agent_zoo/*_v0_2017may.py."""
    name = env_id.replace("-v1", "")
    path = os.path.join(ZOO_DIR, name + "_" + version + ".npz")
    if version == "v1" and not os.path.exists(path):      # the pendulums and the reacher only have a May-2017 policy
        path = os.path.join(ZOO_DIR, name + "_v0.npz")
    z = np.load(path)
    return [z[k] for k in ("dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b")]


class ZooPolicy:
    """relu(relu(x W1 + b1) W2 + b2) W3 + b3 for a batch of observations held in device memory."""

    def __init__(self, weights, device=0):
        self.L = _lib.lib()
        w = [np.ascontiguousarray(a, dtype=np.float32) for a in weights]
        self.in_dim, self.h1 = w[0].shape
        self.h2 = w[2].shape[1]
        self.out_dim = w[4].shape[1]
        h = ctypes.c_void_p()
        _lib.check(self.L.rs_policy_create(self.in_dim, self.h1, self.h2, self.out_dim,
                                           *[a.ctypes.data for a in w], device, ctypes.byref(h)))
        self.h = h

    def forward(self, obs, act, n=None, stream=None):
        """obs/act: torch CUDA float32 tensors [n,in] / [n,out] (or raw device pointers)."""
        if n is None:
            n = obs.shape[0]
        op = obs if isinstance(obs, int) else obs.data_ptr()
        ap = act if isinstance(act, int) else act.data_ptr()
        _lib.check(self.L.rs_policy_forward(self.h, op, ap, int(n), stream))

    def close(self):
        if getattr(self, "h", None):
            self.L.rs_policy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

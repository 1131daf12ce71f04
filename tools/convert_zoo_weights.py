"""Convert agent_zoo/*_v1_2017jul.weights (python literals exec'd by the reference,
agent_zoo/RoboschoolHumanoid_v1_2017jul.py:36-46) into .npz files the GPU box can load.

Usage: python tools/convert_zoo_weights.py [/root/reference/agent_zoo]
"""
import glob, os, sys
import numpy as np
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/agent_zoo"
dst = os.path.join(os.path.dirname(__file__), "..", "roboschool_b200", "zoo")
os.makedirs(dst, exist_ok=True)
for f in sorted(glob.glob(os.path.join(src, "*_v1_2017jul.weights"))):
    ns = {}
    exec(open(f).read(), ns)
    name = os.path.basename(f).replace("_v1_2017jul.weights", "")
    arrs = {k: np.asarray(ns["weights_" + k], dtype=np.float64) for k in
            ["dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b"]}
    np.savez_compressed(os.path.join(dst, name + "_v1.npz"), **arrs)
    print(name, [a.shape for a in arrs.values()])

# The May-2017 policies keep their weights inline in the agent script (agent_zoo/RoboschoolHopper_v0_2017may.py:4-22:
# same relu-relu-linear network); they were trained independently of the July ones, so they are a second behavioural
# check of the restated physics.  Only the module-level `weights_*` literals are evaluated, not the demo loop.
for f in sorted(glob.glob(os.path.join(src, "*_v0_2017may.py"))):
    text = open(f).read()
    if "weights_dense1_w" not in text:
        continue
    body = text[text.index("\nweights_dense1_w = "):text.index("if __name__")]
    ns = {"np": np}
    exec(body, ns)
    name = os.path.basename(f).replace("_v0_2017may.py", "")
    arrs = {k: np.asarray(ns["weights_" + k], dtype=np.float64) for k in
            ["dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b"]}
    np.savez_compressed(os.path.join(dst, name + "_v0.npz"), **arrs)
    print(name, "v0", [a.shape for a in arrs.values()])

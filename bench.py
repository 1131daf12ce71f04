"""bench.py -- env-steps/sec of the batched Humanoid stepper (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E]

A "step" is one env.step of every env of the batch: agent_zoo v1 MLP policy forward on the tensor
cores + apply_action + 4 Bullet-style sub-steps + calc_state/reward/done + on-device auto-reset.
Workload: RoboschoolHumanoid-v1, 32768 envs per GPU (BASELINE.json configs[3]: 262144 envs over 8
GPUs), weak scaling, synthetic (seeded reset distribution, pretrained zoo weights shipped as npz).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "RoboschoolHumanoid-v1"
METRIC = "env-steps/sec, RoboschoolHumanoid-v1 batched"
UNIT = "env-steps/s"


def algorithmic_bytes(m):
    """state read once + written once, actions in, obs/reward/done out, fp32."""
    S = (0 if m.fixed_base else 13) + 2 * m.n_dof
    return 4 * (2 * S + m.n_act + m.obs_dim + 2)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.p = p
        for line in p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            if self.stop_flag:
                break
        try:
            p.kill()
        except Exception:
            pass

    def summary(self):
        self.stop_flag = True
        try:
            self.p.kill()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(n_envs, n_steps, threads, warmup=1, min_seconds=0.0):
    """The CPU restatement (oracle, 'port') on the host cores: same model, same policy (numpy), auto-reset.
    Runs n_steps env steps, then keeps going until min_seconds of wall time have elapsed."""
    from roboschool_b200.envspec import load_env_model
    from roboschool_b200.policy import load_zoo_weights
    from oracle.oracle import OracleWorld, mlp_forward
    m = load_env_model(ENV_ID)
    wts = load_zoo_weights(ENV_ID)
    o = OracleWorld(m, n_envs)
    o.reset(seed=0)
    obs, _, _ = o.obs()
    for _ in range(warmup):
        a = mlp_forward(wts, obs)
        obs, _, _ = o.step(a, auto_reset=True, seed=0, threads=threads)
    t0 = time.perf_counter()
    done_steps = 0
    while done_steps < n_steps or time.perf_counter() - t0 < min_seconds:
        a = mlp_forward(wts, obs)
        obs, _, _ = o.step(a, auto_reset=True, seed=0, threads=threads)
        done_steps += 1
    dt = time.perf_counter() - t0
    return n_envs * done_steps / dt, dt, done_steps


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = max(cores * 16, 64)
    per_step_value = []
    # bounded: the whole --steps K --warmup W run ends within a few minutes whatever K is
    k_steps = min(args.steps, 40)
    for _ in range(min(args.warmup, 3)):
        cpu_baseline_run(n_envs, 1, cores, warmup=0)
    t_total = 0.0
    for _ in range(k_steps):
        v, dt, _n = cpu_baseline_run(n_envs, 4, cores, warmup=0)
        per_step_value.append(v); t_total += dt
    value = float(np.mean(per_step_value))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(k_steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "RoboschoolHumanoid-v1 CPU stepper, %d envs x 4 env-steps per bench step, zoo v1 MLP in numpy, auto-reset" % n_envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d envs x 4 env-steps per step, %d threads (CPU restatement of the Bullet path; the reference's Bullet fork cannot be built here)" % (n_envs, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--envs-per-gpu", type=int, default=32768)
    ap.add_argument("--env", default=ENV_ID)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from roboschool_b200.envspec import load_env_model
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    from roboschool_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; roboschool_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = load_env_model(args.env)
    n = args.envs_per_gpu
    env_id0 = rank * n
    W = BatchedWorld(m, n, local_rank)
    pol = ZooPolicy(load_zoo_weights(args.env), local_rank)
    obs_p, act_p, rew_p, done_p = W.rollout_buffers()
    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    W.rollout_reset(seed=0, env_id0=env_id0, stream=stream)

    def one_step():
        W.rollout(pol, 1, seed=0, env_id0=env_id0, stats=stats, stream=stream)

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    stats.zero_()
    # --- device-resident timed region: K steps, events around the whole region and around each k_step launch ---
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0.record()
    for t in range(args.steps):
        pev[t].record()
        pol.forward(obs_p, act_p, n, stream)
        kev[t][0].record()
        _lib.check(W.L.rs_world_step_stats(W.h, act_p, obs_p, rew_p, done_p, 1, stats.data_ptr(), stream))
        kev[t][1].record()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.summary() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ms_policy = float(np.mean([p.elapsed_time(k[0]) for p, k in zip(pev, kev)]))
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.barrier()
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total_max = float(t_ms.item())
    total_envs = n * world_size
    value = total_envs * args.steps / (ms_total_max * 1e-3)

    # --- rollout statistics: the only collective of the path (NCCL all-reduce of 8 doubles) ---
    if world_size > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)

    # --- end-to-end through the host-buffer C-ABI call: every step copies the observations H2D from (pinned-staged)
    # host memory, runs policy + step, and copies obs/reward/done back D2H ---
    e2e_steps = max(10, args.steps // 4)
    torch.cuda.synchronize()
    bufs = [W.pinned_host_buffers(), W.pinned_host_buffers()]   # page-locked ping-pong: this step's output is the next step's input
    o_np, r_np, d_np = W.policy_step_host(pol, None, out=bufs[0])   # continue from the device-resident rollout: same steady-state workload
    for k in range(3):
        o_np, r_np, d_np = W.policy_step_host(pol, o_np, out=bufs[(k + 1) & 1])
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        o_np, r_np, d_np = W.policy_step_host(pol, o_np, out=bufs[k & 1])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = total_envs * e2e_steps / float(t_e.item())
    h2d = n * m.obs_dim * 4
    d2h = n * (m.obs_dim * 4 + 4 + 1)

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    abytes = algorithmic_bytes(m)
    achieved = abytes * n / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    if os.path.exists(tp) and args.env == ENV_ID and n == 32768:     # the capture is of the default workload
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = max(cores * 16, 64)
        v, dt, ns = cpu_baseline_run(n_cpu, 8, cores, warmup=2, min_seconds=12.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d envs x %d env-steps of the same workload, %d threads, %.1f s (CPU restatement of the Bullet path, not Bullet itself)" % (n_cpu, ns, cores, dt)}

    # the policy forward is the one contraction of the path: reported against the tensor peak
    wshape = [w.shape for w in load_zoo_weights(args.env)]
    flop_env = 2 * (wshape[0][0] * wshape[0][1] + wshape[2][0] * wshape[2][1] + wshape[4][0] * wshape[4][1])
    tpeak = float(peaks.get("bf16_tflops", 1660.3))        # the kernel is timed alone between two events: burst figure
    tachieved = flop_env * n / (ms_policy * 1e-3) / 1e12
    # launches: policy + step per timed step; the host entry point cuts the envs into stream slices (2 launches per slice)
    chunks = 1
    if W.L.rs_world_kernel_variant(W.h) > 0 and n >= 1024:
        chunks = max(1, min(8, int(os.environ.get("RS_E2E_CHUNKS", "6" if n >= 24576 else "1"))))
    st = stats.cpu().numpy()
    line = {
        "metric": METRIC.replace(ENV_ID, args.env), "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s, %d envs/GPU (%d total), agent_zoo v1 MLP policy on tensor cores, frame_skip 4, on-device auto-reset" % (args.env, n, total_envs),
                   "l2": "per-step working set (state+contacts+row workspace, >400 MB/GPU) exceeds the 126 MB L2",
                   "parallelism": "env-sharded, %d process(es), no step-path collective" % world_size},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_step", "algorithmic_bytes_per_env_step": abytes, "kernel_ms": ms_kernel, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "rs_world_policy_step_host (pinned host obs in -> policy -> step -> pinned host obs/reward/done out), %d steps, wall clock" % e2e_steps},
        "roofline_policy": {"bound": "tensor", "achieved": tachieved, "peak": tpeak, "unit": "TFLOP/s", "frac": tachieved / tpeak,
                            "kernel": "k_mlp_tc5", "flop_per_env": flop_env, "kernel_ms": ms_policy,
                            "peak_source": "measured (MEASURED_PEAKS.json, dense bf16 burst)" if "bf16_tflops" in peaks else "fallback"},
        "gpu_launches": 2 * args.steps + 2 * chunks * (e2e_steps + 4),
        "clocks": clocks,
        "rollout_stats": {"mean_reward": st[0] / max(st[6], 1), "episodes": st[1], "mean_episode_len": st[2] / max(st[1], 1),
                          "nonfinite": st[3], "mean_rows": st[4] / max(st[6], 1), "mean_contacts": st[5] / max(st[6], 1), "env_steps": st[6]},
    }
    print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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


_CLOCK_Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def clock_query(index):
    """One synchronous nvidia-smi query (the line's `clocks` block is never empty, whatever the length of the timed region)."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=" + _CLOCK_Q, "--format=csv,noheader,nounits"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip().splitlines()
        return [x.strip() for x in out[0].split(",")] if out else None
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi sampler (50 ms period) running from the warm-up steps to the end of the end-to-end region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []      # (wall time, fields)
        self.stop_flag = False
        self.p = None

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + _CLOCK_Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.p = p
        for line in p.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
            if self.stop_flag:
                break
        try:
            p.kill()
        except Exception:
            pass

    def summary(self, t0=None, t1=None, extra=None):
        """Median SM clock / throttle reasons over the samples taken inside [t0, t1] (the timed regions), plus `extra`
        (a synchronous query made while the timed steps were in flight)."""
        self.stop_flag = True
        try:
            self.p.kill()
        except Exception:
            pass
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        if extra:
            rows.append(extra)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": "device-timed + end-to-end regions"}


class CpuStepper:
    """The CPU restatement (oracle, 'port') on the host cores: ONE world kept at steady state (persistent worker threads,
    >= 64 envs per thread, zoo policy in numpy, auto-reset) -- the world is settled once and then only stepped."""

    def __init__(self, env_id, threads, envs_per_thread=64, settle_steps=60, settle_seconds=3.0):
        from roboschool_b200.envspec import load_env_model
        from roboschool_b200.policy import load_zoo_weights
        from oracle.oracle import OracleWorld
        try:    # BLAS worker threads spin after every matmul and steal the cores from the stepper's own thread pool (measured: 2.4x)
            from threadpoolctl import threadpool_limits
            self._blas = threadpool_limits(limits=1)
        except Exception:
            self._blas = None
        self.m = load_env_model(env_id)
        self.wts = [np.asarray(w, dtype=np.float32) for w in load_zoo_weights(env_id)]   # fp32 like the reference's TF policy
        self.threads = threads
        self.n = max(64, threads * envs_per_thread)
        self.o = OracleWorld(self.m, self.n)
        self.o.reset(seed=0)
        self.obs = self.o.obs()[0]
        # settle: episodes under way, contacts in steady state, worker threads (and the VM's vCPUs) spun up
        t0 = time.perf_counter()
        k = 0
        while k < settle_steps or time.perf_counter() - t0 < settle_seconds:
            self.step(); k += 1

    def step(self):
        w = self.wts
        x = np.maximum(self.obs.astype(np.float32) @ w[0] + w[1], 0)
        x = np.maximum(x @ w[2] + w[3], 0)
        a = (x @ w[4] + w[5]).astype(np.float64)
        self.obs, _, _ = self.o.step(a, auto_reset=True, seed=0, threads=self.threads)

    def measure(self, seconds, min_steps=4):
        t0 = time.perf_counter()
        k = 0
        while k < min_steps or time.perf_counter() - t0 < seconds:
            self.step(); k += 1
        dt = time.perf_counter() - t0
        return self.n * k / dt, dt, k


def cpu_baseline(env_id, cores, seconds_all=12.0, seconds_one=5.0):
    """All-cores figure (median of 3 equal slices of one settled world) + the single-core figure."""
    allc = CpuStepper(env_id, cores)
    runs = [allc.measure(seconds_all / 3.0) for _ in range(3)]
    v = float(np.median([r[0] for r in runs]))
    one = CpuStepper(env_id, 1, settle_steps=40, settle_seconds=1.5)
    v1, dt1, k1 = one.measure(seconds_one)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "runs": [r[0] for r in runs], "single_core": {"value": v1, "cores": 1, "sample": "%d envs x %d env-steps, %.1f s" % (one.n, k1, dt1)},
            "sample": "%d envs (64 per thread) x %d env-steps of the same workload on one settled world, %d persistent threads, median of 3 x %.1f s "
                      "(CPU restatement of the Bullet path, not Bullet itself)" % (allc.n, sum(r[2] for r in runs), cores, seconds_all / 3.0)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # world sized so that the K timed steps cover ~10 s of CPU work at ~17 k env-steps/s per thread (a 20-step run over 64 envs
    # per thread lasted 70 ms and read anything between 144 k and 276 k env-steps/s on the same kind of box)
    ept = int(min(1024, max(64, -(-12 * 17000 // max(args.steps, 1)))))
    cpu = CpuStepper(args.env, cores, envs_per_thread=ept, settle_steps=30, settle_seconds=3.0)
    for _ in range(min(max(args.warmup, 3), 50)):
        cpu.step()
    # a bench step = one env.step of every env of the settled CPU world; bounded so that the run ends within ~1.5 minutes
    budget_s = 60.0
    t0 = time.perf_counter()
    k = 0
    while k < args.steps and (k < 4 or time.perf_counter() - t0 < budget_s):
        cpu.step(); k += 1
    dt = time.perf_counter() - t0
    value = cpu.n * k / dt
    one = CpuStepper(args.env, 1, settle_steps=40, settle_seconds=1.5)
    v1, dt1, k1 = one.measure(5.0)
    line = {
        "impl": "reference", "metric": METRIC.replace(ENV_ID, args.env), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(k, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s CPU stepper, one settled world of %d envs, %d of the %d bench steps timed (%.0f s budget), zoo v1 MLP in numpy, auto-reset"
                               % (args.env, cpu.n, k, args.steps, budget_s)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "single_core": {"value": v1, "cores": 1, "sample": "%d envs x %d env-steps, %.1f s" % (one.n, k1, dt1)},
                         "sample": "%d envs x %d env-steps, %d persistent threads, steady state (CPU restatement of the Bullet path; the reference's Bullet fork cannot be built here)" % (cpu.n, k, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def short_run(env_id, n, device, stream, settle=200, steps=50, random_actions=False):
    """Device-resident env-steps/s of another BASELINE config: settle, stagger the episode clocks, time `steps` steps.
    random_actions: i.i.d. U(-1, 1) actions from the counter-based RNG on the device instead of the zoo policy (the north star's
    'synthetic random-action rollouts': the robots fall within ~20 steps, so this is a reset-heavy workload)."""
    import torch
    from roboschool_b200.envspec import load_env_model
    from roboschool_b200.world import BatchedWorld
    from roboschool_b200.policy import ZooPolicy, load_zoo_weights
    m = load_env_model(env_id)
    W = BatchedWorld(m, n, device)
    pol = None if random_actions else ZooPolicy(load_zoo_weights(env_id), device)
    stats = torch.zeros(8, dtype=torch.float64, device="cuda:%d" % device)
    W.rollout_reset(seed=0, env_id0=0, stream=stream)
    W.rollout(pol, settle, stats=stats, stream=stream)
    torch.cuda.synchronize()
    W.set_episode(frame=np.random.RandomState(7).randint(0, max(m.max_episode_steps - 1, 1), size=n).astype(np.int32))
    W.rollout(pol, m.max_episode_steps, stats=stats, stream=stream)      # every env reset once, at its own time (decorrelated gaits)
    torch.cuda.synchronize()
    stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    W.rollout(pol, steps, stats=stats, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = stats.cpu().numpy()
    W.close()
    if pol is not None:
        pol.close()
    return {"env": env_id, "envs": n, "actions": "uniform random (device RNG)" if random_actions else "agent_zoo policy", "value": n * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "settle_steps": settle + m.max_episode_steps, "mean_rows": st[4] / max(st[6], 1), "mean_contacts": st[5] / max(st[6], 1),
            "algorithmic_bytes_per_env_step": algorithmic_bytes(m)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--envs-per-gpu", type=int, default=32768)
    ap.add_argument("--env", default=ENV_ID)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short sub-results of the other BASELINE configs")
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
        W.rollout(pol, 1, stats=stats, stream=stream)

    # --- set-up, untimed: settle the batch into its steady state whatever --warmup is.  SETTLE policy steps after the
    # synchronised reset (contacts and gait established), then the episode clocks are staggered uniformly over the time
    # limit so that time-limit resets arrive at their steady rate (n / max_episode_steps per step), not all at once ---
    SETTLE = 300
    W.rollout(pol, SETTLE, stats=stats, stream=stream)
    torch.cuda.synchronize()
    _p, _z, _f, frame = W.get_episode()
    rng = np.random.RandomState(1234 + rank)
    if not os.environ.get("RS_BENCH_NO_STAGGER"):       # diagnosis only: synchronised episodes (no resets inside the timed region)
        W.set_episode(frame=rng.randint(0, max(m.max_episode_steps - 1, 1), size=n).astype(np.int32))
        # ... and one more time limit of untimed steps, so that every env has gone through a reset at its own time: the envs of
        # a synchronously reset batch walk in step (the step time swings 0.75 <-> 0.97 ms with the gait, period ~25 steps) and a
        # short timed window would sample one phase of that swing; after this the batch is decorrelated (0.84-0.88 ms)
        W.rollout(pol, m.max_episode_steps, stats=stats, stream=stream)
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    stats.zero_()
    # --- device-resident timed region: K steps, events around the whole region and around each k_step launch ---
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_region0 = time.time()
    ev0.record()
    for t in range(args.steps):
        pev[t].record()
        pol.forward(obs_p, act_p, n, stream)
        kev[t][0].record()
        _lib.check(W.L.rs_world_step_stats(W.h, act_p, obs_p, rew_p, done_p, 1, stats.data_ptr(), stream))
        kev[t][1].record()
    ev1.record()
    clock_now = clock_query(local_rank) if rank == 0 else None     # one query while the timed steps are still in flight
    torch.cuda.synchronize()
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
    e2e_steps = max(200, args.steps)          # >= 200 steps (>= ~0.2 s of wall clock at the default workload)
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
    t_region1 = time.time()
    clocks = sampler.summary(t_region0, t_region1, clock_now) if rank == 0 else None
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
        cpu = cpu_baseline(args.env, cores)

    # short device-timed sub-results for the other BASELINE.json configs (2: Hopper 4096, 3: Ant 65536, 5: FlagrunHarder)
    extra = None
    if world_size == 1 and not args.no_extra and args.env == ENV_ID:
        extra = {}
        for cfg, eid, ne in (("configs[1]", "RoboschoolHopper-v1", 4096), ("configs[2]", "RoboschoolAnt-v1", 65536),
                             ("configs[4]", "RoboschoolHumanoidFlagrunHarder-v1", 32768)):
            extra[cfg] = short_run(eid, ne, local_rank, stream)
        # the headline workload under uniform random actions instead of the policy (north star: random-action rollouts)
        extra["random_actions"] = short_run(ENV_ID, n, local_rank, stream, random_actions=True)

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
    # what one step touches per env: the per-thread local frame of the step kernel (world transforms, contact caches, row scalars;
    # ptxas: 2.6-5.1 KB), state in + out, obs / actions, live constraint rows (~47 words each)
    ws_mb = n * (5056 + abytes + 47 * 4 * 4) / 1e6
    if traffic:
        l2_note = "per-step working set ~%.0f MB/GPU (measured DRAM traffic of the step kernel: %.0f MB per launch) exceeds the 126 MB L2; no flush needed" % (ws_mb, traffic / 1e6)
    elif ws_mb > 126:
        l2_note = "per-step working set ~%.0f MB/GPU exceeds the 126 MB L2; no flush needed" % ws_mb
    else:
        l2_note = ("per-step working set ~%.0f MB/GPU fits the 126 MB L2 and is NOT flushed between steps: the world's own state staying "
                   "resident from one step to the next is this configuration's real steady state" % ws_mb)
    line = {
        "metric": METRIC.replace(ENV_ID, args.env), "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "dtype_policy": "f16xf16->f32 (tcgen05 kind::f16, fp32 accumulate; tolerance 5e-3 of the action scale, tests/test_gpu_parity.py)",
        "data": "synthetic",
        "config": {"workload": "%s, %d envs/GPU (%d total), agent_zoo v1 MLP policy on tensor cores, frame_skip 4, on-device auto-reset; steady state: "
                               "%d untimed settle steps after reset + episode clocks staggered over the time limit + %d untimed steps (every env reset once, at its own time), then --warmup" % (args.env, n, total_envs, SETTLE, m.max_episode_steps),
                   "l2": l2_note,
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
        "extra": extra,
        "rollout_stats": {"mean_reward": st[0] / max(st[6], 1), "episodes": st[1], "mean_episode_len": st[2] / max(st[1], 1),
                          "nonfinite": st[3], "mean_rows": st[4] / max(st[6], 1), "mean_contacts": st[5] / max(st[6], 1), "env_steps": st[6]},
    }
    print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

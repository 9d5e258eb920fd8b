#!/usr/bin/env python
"""bench.py — env-turns/sec of the batched Everglades turn step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one game turn for every match of the batch: the random_actions agent kernel for both
players + the turn-step kernel (--agents fused: one launch of evg_step_agents does both).  Workload (config.workload): BASELINE.json configs[4] at
N = 1 — DemoMap, both players random_actions, 1,048,576 lock-step matches per GPU with in-place
auto-reset; matches shard across ranks with NO collective on the step path (weak scaling; the
only exchange is an end-of-run all_gather of episode statistics).

Printed JSON (rank 0, one line): value = device-timed whole-job env-turns/s with state and inputs
resident in HBM; e2e = the same metric through the public host-buffer API (pinned host actions
H2D + step + D2H of observations/rewards/done flags every step); roofline = the step kernel's
algorithmic bytes / its CUDA-event duration against the measured HBM peak; cpu_baseline = the CPU
oracle port (oracle/evg_oracle.c) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_turns_per_sec"
UNIT = "env-turns/s"
# Algorithmic bytes per match-turn (SURVEY.md §8d / DESIGN.md §5): 1331 B fixed (state 227 B read +
# written, actions 28 B, observations 840 B, rewards 8 B, done 1 B) + 16 B per unit slot of the
# groups that fought that turn (mean 21.6 slots for random-vs-random) = 1676 B.
B_ALG_FIXED = 1331
B_ALG_HEALTH_RANDOM = 345
B_ALG = B_ALG_FIXED + B_ALG_HEALTH_RANDOM


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=450)   # 3 full 150-turn episodes
    ap.add_argument("--warmup", type=int, default=150)  # 1 episode
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--agents", default="kernel", choices=["kernel", "fused"],
                    help="random_actions rows from the agent kernel (2 launches/step) or generated inside the step kernel (1 launch)")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "DemoMap random_actions self-match, %d lock-step matches per GPU, auto-reset "
                    "(BASELINE.json configs[4] at N=1; weak-scaled over ranks)" % args.envs_per_gpu,
        "envs_per_gpu": args.envs_per_gpu, "total_envs": args.envs_per_gpu * world,
        "map": "DemoMap.json", "agents": "random_actions vs random_actions (on-device, Philox tape; %s)" % args.agents,
        "turn_limit": 150, "auto_reset": "terminal-obs", "seed": args.seed,
        "l2": "resident state %.2f GB/GPU >> 126 MB L2; no flush between steps" % (args.envs_per_gpu * (256 + 1600) / 1e9),
        "parallelism": "match-sharded x%d, no step-path collective" % world,
    }


# ------------------------------------------------------------------------------------------------
# CPU oracle port on the host cores (cpu_baseline leg and --impl reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(seconds, seed=0, threads=None, turns=150):
    """Time oracle/evg_oracle.c (evo_run_random: both players random_actions, in-place reset) on
    `threads` host threads for about `seconds`.  Returns (env_turns_per_s, threads, sample, elapsed)."""
    import evgsim
    from oracle import evg_oracle as eo

    cfg = evgsim.load_config()
    threads = threads or os.cpu_count() or 1
    eo.lib()
    t0 = time.perf_counter()
    n, _, _ = eo.run_random(cfg, seed, 0, 64, turns)  # calibration, single thread
    per_thread_rate = n / (time.perf_counter() - t0)
    matches = max(16, int(per_thread_rate * seconds / turns))
    results = [0] * threads

    def work(k):
        results[k] = eo.run_random(cfg, seed, 1_000_000 + k * matches, matches, turns)[0]

    ths = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    el = time.perf_counter() - t0
    total = sum(results)
    sample = "%d threads x %d matches x %d turns of the same workload (oracle/evg_oracle.c, gcc -O2)" % (threads, matches, turns)
    return total / el, threads, sample, el


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's algorithm for this path on the host CPU.  The reference is
    pure Python and cannot travel to the GPU box, so this times its C restatement (the oracle port)
    on all host threads; the unmodified Python server measured in the build container is quoted in
    DESIGN.md §6."""
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    budget = min(60.0, max(5.0, 0.05 * (steps + warm)))
    rate, threads, sample, el = cpu_oracle_rate(budget, seed=args.seed)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32/f64", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "elapsed_s": el,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="evg_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(envs_per_gpu):
    """DRAM bytes per step-kernel launch from the committed ncu --set full capture, if one exists for
    this batch size (profiles/step_kernel_traffic.json), else None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))
        per_env = float(d["dram_bytes_per_env_turn"])
        return per_env * envs_per_gpu
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def _emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner,
    build messages) was diverted to stderr by _quiet_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse_args()
    _quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import evgsim
    from evgsim import dist as evd

    assert torch.cuda.is_available(), "bench.py measures the CUDA path; no GPU visible"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    E = args.envs_per_gpu
    first, _ = rank * E, E
    env = evgsim.BatchedEvergladesEnv(E, device=dev, seed=args.seed, auto_reset=evgsim._capi.AUTORESET_TERMINAL,
                                      env_id_offset=first)
    env.reset()
    stream = torch.cuda.current_stream(dev)

    fused = args.agents == "fused"

    def one_step():
        if fused:
            env.step_agents()                   # ONE launch: both players' random_actions rows + the whole turn
        else:
            env.step(env.random_actions())      # agent kernel, then the turn-step kernel

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize(dev)

    # ---- device-timed region: EXACTLY K steps, barrier + synchronize on both sides
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = env.launch_count
    clocks.start()
    t_begin.record(stream)
    for k in range(K):
        a = None if fused else env.random_actions()
        ev[k][0].record(stream)
        if fused:
            env.step_agents()
        else:
            env.step(a)
        ev[k][1].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    launches = env.launch_count - launches0
    step_kernel_name = {0: "evg_step_kernel (warp per match)", 1: "evg_step_tpm_kernel (thread per match)",
                        2: "evg_step_pair_kernel (lane pair per match)"}.get(env._lib.evg_step_kernel_kind(env._h), "?")
    total_ms = t_begin.elapsed_time(t_end)
    step_kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / K

    # ---- end-to-end through the public host-buffer API (pinned host memory, copies inside)
    Ke = max(1, min(args.e2e_steps, K))
    hb = env.host_buffers()
    ring = []
    for _ in range(4):  # pre-drawn host-side action rows (random_actions ignores observations)
        ring.append(env.random_actions().cpu().pin_memory())
        env.step(env._actions)
    for _ in range(3):
        env.step_host(ring[0])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sink = 0.0
    e0.record(stream)
    for k in range(Ke):
        obs_h, rew_h, done_h, _ = env.step_host(ring[k % 4], sync=True)
        sink += float(rew_h[0, 0])  # the host reads the step's result
    e1.record(stream)
    torch.cuda.synchronize(dev)
    e2e_ms = e0.elapsed_time(e1)

    # ---- max over ranks
    t = torch.tensor([total_ms, step_kernel_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, step_kernel_ms, e2e_ms = (float(x) for x in t.tolist())
    stats = evd.gather_episode_stats(env.episode_stats(), device=dev)  # the only collective; not timed

    if rank == 0:
        total_envs = E * world
        value = total_envs * K / (total_ms / 1e3)
        e2e_value = total_envs * Ke / (e2e_ms / 1e3)
        peak, peak_src = measured_hbm_peak()
        achieved = E * B_ALG / (step_kernel_ms / 1e3) / 1e9  # GB/s per GPU, step kernel alone
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32/f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step() * world,
                    "d2h_bytes_per_step": env.d2h_bytes_per_step() * world, "steps": Ke, "ms_per_step": e2e_ms / Ke,
                    "api": "BatchedEvergladesEnv.step_host (evg_step_host), pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": step_kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(E), "peak_source": peak_src,
                         "bytes_per_env_turn": B_ALG, "kernel_ms": step_kernel_ms,
                         "kernel_env_turns_per_s": E / (step_kernel_ms / 1e3)},
            "episode_stats": stats,
        }
        if not args.no_cpu_baseline and world == 1:
            rate, threads, sample, _ = cpu_oracle_rate(args.cpu_seconds, seed=args.seed)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — path-months/s of the Monte Carlo path engine on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 10 --warmup 3           # this repo's CUDA engine
    python bench.py --impl reference --steps 2 --warmup 1     # the CPU arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] / SURVEY §8d config #3 — synthetic
1,000,000 paths x 720 months (working_months=240, 40 retirement years), config.json values,
equity-inflation correlation -0.5, realized-gains tax on both assets, rebalancing on, 2 income
streams; PER GPU (weak scaling: rank r owns global paths [r*1e6, (r+1)*1e6) of one Philox
stream). One "step" = the whole hot path for one batch: the fused timeline kernel (native
Philox draws, fp64, yearly nominal/real/withdrawal-rate series written to HBM) + every device
aggregation the callers need (success count, median first-year withdrawal rate, medians,
final-balance quantiles and histograms, 7-quantile nominal/real bands, 5-quantile WR bands).

`value`   = nominal path-months (N * 720 * n_gpus) / device-timed step, nothing leaves HBM.
`e2e`     = the same metric through the reference-facing call
            RetirementMonteCarloSimulator.run_monte_carlo_simulations(240, 1_000_000) returning
            the reference's 7-tuple on the HOST (summary_df + band frames), wall-clocked.
`roofline`= the timeline kernel against the unit that binds it. Measured (ncu + the issue-model
            microbenchmarks in tools/microbench, profiles/r02_issue_model*.log): not HBM (2 %), not
            the FP64 pipe (~50 % busy) but warp-instruction ISSUE — one instruction per clock per
            SM sub-partition, of which the integer / FP32 work of the draws and selects (half-rate
            on B200) takes most. `frac` = executed thread-instructions per second / (SMs x 4 x 32
            x clock); the instruction count per path-month comes from an ncu capture of THIS build
            (profiles/timeline_counts.json, keyed by a hash of the kernel sources — stale counts
            are refused), the duration is CUDA events of this run. The FP64-pipe view, the
            reference-order census W = 220 and the HBM side are reported next to it.
`search`  = BASELINE.json configs[3]: find_minimum_working_months at 1e6 paths per candidate
            (wall time; the full 601-candidate grid in one launch sequence; the selected month
            checked against the reference's decision procedure on the same table).
`c5`      = BASELINE.json configs[4] (8 GPUs only): 1e9 paths with bands + histograms, wall time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_SLOTS_PER_PATH_MONTH = 220.0  # 108 simple + 8.33 div*8 + 3 exp*15
PATHS_PER_GPU = 1_000_000
WORKING_MONTHS = 240
METRIC = "path_months_per_sec"
UNIT = "path-months/s"
WORKLOAD = ("synthetic C3: 1e6 paths/GPU x 720 months (wm=240, R=40y), rho=-0.5, realized-gains tax on both "
            "assets, rebalancing, 2 income streams; fp64; native Philox")


def _scenario():
    import scenarios

    return dict(scenarios.SYNTH_C3)


def _silence_logs():
    try:
        from loguru import logger

        logger.remove()
    except Exception:
        pass


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a thread every 10 ms
    (a timed region of ten 9-ms steps is shorter than one `nvidia-smi` start-up on an 8-GPU box);
    `nvidia-smi -lms` is the fallback when the NVML binding is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.handle = None
        self.samples = []
        self.bits = 0
        self.stop_flag = threading.Event()
        self.thread = None
        # polling period (measured: 10 / 25 / 100 ms make no difference to the step time, profiles/README.md)
        self.period_s = float(os.environ.get("MCR_BENCH_SAMPLE_MS", "10")) / 1e3

    def _nvml_handle(self):
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber devices: go through the UUID
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, handle

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.bits |= int(reasons(self.handle))
            except Exception:
                pass
            self.stop_flag.wait(self.period_s)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            names = sorted(name for bit, name in self.REASON_BITS.items() if self.bits & bit)
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(self.samples), "reasons": names, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the reference algorithm + the reference's own numpy
# draws) on all host threads. /root/reference is pure Python and cannot travel to the GPU box,
# so this is kind="port"; the Python reference itself measured 1.09e5 path-months/s/core in the
# build container.
# -------------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_worker_init(cfg):
    from oracle import oracle as orc

    _WORKER["p"] = orc.params_from_config(cfg)
    _WORKER["orc"] = orc


def _cpu_worker(part):
    orc, p = _WORKER["orc"], _WORKER["p"]
    shocks = orc.shocks_for_seeds(p, WORKING_MONTHS, part)  # default_rng(seed).standard_normal((n,3)) per path
    recs, traj, real, wr = orc.run_batch(p, WORKING_MONTHS, shocks, 1, want_series=True)
    return int(recs["success"].sum())


def cpu_port_throughput(n_paths: int, procs: int):
    """Reference algorithm on the host cores: SeedSequence path seeds (parent, as in the
    reference), then one process per core drawing the numpy shocks and stepping the C port of
    the path engine over its share of the paths (the reference fans out with
    multiprocessing.Pool the same way, simulation.py:996-1001)."""
    import multiprocessing as mp

    from oracle import oracle as orc

    cfg = _scenario()
    orc.build()
    sim = orc.OracleSimulator(cfg, n_threads=1)
    sim.use_final_seeds()
    n_rows = WORKING_MONTHS + 12 * cfg["retirement_years"]
    pool = mp.get_context("fork").Pool(procs, initializer=_cpu_worker_init, initargs=(cfg,))
    try:
        pool.map(_cpu_worker, [[1, 2]] * procs)  # workers up (library loaded) before the clock starts
        t0 = time.perf_counter()
        seeds = sim.seeds.path_seeds(n_paths)
        chunk = max(32, n_paths // (procs * 8))
        ok = sum(pool.map(_cpu_worker, [seeds[i:i + chunk] for i in range(0, n_paths, chunk)]))
        dt = time.perf_counter() - t0
    finally:
        pool.terminate()
    return n_paths * n_rows / dt, dt, ok / n_paths


def _reference_worker(q, cfg_dict, wm, n, procs):
    """Child process: the UNMODIFIED reference (oracle/_ref/backend) on the host cores."""
    from oracle import ref as oracle_ref

    sys.path.insert(0, oracle_ref.backend_path())
    try:
        from loguru import logger

        logger.remove()
    except Exception:
        pass
    import config as ref_config
    import simulation as ref_simulation

    cfg = ref_config.Config(**dict(cfg_dict, num_processes=procs))
    sim = ref_simulation.RetirementMonteCarloSimulator(cfg)
    sim.use_final_seeds()
    t0 = time.perf_counter()
    out = sim.run_monte_carlo_simulations(working_months=wm, num_simulations=n)
    dt = time.perf_counter() - t0
    q.put((dt, float(sim._success_probability(out[0]))))


def reference_throughput(n_paths: int, procs: int):
    """RetirementMonteCarloSimulator.run_monte_carlo_simulations of the real reference
    (backend/simulation.py:952-1128, its own multiprocessing.Pool, :996-1001), in a fresh process
    so that neither its flat module names nor its Pool touch this one. None when oracle/_ref is
    not staged."""
    import multiprocessing as mp

    from oracle import ref as oracle_ref

    if not oracle_ref.build():
        return None
    cfg = _scenario()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_reference_worker, args=(q, cfg, WORKING_MONTHS, n_paths, procs))
    p.start()
    dt, p_ok = q.get(timeout=1800)
    p.join()
    n_rows = WORKING_MONTHS + 12 * cfg["retirement_years"]
    return n_paths * n_rows / dt, dt, p_ok


REF_SAMPLE = 4000      # paths per step of the real (CPython) reference: ~1e5 path-months/s/core
PORT_SAMPLE = 100_000  # paths per step of the C port


def cpu_baseline_block(threads: int):
    """cpu_baseline of the bench line: the real reference when oracle/_ref is staged (kind
    "reference"), with the oracle port (same algorithm in C, the reference's numpy draws) beside
    it; the port alone otherwise."""
    port_v, port_dt, p_ok = cpu_port_throughput(PORT_SAMPLE, threads)
    port = {"value": port_v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{PORT_SAMPLE} paths x 720 months of the same workload in {port_dt:.1f} s wall (numpy "
                      f"SeedSequence/PCG64 draws + C port of the path engine, {threads} processes)",
            "success_probability": p_ok * 100.0}
    ref = reference_throughput(REF_SAMPLE, threads)
    if ref is None:
        return port
    v, dt, p_ref = ref
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{REF_SAMPLE} paths x 720 months of the same workload in {dt:.1f} s wall: the unmodified "
                      f"backend/simulation.py run_monte_carlo_simulations (oracle/_ref) with num_processes={threads}",
            "success_probability": p_ref, "port": port}


CPU_SEARCH_N = 1000  # paths per probe of the CPU arm's sequential search (the GPU arm uses 1e6)


def _reference_search_worker(q, cfg_dict, procs):
    from oracle import ref as oracle_ref

    sys.path.insert(0, oracle_ref.backend_path())
    try:
        from loguru import logger

        logger.remove()
    except Exception:
        pass
    import config as ref_config
    import simulation as ref_simulation

    sim = ref_simulation.RetirementMonteCarloSimulator(ref_config.Config(**dict(cfg_dict, num_processes=procs)))
    t0 = time.perf_counter()
    months, prob, curve = sim.find_minimum_working_months(verbose=False)
    q.put((time.perf_counter() - t0, months, prob, len(curve)))


def cpu_search_block(threads: int):
    """BASELINE.json configs[3] on the host cores: the reference's SEQUENTIAL search
    (backend/simulation.py:1138-1342, one run_monte_carlo_simulations per probe) on config.json at
    CPU_SEARCH_N paths per probe; linear in the paths per probe, so the 1e6-path figure is an
    extrapolation and says so."""
    import multiprocessing as mp

    import scenarios
    from oracle import ref as oracle_ref

    cfg = dict(scenarios.CONFIG_JSON, num_simulations_search=CPU_SEARCH_N)
    if oracle_ref.build():
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        p = ctx.Process(target=_reference_search_worker, args=(q, cfg, threads))
        p.start()
        dt, months, prob, probes = q.get(timeout=1800)
        p.join()
        kind = "reference"
    else:
        from oracle import oracle as orc

        orc.build()
        sim = orc.OracleSimulator(cfg, n_threads=threads)
        t0 = time.perf_counter()
        months, prob, curve, _ = sim.find_minimum_working_months()
        dt, probes, kind = time.perf_counter() - t0, len(curve), "port"
    return {"kind": kind, "cores": threads, "paths_per_probe": CPU_SEARCH_N, "probes": probes, "wall_s": dt,
            "selected_working_months": months, "probability": prob,
            "extrapolated_wall_s_at_1e6_paths_per_probe": dt * 1_000_000 / CPU_SEARCH_N}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    from oracle import ref as oracle_ref

    real = oracle_ref.build()
    vals = []
    for i in range(args.warmup + args.steps):
        if real:
            v, dt, _ = reference_throughput(REF_SAMPLE, threads)
        else:
            v, dt, _ = cpu_port_throughput(PORT_SAMPLE, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = statistics.mean(dt for _, dt in vals) * 1e3
    sample = REF_SAMPLE if real else PORT_SAMPLE
    what = ("the unmodified reference (oracle/_ref/backend/simulation.py, run_monte_carlo_simulations with "
            f"num_processes={threads})" if real else
            "C port of backend/simulation.py (oracle/path_oracle.c) + the reference's numpy draws; oracle/_ref not staged")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD + f"; CPU arm: a {sample}-path sample of it per step (linear in paths)",
                   "sample": f"{sample} paths x 720 months per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if real else "port",
                         "sample": f"{sample} paths x 720 months per step ({what})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": what,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# the CUDA arm
# -------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from monte_carlo_retirement_b200 import native
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_line = cpu_search = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (the CPU arms fork / spawn workers)
        threads = os.cpu_count() or 1
        cpu_line = cpu_baseline_block(threads)
        cpu_search = cpu_search_block(threads)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: NCCL prints its version banner to fd 1 when the first
        # communicator comes up, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    native.load_library()

    cfg = Config(**_scenario())
    n = PATHS_PER_GPU
    if world > 1:
        from monte_carlo_retirement_b200.parallel import ShardedSimulator

        # weak scaling: the GLOBAL job is world x 1e6 paths of one Philox stream, rank r owns
        # [r*1e6, (r+1)*1e6); aggregates (counts, histograms, exact quantile bands) are global.
        sim = ShardedSimulator(cfg, device=local_rank)
    else:
        sim = RetirementMonteCarloSimulator(cfg, device=local_rank)
    sim.use_final_seeds()
    n_job = n * world
    R = cfg.retirement_years
    months = WORKING_MONTHS + 12 * R
    ctx = sim.native_context
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(events=None):
        # world > 1: no data-path collective; counters, histograms and the radix-select digit
        # histograms (exact global bands) are all-reduced over NVLink inside aggregates_device
        return sim.aggregates_device(WORKING_MONTHS, n_job, bands=True, timeline_events=events,
                                     pipeline=args.pipeline)

    fp64_peak = ctx.fp64_peak_slots_per_s()
    for _ in range(max(args.warmup, 3)):
        agg = step()
    barrier()

    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        agg = step(k_events[i])
    agg.wait()  # pipelined reductions run on their own stream: the last step's must be inside the timed region
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in k_events)
    host = agg.to_host()
    executed = host["executed_path_months"] / world if world > 1 else host["executed_path_months"]

    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = n * months * world / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing API, host 7-tuple (rank-local shard) ----------
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(3):  # warm: two generations of pinned result blocks, sample columns, allocator
        tup = sim.run_monte_carlo_simulations(WORKING_MONTHS, n_job)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tup = sim.run_monte_carlo_simulations(WORKING_MONTHS, n_job)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n * months * world / float(e2e_t[0])
    h2d = 1024  # scenario constants + launch arguments (kernel parameter blocks); there is no bulk input
    d2h = int(getattr(sim, "last_d2h_bytes", n_job * 45))

    # ---- BASELINE.json configs[3]: the search at 1e6 paths per candidate (outside the timed C3 region)
    search = search_block(world, local_rank) if not args.no_search else None
    # ---- BASELINE.json configs[4]: 1e9 paths over the 8 GPUs of the box
    c5 = c5_block(world, local_rank) if (world >= 8 and not args.no_c5) else None

    if rank == 0:
        from monte_carlo_retirement_b200.build import source_hash

        T = sim._trajectory_len(WORKING_MONTHS)
        alg_bytes = n * ((2 * T + R) * 8 + 5 * 8 + 1 + 4)  # series + summary columns written per launch
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        counts = {}
        try:
            with open(os.path.join(ROOT, "profiles", "timeline_counts.json")) as f:
                counts = json.load(f)
        except Exception:
            pass
        fresh = bool(counts) and counts.get("source_hash") == source_hash()
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clock_hz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
        issue_peak = sm_count * 4 * 32 * clock_hz        # thread-instructions / s: 1 warp-instruction / clk / SMSP
        t_kernel = kernel_ms * 1e-3
        roofline = {
            "bound": "issue", "unit": "Tinst/s", "peak": issue_peak / 1e12,
            "peak_source": f"{sm_count} SMs x 4 sub-partitions x 32 lanes x {clock_hz / 1e6:.0f} MHz (SM clock sampled in this run)",
            "kernel": "k_timeline<fast, philox, bounded-return variant>", "kernel_ms": kernel_ms,
            "kernel_share_of_step": kernel_ms / ms_per_step,
            "executed_path_months_per_launch": executed,
            "why": "ncu: warp-instruction issue is the busiest unit (~70 %), FP64 pipe ~50 %, XU ~36 %, HBM 3 %; "
                   "tools/microbench/issue_model2.cu: integer multiplies / logic / selects issue at half rate on the "
                   "datapath they share with FP32, so the draws and selects, not the FP64 arithmetic, set the pace",
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / t_kernel / 1e9,
                    "peak_gbs": hbm_peak, "frac": alg_bytes / t_kernel / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"},
            # the reference-order census of SURVEY §8d, kept as a secondary figure: how many of the
            # reference's own FP64 issue slots per second this kernel retires (the closed forms need
            # far fewer FP64 instructions than the census, so this exceeds the FP64 peak)
            "reference_order": {"W_slots_per_path_month": W_SLOTS_PER_PATH_MONTH,
                                "slots_per_s": executed * W_SLOTS_PER_PATH_MONTH / t_kernel,
                                "fp64_peak_slots_per_s": fp64_peak,
                                "ratio_to_fp64_peak": executed * W_SLOTS_PER_PATH_MONTH / t_kernel / fp64_peak,
                                "peak_source": "DFMA-chain microbenchmark measured in this run (mcr_fp64_peak_slots_per_s)"},
        }
        if fresh:
            ipm, fpm = counts["instr_per_path_month"], counts["fp64_instr_per_path_month"]
            achieved = executed * ipm / t_kernel
            roofline.update({
                "achieved": achieved / 1e12, "frac": achieved / issue_peak,
                "traffic": counts["ncu"]["dram_bytes_read"] + counts["ncu"]["dram_bytes_write"],
                "issue": {"instr_per_path_month": ipm, "non_fp64_instr_per_path_month": ipm - fpm,
                          "issue_peak_tinst_s": issue_peak / 1e12, "frac": achieved / issue_peak,
                          "ncu_issue_active_pct": counts["ncu"]["issue_active_pct"]},
                "fp64_pipe": {"fp64_instr_per_path_month": fpm, "slots_per_s": executed * fpm / t_kernel,
                              "peak_slots_per_s": fp64_peak, "frac": executed * fpm / t_kernel / fp64_peak,
                              "ncu_fp64_pipe_active_pct": counts["ncu"]["fp64_pipe_active_pct"]},
                "counts_source": counts.get("source"), "counts_source_hash": counts.get("source_hash"),
            })
        else:
            roofline.update({"achieved": None, "frac": None, "traffic": None,
                             "stale_counts": "profiles/timeline_counts.json was taken from another build of the kernel "
                                             f"(hash {counts.get('source_hash')} != {source_hash()}); re-run "
                                             "tools/make_timeline_counts.py on an ncu capture of this build"})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD + (f"; cpu_baseline: a {REF_SAMPLE}-path sample of it (real reference) / "
                                               f"{PORT_SAMPLE}-path sample (C port), linear in paths"
                                               if cpu_line is not None else ""),
                       "paths_per_gpu": n, "months": months,
                       "l2": "no HBM inputs (counter-based Philox); each step writes 1.36 GB of outputs, > 126 MB L2",
                       "success_probability": host["success_probability"],
                       # N > 1: selects repeated with the stepwise protocol because the pooled shortcut
                       # could not finish a row (0 = the timed steps are the whole work)
                       "select_fallbacks": int(getattr(sim, "select_fallbacks", 0)),
                       "parity": "this kernel variant (bounded-return specialisation, lean month steps, MUFU normals) "
                                 "passes the <= 1e-9 / bit-exact-flags replay gate on the reference's numpy draws "
                                 "(tests/test_gpu_parity.py::test_benchmarked_variant_meets_the_replay_gate); its own "
                                 "Philox draws are statistically, not path-wise, comparable to numpy's",
                       "pipeline": bool(args.pipeline)},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(e2e_t[0]) * 1e3,
                    "api": f"RetirementMonteCarloSimulator.run_monte_carlo_simulations(240, {n_job}) -> host 7-tuple"
                           + ("; summary_df: all rows on rank 0 (shared pinned host block), own shard elsewhere"
                              if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        if search is not None:
            if cpu_search is not None:
                search["cpu"] = cpu_search
            line["search"] = search
        if c5 is not None:
            line["c5"] = c5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _sim_for(cfg_dict, world, local_rank, **kw):
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    cfg = Config(**cfg_dict)
    if world > 1:
        from monte_carlo_retirement_b200.parallel import ShardedSimulator

        return ShardedSimulator(cfg, device=local_rank, **kw)
    return RetirementMonteCarloSimulator(cfg, device=local_rank, **kw)


def search_block(world: int, local_rank: int):
    """configs[3]: config.json, 1e6 paths per candidate (global, sharded over the ranks).
    `auto`: the shipped policy (one search launch per probe at this size); `grid`: every month
    start..start+600 in ONE launch, the reference's decisions replayed over the table. Both read
    the same common-random-number table, so they must select the same month."""
    import torch

    import scenarios

    cfg = dict(scenarios.CONFIG_JSON, num_simulations_search=1_000_000)
    out = {"workload": "config.json, 1e6 paths per candidate, candidates 0..600 (grid) / the reference's probes (auto)",
           "n_gpus": world}
    results = {}
    for policy in ("auto", "grid"):
        sim = _sim_for(cfg, world, local_rank, search_policy=policy)
        sim.native_context  # noqa: B018  context creation is not part of the search
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        months, prob, curve = sim.find_minimum_working_months(verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        results[policy] = (months, prob)
        out[policy] = {"wall_s": dt, "selected_working_months": months, "probability": prob, "probes": len(curve),
                       "launches": sim.last_search_stats.get("launches"),
                       "candidates_evaluated": sim.last_search_stats.get("candidates_evaluated")}
    out["policies_agree"] = results["auto"] == results["grid"]
    return out


def c5_block(world: int, local_rank: int):
    """configs[4]: 1e9 paths of config.json at working_months = 233 over all ranks, nominal / real /
    withdrawal-rate bands, histograms and summary statistics; aggregate-only (nothing N-sized
    leaves the GPUs). Wall time of the second call (the first also pays the cudaMalloc of the
    series blocks)."""
    import torch
    import torch.distributed as dist

    import scenarios

    sim = _sim_for(dict(scenarios.CONFIG_JSON), world, local_rank)
    sim.use_final_seeds()
    n, wm = 1_000_000_000, 233
    walls = []
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        agg = sim.run_aggregates(wm, n, bands=True)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    months = wm + 12 * sim.params_model.retirement_years
    return {"workload": "config.json, 1e9 paths, wm=233, bands + histograms, aggregate-only", "n_gpus": world,
            "wall_s": walls[-1], "wall_s_each_call": walls, "nominal_path_months_per_s": n * months / walls[-1],
            "executed_path_months": agg["executed_path_months"], "success_probability": agg["success_probability"],
            "series_passes": [list(g) for g in sim.last_series_plan],
            "select_fallbacks": int(getattr(sim, "select_fallbacks", 0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search", action="store_true", help="skip the configs[3] search block")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] 1e9-path block (8 GPUs)")
    ap.add_argument("--pipeline", type=int, default=0, choices=[0, 1],
                    help="1: reductions of step i on a second stream, under the timeline kernel of step i+1")
    args = ap.parse_args()
    _silence_logs()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

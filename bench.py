#!/usr/bin/env python
"""Headline benchmark: MCMC proposals/sec on BASELINE config C2.

Workload (BASELINE.json configs[1]): 3D N^2-queens, N=12, full_3d state
space, random initial states, the five beta schedules {constant 5.0; linear, exponential,
logarithmic, sinusoidal 1.0->3.0}, 4096 replicas per schedule (20480 chains) per GPU,
n_steps = 1e6 proposals per chain.  One bench "step" = one pass of that batch through the hot
path (2.048e10 proposals per GPU), producing what the experiment consumes: per-schedule
sum E / sum E^2 per step (mean +- std curves), per-chain best / final energy, steps-to-best,
accept counts, 100-bin acceptance histograms and best / final states.

  value  proposals/s with inputs (seeds, schedule tables, group ids) resident in HBM and outputs
         left in HBM (MCQ_MEM_DEVICE), timed with CUDA events on the launching stream.
  e2e    the same pass through the public host-buffer API (Engine.run -> mcq_run, MCQ_MEM_HOST):
         pinned host inputs are copied H2D and every result is copied D2H inside the timed region.
  N>1    one process per GPU (torchrun); replicas are sharded (each GPU runs its own 4096
         replicas per schedule: weak scaling), no data-path collective; the per-schedule
         statistics and the global best energy are reduced once with NCCL inside the timed step.

`--impl reference` times the reference's CPU algorithm (oracle/queens_numpy.py, a line-by-line
NumPy restatement pinned to the reference by tests/golden; the reference itself is pure Python and
does not exist on the GPU box) on all host cores, same workload definition, bounded sample.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BOARD = 12
REPLICAS = 4096
CHAIN_STEPS = 1_000_000
BASE_SEED = 42
SCHEDULES = [
    {"type": "constant", "beta_const": 5.0},
    {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "logarithmic_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0},
]
# algorithmic warp-instructions per proposal, one warp per chain
I_ALG = {"full_3d": 48.0, "board": 40.0}
ISSUE_PER_CLK_PER_SM = 4
# ncu measurements of the dominant kernel on this workload (profiles/README.md says which capture)
AS_BUILT = {"source": "profiles/r1_spec_kernel_raw.txt (ncu --set full, chunk launch of 20480 chains x 26208 steps)",
            "warp_inst_per_proposal": 9.8, "issue_active_pct": 65.0, "warps_active_per_scheduler": 6.3,
            "registers_per_thread": 72, "smem_wavefronts_per_proposal": 2.60, "smem_wavefront_pct_of_peak": 66.2,
            "speculated_steps_per_round": 32, "proposals_retired_per_round": 19.4, "cycles_per_round_per_warp": 1950,
            "dram_bytes_per_launch": 1.057e9, "algorithmic_bytes_per_launch": 1.073e9}


_REAL_STDOUT = None


def emit(line):
    """Print the result line on the process's original stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "", 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def chain_seeds(rank, replicas):
    """Seed of replica r of any schedule on this rank: base_seed + global replica index (the
    reference gives every schedule the same base_seed, experiments.py:157,194)."""
    r = np.arange(replicas, dtype=np.uint64) + np.uint64(rank * replicas) + np.uint64(BASE_SEED)
    return np.tile(r, len(SCHEDULES))


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (NumPy port) on the host cores
# ----------------------------------------------------------------------------------------------
def _cpu_chain(args):
    from oracle import queens_numpy as qn
    sched_idx, seed, steps = args
    sched = qn.schedule_from_params(SCHEDULES[sched_idx], steps)
    r = qn.chain_full(N_BOARD, steps, "random", sched, seed=seed)
    return r["best_energy"]


def cpu_pass(cores, steps):
    """One bounded sample: `cores` chains (schedules round-robin) x `steps` proposals, one process per core
    (what run_experiment does with its ProcessPoolExecutor, experiments.py:513)."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(c % len(SCHEDULES), BASE_SEED + c, steps) for c in range(cores)]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as ex:
        best = list(ex.map(_cpu_chain, jobs))
    dt = time.perf_counter() - t0
    return cores * steps / dt, dt, best


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    steps = args.cpu_steps
    for _ in range(args.warmup):
        cpu_pass(cores, max(200, steps // 20))
    rates, times = [], []
    for _ in range(args.steps):
        pps, dt, _ = cpu_pass(cores, steps)
        rates.append(pps)
        times.append(dt)
    total = cores * steps * args.steps / sum(times)
    sample = f"{cores} chains x {steps} proposals per step (schedules round-robin), N=12 full_3d random init"
    line = {
        "impl": "reference", "metric": "MCMC proposals/sec", "value": total, "unit": "proposals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
        "config": workload_config(args, 1, cpu=True),
        "cpu_baseline": {"value": total, "unit": "proposals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": total, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world, cpu=False):
    return {
        "workload": "C2 single_N: N=12 full_3d, 5 beta schedules x %d replicas per GPU, %d proposals per chain, random init"
                    % (args.replicas, args.chain_steps),
        "N": N_BOARD, "mcmc_type": "full_3d", "schedules": [s["type"] for s in SCHEDULES],
        "replicas_per_schedule_per_gpu": args.replicas, "chains_total": args.replicas * len(SCHEDULES) * world,
        "chain_steps": args.chain_steps, "history": "per-schedule sum E / sum E^2 per step (stats) + per-chain results",
        "l2": "flushed between timed passes (256 MiB write)" if not cpu else "n/a",
        "parallelism": f"replica-sharded x{world}",
    }


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=REPLICAS, help="replicas per schedule per GPU")
    ap.add_argument("--chain-steps", type=int, default=CHAIN_STEPS, help="proposals per chain per pass")
    ap.add_argument("--cpu-steps", type=int, default=20000, help="proposals per chain in the CPU sample")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print on the way (NCCL's version banner, build
    # logs) is sent to stderr, and the real stdout is put back for the final print
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    if rank == 0 or not os.path.isfile(os.path.join(ROOT, "monte_carlo_collective_b200", "libmcq.so")):
        ge.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        dist.barrier()
    import monte_carlo_collective_b200 as mcq
    from monte_carlo_collective_b200 import schedules

    eng = mcq.Engine(local)
    ns, reps, ng = args.chain_steps, args.replicas, len(SCHEDULES)
    nc = reps * ng
    dev = torch.device(f"cuda:{local}")

    # ---- synthetic inputs ----
    seeds_h = chain_seeds(rank, reps)
    groups_h = np.repeat(np.arange(ng, dtype=np.int32), reps)
    seeds_d = torch.from_numpy(seeds_h.view(np.int64)).to(dev)
    groups_d = torch.from_numpy(groups_h).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    out_d = {}

    def device_pass():
        r = eng.run("full_3d", N_BOARD, ns, seeds_d, schedules=SCHEDULES, groups=groups_d, history="stats",
                    n_bins=100, device_buffers=True, lanes_per_chain=args.lanes, stream=stream.cuda_stream, out=out_d)
        for k in ("stat_sum_e", "stat_sum_e2", "best_energy", "final_energy", "steps_to_best", "n_accepted",
                  "steps_done", "initial_energy", "final_state", "best_state", "accept_hist"):
            out_d[k] = getattr(r, k)
        if world > 1:   # the only exchange of the path: final reductions over NCCL
            dist.all_reduce(r.stat_sum_e, op=dist.ReduceOp.SUM)
            dist.all_reduce(r.stat_sum_e2, op=dist.ReduceOp.SUM)
            gmin = r.best_energy.min().reshape(1)
            dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
            acc = r.n_accepted.sum(dtype=torch.int64).reshape(1)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        return r

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_pass()
        flush.fill_(1)
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, launches = 0.0, 0
    with ClockSampler(local) as clocks:
        for a, b in ev:
            flush.fill_(1)
            a.record(stream)
            r = device_pass()
            b.record(stream)
            kernel_ms += r.kernel_ms
            launches += r.gpu_launches
        sync_all()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        clk = clocks.summary()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    proposals_per_pass = nc * ns * world
    value = proposals_per_pass * args.steps / (total_ms * 1e-3)
    gmin = r.best_energy.min().to(torch.int64).reshape(1)
    gacc = r.n_accepted.sum(dtype=torch.int64).reshape(1)
    if world > 1:   # job-wide figures for the report line (the timed reduction is inside device_pass)
        dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(gacc, op=dist.ReduceOp.SUM)
    best_min = int(gmin.item())
    acc_rate = float(gacc.item()) / (nc * ns * world)

    # ---- end to end through the host-buffer API ----
    e2e = None
    if not args.no_e2e:
        out_h = {}
        h2d = seeds_h.nbytes + groups_h.nbytes + 32 * ng   # seeds, group ids, five schedule parameter records
        times = []
        for it in range(min(args.warmup, 1) + args.steps):
            sync_all()
            t0 = time.perf_counter()
            rh = eng.run("full_3d", N_BOARD, ns, seeds_h, schedules=SCHEDULES, groups=groups_h,
                         history="stats", n_bins=100, lanes_per_chain=args.lanes, out=out_h)
            sync_all()
            dt = time.perf_counter() - t0
            if it >= min(args.warmup, 1):
                times.append(dt)
            for k in ("stat_sum_e", "stat_sum_e2", "best_energy", "final_energy", "steps_to_best", "n_accepted",
                      "steps_done", "initial_energy", "final_state", "best_state", "accept_hist"):
                out_h[k] = getattr(rh, k)
        d2h = sum(out_h[k].nbytes for k in out_h)
        tmax = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e = {"value": proposals_per_pass * args.steps / float(tmax.item()), "unit": "proposals/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "Engine.run(host NumPy buffers) -> mcq_run(MCQ_MEM_HOST)"}

    # ---- roofline of the dominant kernel (spec_kernel<FULL>), DESIGN.md section 4 ----
    # Bound: SM issue slots (4 warp-instructions / clk / SM).  Algorithmic warp-instructions per proposal of
    # the speculative mapping: a round of 32 lanes costs 54 and retires
    # adv(p) = (1-(1-p)^32)/p proposals at acceptance probability p; the random words of a step cost
    # 64/32 (one Philox4x32-10 call + ring store per lane per 32 consumed steps); an accepted move costs 70.
    kernel_pps = nc * ns * args.steps / (kernel_ms * 1e-3)        # this rank's kernel-only rate (CUDA events in mcq_run)
    f_mhz = clk["sm_mhz"] or 1965.0
    peak = ISSUE_PER_CLK_PER_SM * eng.sm_count * f_mhz * 1e6
    p_acc = max(acc_rate, 1e-6)
    p_round = 1.0 - (1.0 - p_acc) ** 32
    i_alg = 54.0 * p_acc / p_round + 2.0 + 70.0 * p_acc
    achieved = kernel_pps * i_alg
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.isfile(peaks_file) else 6650.0
    roofline = {
        "bound": "issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "Gwarp-inst/s",
        "frac": achieved / peak, "traffic": AS_BUILT["dram_bytes_per_launch"],
        "kernel": "spec_kernel<FULL=1,REPLAY=0,EARLY=0,NR=5,LPC=32,N=12,HIST=u16>", "i_alg_warp_inst_per_proposal": i_alg,
        "i_alg_formula": "54/adv(p) + 2 + 70*p, adv(p) = (1-(1-p)^32)/p, p = measured acceptance rate",
        "kernel_proposals_per_s": kernel_pps, "sm_clock_mhz_used": f_mhz, "sm_count": eng.sm_count,
        "as_built": AS_BUILT,
        "frac_under_survey_mapping": kernel_pps * I_ALG["full_3d"] / peak,
        # shared-memory utilisation with the survey's W_alg = 3 + 1/(N-1) + 3 p wavefronts per proposal
        "smem_util_survey": kernel_pps * (3.0 + 1.0 / 11.0 + 3.0 * p_acc) / (eng.sm_count * f_mhz * 1e6),
        "hbm": {"algorithmic_bytes_per_proposal": 2.0, "achieved_gbs": kernel_pps * 2.0 / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if os.path.isfile(peaks_file) else "fallback"},
        "note": "frac = pps * I_alg(p) / (4 * n_SM * f_measured); as_built = ncu on the same kernel (profiles/); "
                "frac_under_survey_mapping uses SURVEY 8(d)'s one-warp-per-chain I_alg = 48, which the speculative "
                "conflict-table mapping is designed to beat (hence > 1)",
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        cpu_pps, dt, _ = cpu_pass(cores, args.cpu_steps)
        cpu_baseline = {"value": cpu_pps, "unit": "proposals/s", "cores": cores, "kind": "port",
                        "sample": f"{cores} chains x {args.cpu_steps} proposals (schedules round-robin), N=12 full_3d, "
                                  f"{dt:.1f} s wall, oracle/queens_numpy.py (NumPy restatement of the reference)"}

    if rank == 0:
        line = {
            "metric": "MCMC proposals/sec", "value": value, "unit": "proposals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "min_energy_reached": best_min, "acceptance_rate": acc_rate,
            "metric_full": "MCMC proposals/sec (N=12, 1/2/4/8 B200) vs host-CPU ref; min energy reached",
            "dtype_note": "uint8 conflict table, int32 energies, float32 ex2 accept threshold on a 32-bit uniform word",
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

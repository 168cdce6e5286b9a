#!/usr/bin/env python
"""bench.py — throughput of the conclave voting engine on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] — the README "optimized" parameters
(beta 1.5 +0.3/round, bandwagon 0.7, papabile 2.0, stickiness 0.8, threshold 0.56, run-off 5/2,
max 20 rounds) on a synthetic 135-elector roster (SURVEY §8d recipe), `ref_cli` wiring (what the
reference CLI computes).  configs[1] names 1000 trials, which a B200 finishes in well under a
millisecond; one STEP here is therefore one batch of --trials-per-gpu trials per GPU (default
1,250,000 = configs[2]'s 10M-trial job split over 8 GPUs, so --gpus 8 IS configs[2]; weak scaling).

`value` / `roofline`: one step = one launch of the fp32 DENSE trial kernel (every trial re-evaluates its E x C scores each
  round: the canonical per-trial formulation north_star describes and SURVEY 8d puts the roofline on) + ONE NCCL all-reduce
  of the winner / rounds histograms issued through the C ABI (csim_allreduce_hists); inputs resident in HBM.
`e2e`: the same engine through the PUBLIC call a user makes — conclave_sim_b200.simulate.run_simulations(DataFrame, ...) — with
  host buffers: roster + parameters uploaded and the per-trial rows + histograms copied back inside every timed step; under
  torchrun that one call shards the sim ids over the ranks, all-reduces the histograms and gathers the rows on rank 0.
`engines`: every other engine on the same workload, each with its own timed steps (never mixed into value / roofline):
  `auto` = what CSIM_ENGINE_AUTO — the default of the public API — runs for this workload (the TABLE engine), first-class with
  its own e2e; prefix (both wirings), full (fatigue + stop-candidate live), dense under the `model` wiring.
`configs` (N = 1 only): the remaining BASELINE configurations and the bit-exact fp64 engines, each timed with CUDA events:
  fp64 table (bit-exact) at the full step size, exact64 under the `model` wiring, cfg-4 (256 sets x 100 k), cfg-5
  (1 M x 1024^2, R = 50), and the literal cfg-1 / cfg-2 sizes through the host API.

The oracle (oracle/) is only ever executed here for the cpu_baseline leg and for --impl reference.
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E_HEADLINE = 135
CFG2 = dict(initial_beta_weight=1.5, enable_dynamic_beta=True, beta_increment_amount=0.3,
            beta_increment_interval_rounds=1, stickiness_factor=0.8, bandwagon_strength=0.7,
            regional_affinity_bonus=0.1, papabile_weight_factor=2.0)
CFG1 = dict(initial_beta_weight=1.0, beta_increment_amount=0.05, beta_increment_interval_rounds=10,
            stickiness_factor=0.5, regional_affinity_bonus=0.1, papabile_weight_factor=1.5)        # the CLI defaults (src/simulate.py:236-244)
ENGINE_KW = dict(max_rounds=20, supermajority_threshold=0.56, runoff_threshold_rounds=5, runoff_candidate_count=2)
SEED = 20250507
# canonical dense work per (elector, candidate) pair
FP32_PER_PAIR, MUFU_PER_PAIR = 10, 1
# what the factorised kernel EXECUTES per pair in its inner loop: one packed multiply, two mins, one packed add per TWO pairs
FACT_ISSUE_SLOTS_PER_PAIR = 2.0
FALLBACK_FP32_GINSTR = 148 * 128 * 1.965     # architectural: SMs x FP32 lanes x GHz (used only if the microbench fails)
FALLBACK_MUFU_GINSTR = 148 * 16 * 1.965
NCU_DIGEST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02b_ncu_digest.json")


def synthetic_roster(E: int, seed: int = 20250507):
    """SURVEY §8(d) recipe (kept here so the product arm needs nothing from oracle/)."""
    rng = np.random.default_rng(seed)
    ideology = rng.beta(2, 4, E)
    p = np.array([53, 23, 18, 18, 10, 5, 4, 3], dtype=np.float64)
    region = rng.choice(8, size=E, p=p / p.sum()).astype(np.int32)
    pap = rng.random(E) < 0.15
    return ideology, region, pap


def cfg4_grid():
    """4 x 4 x 4 x 4 = 256 parameter sets (beta0, beta increment, papabile factor, threshold), cfg-2 otherwise."""
    return list(itertools.product([0.5, 1.0, 1.5, 2.0], [0.0, 0.1, 0.3, 0.5], [1.0, 1.5, 2.0, 3.0], [0.5, 0.56, 0.6, 2 / 3]))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        # the first samples may predate the load; take the median of the upper half
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pair_evals(rounds: np.ndarray, E: int, rr: int, k: int) -> int:
    """Algorithmic (elector, candidate) score evaluations of a batch: E*C in full-field rounds,
    E*k in run-off rounds (rounds after round rr; simulate.py:144-154,183-186)."""
    rounds = rounds.astype(np.int64)
    full = np.minimum(rounds, rr) if k > 0 else rounds
    return int(full.sum()) * E * E + int((rounds - full).sum()) * E * k


def ncu_digest(kernel_key: str):
    """The committed ncu evidence for one kernel (profiles/r02b_ncu_digest.json, written by profiles/ncu_summary.py --digest from the
    .ncu-rep captures of this round): a pointer + the few numbers, read from the file — nothing is hard-coded here."""
    try:
        with open(NCU_DIGEST) as f:
            d = json.load(f)
        rec = d.get(kernel_key)
        if rec is not None:
            rec = dict(rec, source=f"profiles/{os.path.basename(NCU_DIGEST)}")
        return rec
    except Exception:
        return None


def cpu_oracle_rate(trials: int, threads: int, E: int, faithful: bool = True, model_kw=None, engine_kw=None):
    """Times the C oracle (oracle/conclave_oracle.c, OpenMP) on `trials` trials of the workload.
    faithful=True: the whole E x C matrix (exp included) is re-evaluated per trial per round, as the
    reference does (src/simulate.py:121, src/model.py:311-486) -- the like-for-like baseline of the
    dense GPU engine.  faithful=False: the same port with trial-invariant per-round CDF tables."""
    from oracle import conclave_oracle as O
    from oracle import c_oracle as CO
    CO.build()
    CO.set_faithful(faithful)
    roster = O.synthetic_roster(E)
    mp = O.ModelParams(**(model_kw or CFG2))
    ep = O.EngineParams(**(engine_kw or dict(max_rounds=20, supermajority_threshold=0.56)))
    t0 = time.perf_counter()
    out = CO.run(roster, [(mp, ep)], O.WIRING_REF_CLI, trials, seed=SEED, want_final=False, threads=threads)
    dt = time.perf_counter() - t0
    er = int(out["totals"][0, 0]) * E
    return er / dt, trials / dt, dt


PYTHON_REFERENCE_NOTE = ("the UNMODIFIED Python reference cannot travel to the GPU box (/root/reference exists only in the build container); "
                         "measured THERE on 8 cores, not on this box: 38.2 conclaves/s = 5.87e4 elector-rounds/s on cfg-2 "
                         "(20 000 independently seeded trials, tests/golden/dist_cfg2.json), 31.7 conclaves/s on cfg-1 defaults, "
                         "4.8 conclaves/s on the cfg-5 shape (tests/golden/dist_cfg1.json, dist_cfg5.json)")


def reference_arm(args):
    """--impl reference: the CPU implementation of the path on this box's host cores.  The
    reference itself is Python and cannot travel to the GPU box (and /root/reference does not exist
    there), so this is the oracle PORT (kind "port"): the C restatement, OpenMP over all cores,
    re-evaluating the E x C matrix per trial per round exactly as the reference does (the dense
    formulation the GPU arm is timed on).  The table-optimised variant of the same port is reported
    beside it (cpu_baseline.table_optimised_value)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    probe_er, probe_c, _ = cpu_oracle_rate(2000, cores, E_HEADLINE)
    per_step = max(2000, int(probe_c * 4.0))          # ~4 s of CPU work per step
    for _ in range(args.warmup):
        cpu_oracle_rate(max(500, per_step // 8), cores, E_HEADLINE)
    t_er, t_c, t_dt = [], [], 0.0
    for _ in range(args.steps):
        er, c, dt = cpu_oracle_rate(per_step, cores, E_HEADLINE)
        t_er.append(er * dt); t_c.append(c * dt); t_dt += dt
    value = sum(t_er) / t_dt
    tab_er, tab_c, _ = cpu_oracle_rate(max(20000, int(probe_c * 40)), cores, E_HEADLINE, faithful=False)
    line = {
        "impl": "reference", "metric": "elector_rounds_per_sec", "value": value, "unit": "elector-rounds/s",
        "conclaves_per_sec": sum(t_c) / t_dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg2 README params, synthetic E={E_HEADLINE}, ref_cli wiring, max_rounds 20",
                   "trials_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "elector-rounds/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} trials per step x {args.steps} steps, C oracle (dense per-trial matrix like the "
                                   f"reference), OpenMP on {cores} threads",
                         "table_optimised_value": tab_er, "python_reference": PYTHON_REFERENCE_NOTE},
        "e2e": {"value": value, "unit": "elector-rounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trials-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the N = 1 `configs` block (cfg-4 / cfg-5 / fp64 engines)")
    ap.add_argument("--no-reduce", action="store_true", help="diagnostic: skip the histogram all-reduce in --only mode (each rank's own step time)")
    ap.add_argument("--only", default=None, help="profiling aid: run only the timed steps of ONE engine (dense|auto|table|prefix|prefix_model|full) and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    # stdout carries ONE JSON line: libraries that print banners to fd 1 (NCCL's version line) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from conclave_sim_b200 import (Engine, make_params, FP32, FP64, WIRING_REF_CLI, WIRING_MODEL, ENGINE_AUTO, ENGINE_DENSE, ENGINE_TABLE,
                                   ENGINE_PREFIX, ENGINE_FULL)
    from conclave_sim_b200.dist import DeviceBatch, shard_range, csim_comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    E = E_HEADLINE
    ideology, region, pap = synthetic_roster(E)
    eng = Engine(local_rank)
    eng.upload_roster(ideology, region, pap)
    params = make_params(CFG2, **ENGINE_KW)
    R, rr, k = ENGINE_KW["max_rounds"], ENGINE_KW["runoff_threshold_rounds"], ENGINE_KW["runoff_candidate_count"]
    n_local = args.trials_per_gpu
    n_global = n_local * world
    comm = csim_comm(eng)                                 # ncclComm_t made through the C ABI (None on one GPU): the data path's only collective
    if os.environ.get("CSIM_BENCH_TORCH_ALLREDUCE"):      # diagnostic: torch.distributed's all_reduce instead
        comm = None
    batch = DeviceBatch(eng, 1, n_local, R, want_reason=True, comm=comm)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i: int, reduce: bool = True, engine: int = ENGINE_DENSE, wiring: int = WIRING_REF_CLI, p=None, precision: int = FP32, b=None,
             sets=None):
        b = batch if b is None else b
        begin, _ = shard_range(n_global, rank, world)
        b.run(sets if sets is not None else [params if p is None else p], sim_id_begin=i * n_global + begin, seed=SEED, wiring=wiring,
              engine=engine, precision=precision, reduce=reduce, job_sims=b.n_sims_local * world)

    def timed(engine: int, wiring: int, base: int, p=None, precision: int = FP32, b=None, sets=None, steps=None, warm: int = 2):
        """`steps` timed steps of one engine: L2 flushed between steps (untimed), CUDA events on the launch stream, max over ranks."""
        b = batch if b is None else b
        steps = args.steps if steps is None else steps
        n_sets = 1 if sets is None else len(sets)
        for i in range(warm):
            step(base + i, engine=engine, wiring=wiring, p=p, precision=precision, b=b, sets=sets)
        barrier()
        name = eng.last_run_info()[0:2]
        l0 = eng.stats()[0]
        tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        rounds_dev = torch.zeros((), dtype=torch.int64, device=dev)
        t_host0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()
            if world > 1:                      # untimed, like the flush: every rank starts the step together.  Without it rank 0 enters
                barrier()                      # each step ~2 ms late (its host-side gap between steps, rank_untimed_gap_ms) and the
            tev[i][0].record()                 # all-reduce bills that wait to every other rank's step (8.5 vs 10.3 ms at 2 and 8 GPUs)
            step(base + warm + i, engine=engine, wiring=wiring, p=p, precision=precision, b=b, sets=sets, reduce=not args.no_reduce)
            tev[i][1].record()
            rounds_dev += b.totals.view(n_sets, 2)[:, 0].sum()                 # already summed over ranks by the step's all-reduce
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps            # host time to ENQUEUE a step (it runs ahead of the GPU)
        barrier()
        k_ms = eng.stats()[1] * steps                                          # the trial kernel of the last step (every step is the same size)
        t_rounds = int(rounds_dev.item())
        l1 = eng.stats()[0]
        gap = sum(tev[i][1].elapsed_time(tev[i + 1][0]) for i in range(steps - 1)) / max(1, steps - 1)     # untimed: L2 flush + bookkeeping
        tt = torch.tensor([sum(a.elapsed_time(c) for a, c in tev), k_ms, gap, host_enqueue_ms], dtype=torch.float64, device=dev)
        per_rank = [round(tt[0].item() / steps, 4)]
        per_rank_gap = [round(gap, 4)]
        per_rank_host = [round(host_enqueue_ms, 4)]
        if world > 1:
            every = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(every, tt)
            per_rank = [round(x[0].item() / steps, 4) for x in every]
            per_rank_gap = [round(x[2].item(), 4) for x in every]
            per_rank_host = [round(x[3].item(), 4) for x in every]
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt[0].item()
        trials = b.n_sims_local * world * n_sets * steps
        Eb = eng.n_electors
        return {"value": float(t_rounds) * Eb / (ms * 1e-3), "unit": "elector-rounds/s", "conclaves_per_sec": trials / (ms * 1e-3),
                "ms_per_step": ms / steps, "trial_kernel_ms_per_step": tt[1].item() / steps, "steps": steps,
                "trials_per_step": trials // steps, "mean_rounds": float(t_rounds) / trials, "engine_ran": name[0], "precision_ran": name[1],
                "gpu_launches": int(l1 - l0), "rank_ms_per_step": per_rank, "rank_untimed_gap_ms": per_rank_gap, "rank_host_enqueue_ms_per_step": per_rank_host}

    if args.only:
        sel = {"dense": (ENGINE_DENSE, WIRING_REF_CLI, None), "auto": (ENGINE_AUTO, WIRING_REF_CLI, None), "table": (ENGINE_TABLE, WIRING_REF_CLI, None),
               "prefix": (ENGINE_PREFIX, WIRING_REF_CLI, None), "prefix_model": (ENGINE_PREFIX, WIRING_MODEL, None),
               "full": (ENGINE_FULL, WIRING_MODEL, make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3),
                                                               enable_candidate_fatigue=True, enable_stop_candidate=True, **ENGINE_KW))}[args.only]
        r = timed(sel[0], sel[1], 100, p=sel[2], warm=args.warmup)
        os.dup2(real_stdout, 1)
        if rank == 0:
            print(json.dumps({"only": args.only, **r}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region: K steps, CUDA events on the launch stream, L2 flushed between steps (not timed)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0, _ = eng.stats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, tot_rounds, tot_pairs, kept_rounds = [], 0, 0, []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
        ev[i][1].synchronize()
        kernel_ms.append(eng.stats()[1])
        kept_rounds.append(batch.rounds.clone())          # device-side copy after the step's end event: not timed, and no host
                                                          # work between steps that would skew the ranks against each other
    barrier()
    for rd in kept_rounds:
        rounds_host = rd.cpu().numpy()
        tot_rounds += int(rounds_host.astype(np.int64).sum())
        tot_pairs += pair_evals(rounds_host, E, rr, k)
    del kept_rounds
    launches1, _ = eng.stats()
    clocks = sampler.stop()
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_local, float(tot_rounds), float(tot_pairs), float(sum(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, all_rounds, all_pairs, kern_ms_max = tmax[0].item(), tsum[1].item(), tsum[2].item(), tmax[3].item()
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        rank_ms = [round(x[0].item() / args.steps, 4) for x in every]     # per-rank step time: `value` uses the slowest
    else:
        ms_total, all_rounds, all_pairs, kern_ms_max = t[0].item(), t[1].item(), t[2].item(), t[3].item()
        rank_ms = [round(ms_total / args.steps, 4)]
    value = all_rounds * E / (ms_total * 1e-3)
    conclaves = n_global * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (k_trials_dense32), per launch, this rank.
    # frac = EXECUTED work against the measured peak of the same instruction mix: pair evaluations per second over the rate at
    # which this GPU retires the kernel's own inner-loop pair body (csim_microbench(3): packed FMUL2 / FMNMX / FADD2).  The
    # canonical SURVEY-8d accounting (10 FP32 + 1 MUFU per pair, which the factorised kernel does NOT execute) is kept beside it
    # as `algorithmic_equivalent_frac` so that rounds stay comparable.
    try:
        peak_fp32 = eng.microbench(0); peak_mufu = eng.microbench(1); peak_mix = eng.microbench(2); peak_fact = eng.microbench(3)
        peak_src = "measured on this GPU by csim_microbench (k_microbench<0..3>); MEASURED_PEAKS.json has only HBM / bf16 entries, which do not bound this kernel"
    except Exception:
        peak_fp32, peak_mufu, peak_mix, peak_fact, peak_src = FALLBACK_FP32_GINSTR, FALLBACK_MUFU_GINSTR, None, None, "architectural fallback (148 SMs x 128 lanes x 1.965 GHz)"
    pairs_per_launch = tot_pairs / args.steps
    kern_s = (sum(kernel_ms) / args.steps) * 1e-3
    pair_rate = pairs_per_launch / kern_s
    fact_peak_pairs = (peak_fact / 2.0 * 1e9) if peak_fact else FALLBACK_FP32_GINSTR * 1e9 / 2.0     # pair body = 2 thread-instr per pair (1 FMUL2 + 2 FMNMX + 1 FADD2 per two pairs)
    out_bytes = n_local * 9 + (E + 1 + R + 1 + 2) * 8
    dig = ncu_digest("k_trials_dense32")
    roofline = {"bound": "fp32", "kernel": "k_trials_dense32<LC=8,MODEL=false,FACT=true>",
                "achieved": pair_rate * 1e-9, "peak": fact_peak_pairs * 1e-9, "unit": "G pair-evaluations/s", "frac": pair_rate / fact_peak_pairs,
                "algorithmic": f"{pairs_per_launch:.4g} (elector, candidate) pair evaluations per launch (E*C in full-field rounds, E*k in run-off "
                               f"rounds; SURVEY 8d) over the kernel's CUDA-event duration; peak = the rate at which this GPU retires the kernel's own "
                               f"inner-loop pair body exactly as its SASS has it (1 FMUL2 + 2 FMNMX + 1 FADD2 per two pairs, k_microbench<3>; the round-1 "
                               f"body with two FMUL2 — 1.007e13 pairs/s — was heavier than what the round-2 kernel executes and is no longer the denominator)",
                "peak_source": peak_src, "kernel_ms_per_launch": kern_s * 1e3,
                "algorithmic_equivalent_frac": pairs_per_launch * FP32_PER_PAIR / kern_s * 1e-9 / peak_fp32,
                "algorithmic_equivalent_note": f"the canonical accounting of SURVEY 8d ({FP32_PER_PAIR} FP32 + {MUFU_PER_PAIR} MUFU per pair) against the measured "
                                               f"FFMA rate ({peak_fp32:.4g} G thread-instr/s); the factorised kernel issues ~{FACT_ISSUE_SLOTS_PER_PAIR:g} slots per "
                                               "pair and no MUFU, so this is NOT pipe utilisation — kept only so that rounds stay comparable",
                "measured_peaks_ginstr": {"ffma": peak_fp32, "mufu_ex2": peak_mufu, "mufu_pair_body": peak_mix, "factorised_pair_body": peak_fact},
                "traffic": None if dig is None else dig.get("dram_bytes_per_launch"),
                "traffic_note": "dram__bytes_read + dram__bytes_write of ONE launch of this kernel at this step size from the committed ncu --set full "
                                "capture (read from profiles/r02b_ncu_digest.json at run time; null when the digest is absent); algorithmic HBM bytes per "
                                f"launch = {out_bytes} (9 B/trial result rows + histograms): the kernel is compute-bound",
                "hbm_writeback_gbs": out_bytes / kern_s * 1e-9, "ncu": dig}

    # ---- the other engines on the same workload, reported SEPARATELY from the dense kernel's value / roofline
    auto_engine = timed(ENGINE_AUTO, WIRING_REF_CLI, 500)
    auto_engine.update(kernel="k_trials_table<float> (+ k_base_scores64, k_cdf_tables64, k_f64_to_f32 per step)",
                       note="what CSIM_ENGINE_AUTO — the default of run_simulations / the CLI — runs for this workload: under ref_cli the probability "
                            "matrix is trial-invariant, so CDF rows are built once per step and a draw is an 8-probe binary search in shared memory; "
                            "NOT comparable with the dense roofline",
                       bound={"kind": "issue slots / shared-memory latency (dependent 8-probe searches)", "ncu": ncu_digest("k_trials_table")})
    prefix_engine = {"ref_cli": timed(ENGINE_PREFIX, WIRING_REF_CLI, 600), "model": timed(ENGINE_PREFIX, WIRING_MODEL, 700),
                     "kernel": "k_trials_prefix<LC8, MODEL, STAGE> (+ k_prefix_tables32 per step)",
                     "note": "trial-invariant chunk-prefix tables + per-trial bandwagon / stickiness corrections: a draw is a binary search over 17 chunk "
                             "prefixes and a re-scan of one chunk of 8 candidates; both wirings; what ENGINE_AUTO picks for the `model` wiring",
                     "bound": {"kind": "issue slots, then shared-memory wavefronts", "ncu": ncu_digest("k_trials_prefix")}}
    params_x = make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), enable_candidate_fatigue=True,
                           enable_stop_candidate=True, **ENGINE_KW)
    full_engine = timed(ENGINE_FULL, WIRING_MODEL, 900, p=params_x)
    full_pairs = pair_evals(batch.rounds.cpu().numpy(), E, rr, k)             # this rank's last step (every step is the same size)
    full_kern_s = full_engine["trial_kernel_ms_per_step"] * 1e-3
    full_engine.update(kernel="k_trials_dense32<LC=8,MODEL=true,FACT=true,EXT=true> (the FULL engine's fast path, launched by eng_full.cu)",
                       note="`model` wiring with EVERY term live: cfg-2 parameters + candidate fatigue (3 rounds below 5 %) + stop-candidate "
                            "(unacceptable distance 0.3).  Round 2: the dense kernel's factorised two-elector pair loop with fatigue folded into a "
                            "per-warp copy of the round table and stop-candidate as a byte-masked second accumulator; k_trials_full32 (MUFU form, "
                            "every term inline) remains for negative factors, E > 255 and non-factorisable beta schedules",
                       roofline={"bound": "fp32", "unit": "G pair-evaluations/s", "achieved": full_pairs / full_kern_s * 1e-9, "peak": fact_peak_pairs * 1e-9,
                                 "frac": full_pairs / full_kern_s / fact_peak_pairs,
                                 "note": "executed work: pair evaluations per second over the rate of the PLAIN factorised pair body (k_microbench<3>); in "
                                         "rounds with stop-candidate threats the masked body executes 5.5 issue slots of arithmetic per pair (~8 with its operand loads) instead of 2, so this "
                                         "fraction understates the pipe utilisation of those rounds",
                                 "algorithmic_equivalent_frac": full_pairs * FP32_PER_PAIR / full_kern_s * 1e-9 / peak_fp32,
                                 "algorithmic_equivalent_note": "SURVEY-8d canonical accounting (10 FP32 per pair) against the measured FFMA rate; round 1 "
                                                                "reported this figure (0.26) for the MUFU-form kernel, which executed it",
                                 "ncu": ncu_digest("k_trials_dense32_ext")})
    if not args.no_configs and world == 1:
        os.environ["CSIM_FULL_LEGACY"] = "1"                      # the same step on k_trials_full32 (what round 1 and the first half of round 2 ran)
        try:
            lg = timed(ENGINE_FULL, WIRING_MODEL, 950, p=params_x, steps=min(args.steps, 3), warm=1)
            full_engine["legacy_k_trials_full32"] = {k_: lg[k_] for k_ in ("value", "unit", "ms_per_step", "trial_kernel_ms_per_step", "steps")}
        finally:
            os.environ.pop("CSIM_FULL_LEGACY", None)
    model_wiring_dense = timed(ENGINE_DENSE, WIRING_MODEL, 800)
    model_wiring_dense.update(kernel="k_trials_dense32<LC=8,MODEL=true,FACT=true>",
                              note="`model` wiring (bandwagon + stickiness live, SURVEY 8f.1) on the dense kernel, for comparison")

    # ---- e2e: the PUBLIC call (simulate.run_simulations), host buffers, every step: roster + parameters up, rows + histograms back
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    df = pd.DataFrame({"ideology_score": ideology, "region": region, "is_papabile": pap},
                      index=pd.Index([str(i) for i in range(E)], name="elector_id"))
    api_kw = dict(max_rounds=R, supermajority_threshold=ENGINE_KW["supermajority_threshold"], runoff_threshold_rounds=rr,
                  runoff_candidate_count=k)

    def api_e2e(wiring: str, engine: str, base: int):
        """run_simulations(DataFrame, ...): the roster is re-uploaded every step, results land on the host (rank 0 receives every rank's
        rows under torchrun).  CUDA events around the call on this rank's stream + the wall clock; max over ranks."""
        S.RUNTIME.update(seed=SEED, device=local_rank, precision="fp32", engine=engine, wiring=wiring)
        warm = [S.run_simulations(df, CFG2, n_global, sim_id_begin=j, **api_kw) for j in range(2)]   # two live results: the
        del warm                                                  # page-locked pool reaches its steady-state size before the clock starts
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        rounds_sum, ran = 0, None
        for i in range(args.steps):
            S.get_engine(local_rank)._roster_key = None           # force the roster upload inside the timed region
            r = S.run_simulations(df, CFG2, n_global, sim_id_begin=(base + i) * n_global, **api_kw)
            rounds_sum += int(r.totals[0, 0])                     # the reduced histograms: the whole job's rounds on every rank
            ran = (r.engine, r.precision)
        e1.record()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_wall, dt_dev = tt[0].item(), tt[1].item()
        return {"value": rounds_sum * E / dt_wall, "unit": "elector-rounds/s", "conclaves_per_sec": n_global * args.steps / dt_wall, "steps": args.steps,
                "ms_per_step": 1e3 * dt_wall / args.steps, "ms_per_step_cuda_events": 1e3 * dt_dev / args.steps, "engine_ran": ran[0], "precision_ran": ran[1],
                "api": f"conclave_sim_b200.simulate.run_simulations(DataFrame, cfg2, {n_global} trials, wiring='{wiring}', engine='{engine}')"
                       + (f" inside the {world}-rank torchrun job: sim ids sharded over the ranks, ONE NCCL all-reduce of the histograms through "
                          "csim_allreduce_hists, every rank's rows gathered on rank 0" if world > 1 else "")}

    e2e_dense = api_e2e("ref_cli", "dense", 1000)
    auto_engine["e2e"] = api_e2e("ref_cli", "auto", 2000)
    prefix_engine["model"]["e2e"] = api_e2e("model", "auto", 3000)
    S.RUNTIME.update(wiring="ref_cli", engine="auto")
    h2d = E * (8 + 4 + 1) + 152 + R * 12
    d2h = n_local * 9 + ((E + 1) + (R + 1) + 2) * 8

    line = {
        "metric": "elector_rounds_per_sec", "value": value, "unit": "elector-rounds/s",
        "conclaves_per_sec": conclaves, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "rank_ms_per_step": rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2 README params x {n_local} trials/GPU/step (cfg3 per-GPU share), synthetic E={E}, "
                               "ref_cli wiring, dense fp32 engine, Philox4x32-10",
                   "trials_per_step_global": n_global, "max_rounds": R, "l2": "flushed between steps (256 MiB memset, untimed)",
                   "collective": "1 NCCL all-reduce (int64 sum) of winner/rounds histograms per step, issued through the C ABI (csim_allreduce_hists)"
                                 if world > 1 else "none (1 GPU)"},
        "roofline": roofline,
        "e2e": {"value": e2e_dense["value"], "unit": "elector-rounds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                **{kk: vv for kk, vv in e2e_dense.items() if kk not in ("value", "unit")}},
        "engines": {"auto": auto_engine, "prefix": prefix_engine, "full": full_engine, "dense_model_wiring": model_wiring_dense},
        "gpu_launches": int(launches1 - launches0),
        "clocks": clocks,
        "mean_rounds": all_rounds / (n_global * args.steps),
    }

    # ---- N = 1: the remaining BASELINE configurations and the bit-exact fp64 engines, each with CUDA-event time
    if world == 1 and not args.no_configs:
        cfgs = {}
        cfgs["table_fp64_bit_exact"] = dict(timed(ENGINE_TABLE, WIRING_REF_CLI, 4000, precision=FP64, steps=min(args.steps, 3), warm=1),
                                            note="fp64 CDF tables + 53-bit uniforms: bit-identical to the oracle / the reference under shared draws "
                                                 "(tests: 100 000 trials); same workload and step size as `value`")
        small = DeviceBatch(eng, 1, 100_000, R, want_reason=True)
        cfgs["exact64_model_wiring"] = dict(timed(ENGINE_DENSE, WIRING_MODEL, 4100, precision=FP64, b=small, steps=min(args.steps, 3), warm=1),
                                            note="fp64 exact DENSE engine, `model` wiring (bandwagon + stickiness live): the parity workhorse, 100 000 trials per step")
        cfgs["exact64_model_wiring_fatigue_stop"] = dict(timed(ENGINE_DENSE, WIRING_MODEL, 4200, p=params_x, precision=FP64, b=small, steps=min(args.steps, 3), warm=1),
                                                         note="the same with candidate fatigue + stop-candidate live (what FULL replaces in fp32)")
        del small
        # cfg-4: 256 parameter sets x 100 000 trials, ONE launch with per-CTA parameters
        sets4 = [make_params(dict(CFG2, initial_beta_weight=b0, beta_increment_amount=db, papabile_weight_factor=pw), max_rounds=20,
                             supermajority_threshold=thr) for b0, db, pw, thr in cfg4_grid()]
        b4 = DeviceBatch(eng, len(sets4), 100_000, R, want_reason=False)
        cfgs["cfg4_sweep_256x100k"] = {"auto_ref_cli": timed(ENGINE_AUTO, WIRING_REF_CLI, 5000, b=b4, sets=sets4, steps=min(args.steps, 3), warm=1),
                                       "auto_model": timed(ENGINE_AUTO, WIRING_MODEL, 5100, b=b4, sets=sets4, steps=min(args.steps, 3), warm=1),
                                       "dense_ref_cli": timed(ENGINE_DENSE, WIRING_REF_CLI, 5200, b=b4, sets=sets4, steps=1, warm=1),
                                       "note": "BASELINE configs[3]: SURVEY 8d grid (beta0 x increment x papabile x threshold), 25.6 M trials per step, "
                                               "per-set histograms on the device"}
        del b4
        # cfg-5: synthetic 1024 x 1024 roster, R = 50, 1 M trials
        x5, g5, p5 = synthetic_roster(1024)
        eng.upload_roster(x5, g5, p5)
        p5r = make_params(CFG2, max_rounds=50, supermajority_threshold=0.56)
        b5 = DeviceBatch(eng, 1, 1_000_000, 50, want_reason=False)
        cfgs["cfg5_1024x1024_R50_1M"] = {"auto_ref_cli": timed(ENGINE_AUTO, WIRING_REF_CLI, 6000, p=p5r, b=b5, steps=min(args.steps, 3), warm=1),
                                         "auto_model": timed(ENGINE_AUTO, WIRING_MODEL, 6100, p=p5r, b=b5, steps=min(args.steps, 3), warm=1),
                                         "table_ref_cli": timed(ENGINE_TABLE, WIRING_REF_CLI, 6200, p=p5r, b=b5, steps=1, warm=1),
                                         "note": "BASELINE configs[4] at full size: 1 M trials of 1024 electors x 1024 candidates, 50 rounds"}
        del b5
        eng.upload_roster(ideology, region, pap)
        # the literal cfg-1 / cfg-2 sizes through the host API (too short for events to mean much: wall clock, best of 5)
        x134, g134, p134 = synthetic_roster(134)
        eng.upload_roster(x134, g134, p134)
        lit = {}
        for nm, mp_kw, thr, nn in (("cfg1_defaults_x100", CFG1, 2 / 3, 100), ("cfg2_readme_x1000", CFG2, 0.56, 1000)):
            pp = make_params(mp_kw, max_rounds=20, supermajority_threshold=thr)
            eng.run([pp], nn, seed=SEED)
            best = None
            for _ in range(5):
                t0 = time.perf_counter()
                rr_ = eng.run([pp], nn, seed=SEED)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            lit[nm] = {"wall_ms": best * 1e3, "conclaves_per_sec": nn / best, "value": rr_.elector_rounds / best, "unit": "elector-rounds/s", "engine_ran": rr_.engine}
        cfgs["literal_small_configs_host_api"] = lit
        eng.upload_roster(ideology, region, pap)
        line["configs"] = cfgs

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        probe_er, probe_c, _ = cpu_oracle_rate(2000, cores, E)
        n_cpu = max(2000, int(probe_c * 12.0))       # ~12 s of CPU work
        er, c, dt = cpu_oracle_rate(n_cpu, cores, E)
        tab_er, _, _ = cpu_oracle_rate(max(20000, int(probe_c * 40)), cores, E, faithful=False)
        # the as-shipped command line of BASELINE configs[0] (CLI defaults, 100 trials, 20 rounds), same port, same cores
        s_er, s_c, s_dt = cpu_oracle_rate(100, cores, 134, model_kw=CFG1, engine_kw=dict(max_rounds=20, supermajority_threshold=2 / 3))
        line["cpu_baseline"] = {"value": er, "unit": "elector-rounds/s", "conclaves_per_sec": c, "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} trials of the same workload, C oracle re-evaluating the E x C matrix per trial "
                                          f"per round like the reference (OpenMP, {cores} threads), {dt:.1f} s",
                                "table_optimised_value": tab_er,
                                "as_shipped_cfg1_x100": {"value": s_er, "conclaves_per_sec": s_c, "seconds": s_dt,
                                                         "what": "BASELINE configs[0] (`--num-simulations 100 --max-rounds 20`, CLI defaults) on the same port and cores"},
                                "python_reference": PYTHON_REFERENCE_NOTE}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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

One step = one launch of the fp32 DENSE trial kernel (every trial re-evaluates its E x C scores
each round) + one NCCL all-reduce of the winner / rounds histograms.  `value` is whole-job
elector-rounds/s with inputs resident in HBM; `e2e` is the same metric through the host-buffer
public API (roster + parameters uploaded, results copied back, every step).

The oracle (oracle/) is only ever executed here for the cpu_baseline leg and for --impl reference.
"""
from __future__ import annotations

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

E_HEADLINE = 135
CFG2 = dict(initial_beta_weight=1.5, enable_dynamic_beta=True, beta_increment_amount=0.3,
            beta_increment_interval_rounds=1, stickiness_factor=0.8, bandwagon_strength=0.7,
            regional_affinity_bonus=0.1, papabile_weight_factor=2.0)
ENGINE_KW = dict(max_rounds=20, supermajority_threshold=0.56, runoff_threshold_rounds=5, runoff_candidate_count=2)
SEED = 20250507
# canonical dense work per (elector, candidate) pair
FP32_PER_PAIR, MUFU_PER_PAIR = 10, 1
FALLBACK_FP32_GINSTR = 148 * 128 * 1.965     # architectural: SMs x FP32 lanes x GHz (used only if the microbench fails)
FALLBACK_MUFU_GINSTR = 148 * 16 * 1.965


def synthetic_roster(E: int, seed: int = 20250507):
    """SURVEY §8(d) recipe (kept here so the product arm needs nothing from oracle/)."""
    rng = np.random.default_rng(seed)
    ideology = rng.beta(2, 4, E)
    p = np.array([53, 23, 18, 18, 10, 5, 4, 3], dtype=np.float64)
    region = rng.choice(8, size=E, p=p / p.sum()).astype(np.int32)
    pap = rng.random(E) < 0.15
    return ideology, region, pap


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


def cpu_oracle_rate(trials: int, threads: int, E: int, faithful: bool = True):
    """Times the C oracle (oracle/conclave_oracle.c, OpenMP) on `trials` trials of the workload.
    faithful=True: the whole E x C matrix (exp included) is re-evaluated per trial per round, as the
    reference does (src/simulate.py:121, src/model.py:311-486) -- the like-for-like baseline of the
    dense GPU engine.  faithful=False: the same port with trial-invariant per-round CDF tables."""
    from oracle import conclave_oracle as O
    from oracle import c_oracle as CO
    CO.build()
    CO.set_faithful(faithful)
    roster = O.synthetic_roster(E)
    mp = O.ModelParams(**CFG2)
    ep = O.EngineParams(max_rounds=20, supermajority_threshold=0.56)
    t0 = time.perf_counter()
    out = CO.run(roster, [(mp, ep)], O.WIRING_REF_CLI, trials, seed=SEED, want_final=False, threads=threads)
    dt = time.perf_counter() - t0
    er = int(out["totals"][0, 0]) * E
    return er / dt, trials / dt, dt


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
                         "table_optimised_value": tab_er},
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
    from conclave_sim_b200 import Engine, make_params, FP32, WIRING_REF_CLI, WIRING_MODEL, ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX, ENGINE_FULL
    from conclave_sim_b200.dist import DeviceBatch, shard_range

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
    batch = DeviceBatch(eng, 1, n_local, R, want_reason=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i: int, reduce: bool = True, engine: int = ENGINE_DENSE, wiring: int = WIRING_REF_CLI, p=None):
        begin, _ = shard_range(n_global, rank, world)
        batch.run([params if p is None else p], sim_id_begin=i * n_global + begin, seed=SEED, wiring=wiring,
                  engine=engine, precision=FP32, reduce=reduce)

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

    # ---- roofline of the dominant kernel (k_trials_dense32), per launch, this rank
    try:
        peak_fp32 = eng.microbench(0); peak_mufu = eng.microbench(1); peak_mix = eng.microbench(2); peak_fact = eng.microbench(3)
        peak_src = "measured (csim_microbench on this GPU)"
    except Exception:
        peak_fp32, peak_mufu, peak_mix, peak_fact, peak_src = FALLBACK_FP32_GINSTR, FALLBACK_MUFU_GINSTR, None, None, "architectural fallback"
    pairs_per_launch = tot_pairs / args.steps
    kern_s = (sum(kernel_ms) / args.steps) * 1e-3
    ach_fp32 = pairs_per_launch * FP32_PER_PAIR / kern_s * 1e-9
    ach_mufu = pairs_per_launch * MUFU_PER_PAIR / kern_s * 1e-9
    out_bytes = n_local * 9 + (E + 1 + R + 1 + 2) * 8
    roofline = {"bound": "fp32", "kernel": "k_trials_dense32<LC=8,MODEL=false,FACT=true>", "achieved": ach_fp32, "peak": peak_fp32,
                "unit": "G thread-instr/s", "frac": ach_fp32 / peak_fp32, "traffic": 111872,
                "traffic_note": "ncu dram__bytes_read+write of one launch at 250k trials (profiles/r01_ncu_dense32_v5_summary.txt); "
                                "the kernel is compute-bound, HBM only sees the 9 B/trial result writeback",
                "algorithmic": f"{FP32_PER_PAIR} FP32-pipe instr per (elector,candidate) pair evaluation (SURVEY 8d); "
                               f"{pairs_per_launch:.4g} pair evaluations per launch (E*C in full-field rounds, E*k in run-off rounds)",
                "peak_source": peak_src, "kernel_ms_per_launch": kern_s * 1e3,
                "sfu": {"achieved": ach_mufu, "peak": peak_mufu, "unit": "G MUFU/s", "frac": ach_mufu / peak_mufu,
                        "note": "canonical 1 MUFU.EX2 per pair; the factorised kernel (FACT) issues none per pair"},
                "note": "achieved uses the CANONICAL 10 FP32 instr/pair of SURVEY 8d so that rounds are comparable; the factorised "
                        "kernel executes ~2.5 issue slots per pair in phase 1 (FMUL2/FMNMX/FADD2), see DESIGN.md 5 and "
                        "profiles/*ncu_dense32*summary.txt for executed-instruction pipe utilisation",
                "pair_evals_per_s": pairs_per_launch / kern_s,
                "mufu_pair_body_peak_pairs_per_s": None if peak_mix is None else peak_mix / 4 * 1e9,
                "factorised_pair_body_peak_pairs_per_s": None if peak_fact is None else peak_fact / 2.5 * 1e9,
                "hbm_writeback_gbs": out_bytes / kern_s * 1e-9}

    # ---- the other engines on the same workload, reported SEPARATELY from the dense kernel's value / roofline
    #      (never mixed): same steps, same L2 flush, CUDA events, max over ranks
    def timed_engine(engine: int, wiring: int, base: int, p=None):
        for i in range(2):
            step(base + i, engine=engine, wiring=wiring, p=p)
        barrier()
        tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_rounds, k_ms = 0, 0.0
        for i in range(args.steps):
            flush.zero_()
            tev[i][0].record()
            step(base + 2 + i, engine=engine, wiring=wiring, p=p)
            tev[i][1].record()
            tev[i][1].synchronize()
            k_ms += eng.stats()[1]
            t_rounds += int(batch.totals[0].item())       # already summed over ranks by the step's all-reduce
        barrier()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in tev), k_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt[0].item()
        return {"value": float(t_rounds) * E / (ms * 1e-3), "unit": "elector-rounds/s",
                "conclaves_per_sec": n_global * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                "trial_kernel_ms_per_step": tt[1].item() / args.steps, "mean_rounds": float(t_rounds) / (n_global * args.steps)}

    table_engine = timed_engine(ENGINE_TABLE, WIRING_REF_CLI, 500)
    table_engine.update(kernel="k_trials_table<float> (+ k_base_scores64, k_cdf_tables64, k_f64_to_f32 per step)",
                        note="ref_cli wiring only: the probability matrix is trial-invariant there, so CDF rows are built once per "
                             "step and a draw is a binary search; NOT comparable with the dense roofline")
    prefix_engine = {"ref_cli": timed_engine(ENGINE_PREFIX, WIRING_REF_CLI, 600),
                     "model": timed_engine(ENGINE_PREFIX, WIRING_MODEL, 700),
                     "kernel": "k_trials_prefix<LC8, MODEL, STAGE> (+ k_prefix_tables32 per step)",
                     "note": "trial-invariant chunk-prefix tables + per-trial bandwagon / stickiness corrections: a draw is a binary "
                             "search over 17 chunk prefixes and a re-scan of one chunk of 8 candidates; both wirings; what "
                             "ENGINE_AUTO picks for the `model` wiring.  NOT comparable with the dense roofline"}
    params_x = make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), enable_candidate_fatigue=True,
                           enable_stop_candidate=True, **ENGINE_KW)
    full_engine = timed_engine(ENGINE_FULL, WIRING_MODEL, 900, p=params_x)
    full_pairs = pair_evals(batch.rounds.cpu().numpy(), E, rr, k)             # this rank's last step (every step is the same size)
    full_kern_s = full_engine["trial_kernel_ms_per_step"] * 1e-3
    full_engine.update(kernel="k_trials_full32",
                       note="`model` wiring with EVERY term live: cfg-2 parameters + candidate fatigue (3 rounds below 5 %) + stop-candidate "
                            "(unacceptable distance 0.3): all E x C scores inline per trial per round; the engine ENGINE_AUTO picks for what the "
                            "table-based engines refuse (these terms used to run on the fp64 exact kernel only, 3.5e5 conclaves/s)")
    full_engine["roofline"] = {"bound": "fp32", "kernel": "k_trials_full32<EXT=true,NEG=false>", "unit": "G thread-instr/s",
                               "achieved": full_pairs * FP32_PER_PAIR / full_kern_s * 1e-9, "peak": peak_fp32,
                               "frac": full_pairs * FP32_PER_PAIR / full_kern_s * 1e-9 / peak_fp32,
                               "sfu_frac": full_pairs * MUFU_PER_PAIR / full_kern_s * 1e-9 / peak_mufu,
                               "note": "the same canonical accounting as the dense roofline (10 FP32 + 1 MUFU per pair evaluation); this "
                                       "kernel is the MUFU form with every term inline, so the fraction is what the canonical formulation "
                                       "itself reaches here — the factorised dense kernel's fraction above is the optimised form of the same work"}
    # what bounds those kernels, from the committed ncu captures (static evidence, not measured by this run)
    table_engine["ncu"] = {"issue_slots_busy_pct_while_resident": 67, "bound": "shared-memory latency (serial 8-probe searches)",
                           "source": "profiles/r01_ncu_table_r1b_summary.txt"}
    prefix_engine["ncu"] = {"issue_slots_busy_pct": 76, "shared_wavefronts_pct_of_peak": 58, "bound": "issue slots, then shared-memory wavefronts",
                            "source": "profiles/r01_ncu_prefix_model_final_summary.txt"}
    full_engine["ncu"] = {"issue_slots_busy_pct": 54, "top_stall": "no_instruction", "source": "profiles/r01_ncu_full32_summary.txt"}
    model_wiring_dense = timed_engine(ENGINE_DENSE, WIRING_MODEL, 800)
    model_wiring_dense.update(kernel="k_trials_dense32<LC=8,MODEL=true,FACT=true>",
                              note="`model` wiring (bandwagon + stickiness live, SURVEY 8f.1) on the dense kernel, for comparison")

    # ---- e2e: the host-buffer public API (roster + params uploaded, results copied back, every step)
    e2e_steps = max(2, min(args.steps, 3))
    begin, _ = shard_range(n_global, rank, world)
    hist_dev = torch.zeros((E + 1) + (R + 1) + 2, dtype=torch.int64, device=dev)
    pinned = {"winner": torch.empty(n_local, dtype=torch.int32).pin_memory().numpy(),
              "rounds": torch.empty(n_local, dtype=torch.int32).pin_memory().numpy(),
              "reason": torch.empty(n_local, dtype=torch.uint8).pin_memory().numpy()}
    def e2e_step(i):
        eng._roster_key = None                      # force the roster upload inside the timed region
        eng.upload_roster(ideology, region, pap)
        r = eng.run([params], n_local, sim_id_begin=(1000 + i) * n_global + begin, seed=SEED, wiring=WIRING_REF_CLI,
                    engine=ENGINE_DENSE, precision=FP32, want_reason=True, out=pinned)
        if world > 1:
            h = torch.from_numpy(np.concatenate([r.winner_hist.ravel(), r.rounds_hist.ravel(), r.totals.ravel()])).to(dev)
            dist.all_reduce(h)
            h.cpu()
        return r
    e2e_step(-1)
    barrier()
    t0 = time.perf_counter()
    e2e_rounds = 0
    for i in range(e2e_steps):
        r = e2e_step(i)
        e2e_rounds += int(r.totals[0, 0])
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_rounds)], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rounds = a[0].item(), b[1].item()
    e2e_value = e2e_rounds * E / e2e_s
    # the same through the reference-facing Python call (DataFrame roster, default engine = table for ref_cli)
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    df = pd.DataFrame({"ideology_score": ideology, "region": region, "is_papabile": pap},
                      index=pd.Index([str(i) for i in range(E)], name="elector_id"))
    api_kw = dict(max_rounds=R, supermajority_threshold=ENGINE_KW["supermajority_threshold"], runoff_threshold_rounds=rr,
                  runoff_candidate_count=k)

    def api_e2e(wiring: str, base: int):
        """run_simulations(DataFrame, ...) with the default engine, roster re-uploaded every step, results on the host"""
        S.RUNTIME.update(seed=SEED, device=local_rank, precision="fp32", engine="auto", wiring=wiring)
        warm = [S.run_simulations(df, CFG2, n_local, sim_id_begin=begin + j, **api_kw) for j in range(2)]   # two live results: the
        del warm                                                  # page-locked pool reaches its steady-state size before the clock starts
        barrier()
        t0 = time.perf_counter()
        rounds_sum = 0
        for i in range(e2e_steps):
            S.get_engine(local_rank)._roster_key = None
            r = S.run_simulations(df, CFG2, n_local, sim_id_begin=(base + i) * n_global + begin, **api_kw)
            rounds_sum += int(r.totals[0, 0])
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, float(rounds_sum)], dtype=torch.float64, device=dev)
        if world > 1:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            dt, rounds_sum = a[0].item(), b[1].item()
        return {"value": rounds_sum * E / dt, "unit": "elector-rounds/s", "conclaves_per_sec": n_global * e2e_steps / dt,
                "api": f"conclave_sim_b200.simulate.run_simulations(DataFrame, ..., wiring='{wiring}') with the default engine"}

    table_engine["e2e_default_api"] = api_e2e("ref_cli", 2000)
    prefix_engine["model"]["e2e_default_api"] = api_e2e("model", 3000)
    S.RUNTIME.update(wiring="ref_cli")
    h2d = E * (8 + 4 + 1) + 152 + 20 * 12
    d2h = n_local * 9 + ((E + 1) + (R + 1) + 2) * 8

    line = {
        "metric": "elector_rounds_per_sec", "value": value, "unit": "elector-rounds/s",
        "conclaves_per_sec": conclaves, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "rank_ms_per_step": rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2 README params x {n_local} trials/GPU/step (cfg3 per-GPU share), synthetic E={E}, "
                               "ref_cli wiring, dense fp32 engine, Philox4x32-10",
                   "trials_per_step_global": n_global, "max_rounds": R, "l2": "flushed between steps (256 MiB memset, untimed)",
                   "collective": "1 NCCL all-reduce (int64 sum) of winner/rounds histograms per step" if world > 1 else "none (1 GPU)"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "elector-rounds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "conclave_sim_b200.Engine.run (host buffers; results land in page-locked arrays) -> csim_run"},
        "table_engine": table_engine,
        "prefix_engine": prefix_engine,
        "full_engine": full_engine,
        "model_wiring_dense": model_wiring_dense,
        "gpu_launches": int(launches1 - launches0),
        "clocks": clocks,
        "mean_rounds": all_rounds / (n_global * args.steps),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        probe_er, probe_c, _ = cpu_oracle_rate(2000, cores, E)
        n_cpu = max(2000, int(probe_c * 12.0))       # ~12 s of CPU work
        er, c, dt = cpu_oracle_rate(n_cpu, cores, E)
        tab_er, _, _ = cpu_oracle_rate(max(20000, int(probe_c * 40)), cores, E, faithful=False)
        line["cpu_baseline"] = {"value": er, "unit": "elector-rounds/s", "conclaves_per_sec": c, "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} trials of the same workload, C oracle re-evaluating the E x C matrix per trial "
                                          f"per round like the reference (OpenMP, {cores} threads), {dt:.1f} s",
                                "table_optimised_value": tab_er,
                                "python_reference_in_build_container": "38.2 conclaves/s, 58.7k elector-rounds/s on 8 cores "
                                                                       "(unmodified reference, 20000 trials; tests/golden/dist_cfg2.json)"}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

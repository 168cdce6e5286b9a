"""Drop-in for `python -m src.simulate` (mirror of src/simulate.py) with the trials on the GPU.

Unchanged surface: the CLI flags (names, short forms, types, defaults — src/simulate.py:220-259),
`model_params_from_args` (:261-295), `_run_single_simulation_worker` (:30-218, same arguments and
result dict), the results CSV (`sim_id,winner,rounds,reason`, :395-412) and the stats JSON
(:374-387,423) that viz.py / generate_visualizations.py consume.

What replaces the reference's ProcessPoolExecutor fan-out (:330-351): ONE batched call into the
CUDA library for all trials (`run_simulations`), sharded over the GPUs of the box by disjoint sim-id
ranges when more than one device is asked for (`--devices N`, or the reference's own `-w/--workers`
knob read as a device count; also under torchrun, one process per GPU): the per-device histograms are
combined by ONE NCCL all-reduce and the per-trial rows are concatenated on the host, so the CSV / JSON
are the same bytes whatever the device count.  New, opt-in flags only: --wiring, --seed, --device,
--devices, --precision, --engine, --tape.

Wiring: `ref_cli` (default) reproduces what the reference's CLI computes — the
previous round's votes never reach the model, so bandwagon / stickiness / fatigue /
stop-candidate are inert; `model` connects them as the model's own signature documents.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time
from collections import Counter
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from . import _cabi
from .engine import Engine, EngineGroup, get_engine, get_group, device_count, make_params, RunResult
from .ingest import ElectorDataIngester, roster_arrays

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
log = logging.getLogger(__name__)
worker_log = logging.getLogger(f"{__name__}.worker")

DEFAULT_NUM_SIMULATIONS = 100
DEFAULT_MAX_ROUNDS = 100
DEFAULT_SUPERMAJORITY_THRESHOLD = 2 / 3
DEFAULT_RUNOFF_THRESHOLD_ROUNDS = 5
DEFAULT_RUNOFF_CANDIDATE_COUNT = 2

# Process-wide defaults of the opt-in knobs (the reference draws from an unseeded global RNG, so
# every process sees a different stream; we draw a fresh Philox key per process unless told).
RUNTIME: Dict[str, Any] = dict(wiring="ref_cli", seed=None, device=0, devices=None, precision="fp32", engine="auto")
_WIRING = {"ref_cli": _cabi.WIRING_REF_CLI, "model": _cabi.WIRING_MODEL}
_PRECISION = {"fp32": _cabi.FP32, "fp64": _cabi.FP64}
_ENGINE = {"auto": _cabi.ENGINE_AUTO, "dense": _cabi.ENGINE_DENSE, "table": _cabi.ENGINE_TABLE, "prefix": _cabi.ENGINE_PREFIX, "full": _cabi.ENGINE_FULL}


def _seed() -> int:
    if RUNTIME["seed"] is None:
        RUNTIME["seed"] = int.from_bytes(os.urandom(8), "little")
    return int(RUNTIME["seed"])


def resolve_devices(devices=None, device: Optional[int] = None, workers: Optional[int] = None) -> tuple:
    """The CUDA devices one call shards its trials over.
    `devices`: None, a count (int or digit string: the first N visible devices), "all", a comma list "0,2,3" or a sequence of
    ordinals.  With `devices` None the reference's own parallelism knob applies: `workers` (-w/--workers, src/simulate.py:231)
    is read as a device count, clamped to what the box has; with neither, the single `device`."""
    first = RUNTIME["device"] if device is None else int(device)
    if devices is None:
        if workers is None or int(workers) <= 1:
            return (first,)
        n = max(1, min(int(workers), device_count()))
        return tuple(range(n)) if n > 1 else (first,)
    if isinstance(devices, str):
        d = devices.strip().lower()
        if d == "all":
            return tuple(range(max(1, device_count())))
        if "," in d:
            devices = [int(v) for v in d.split(",") if v.strip() != ""]
        else:
            devices = int(d)
    if isinstance(devices, (int, np.integer)):
        n = int(devices)
        if n < 1:
            raise ValueError("devices must be >= 1")
        have = device_count()
        if n > have:
            raise ValueError(f"{n} devices requested, {have} visible")
        return tuple(range(n))
    out = tuple(int(v) for v in devices)
    if not out or len(set(out)) != len(out):
        raise ValueError(f"bad device list {devices!r}")
    return out


def _torchrun_world():
    """(rank, world, local_rank) when this process is one rank of an initialised torch.distributed job, else None."""
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))


def reason_text(code: int, supermajority_threshold: float) -> str:
    """The three strings of src/simulate.py:61,130,178."""
    if code == _cabi.REASON_WINNER:
        return f"Winner found by supermajority ({supermajority_threshold*100:.0f}% threshold)"
    if code == _cabi.REASON_NO_PROBS:
        return "No probabilities generated"
    return "Max rounds reached"


def run_simulations(elector_data_full: pd.DataFrame, model_params: Dict[str, Any], num_simulations: int, *,
                    max_rounds: int, supermajority_threshold: float, runoff_threshold_rounds: int,
                    runoff_candidate_count: int, enable_candidate_fatigue: bool = False,
                    fatigue_threshold_rounds: int = 3, fatigue_vote_share_threshold: float = 0.05,
                    fatigue_top_n_immune: int = 3, enable_stop_candidate: bool = False,
                    sim_id_begin: int = 0, seed: Optional[int] = None, wiring: Optional[str] = None,
                    device: Optional[int] = None, precision: Optional[str] = None, engine: Optional[str] = None,
                    want_final_votes: bool = False, want_round_tallies: bool = False,
                    uniform_tapes: Optional[np.ndarray] = None, devices=None, workers: Optional[int] = None,
                    distributed: Optional[bool] = None, gather_rows: bool = True) -> RunResult:
    """All trials of one parameter set in one batch (replaces src/simulate.py:330-351).

    One device: one launch.  `devices` (count / list / "all", or `workers` read as a device count — resolve_devices): the trials are
    sharded over the devices by disjoint sim-id ranges from this ONE process (csim_group_run: one NCCL all-reduce of the
    histograms, rows concatenated on the host).  Inside an initialised torch.distributed job (torchrun, one process per GPU;
    `distributed` None = detect): this rank runs its shard on its LOCAL_RANK device, the histograms are all-reduced over NCCL
    through the C ABI (csim_allreduce_hists) and rank 0 receives every rank's rows (gather_rows=False: every rank keeps only its
    own rows); the other ranks get their own rows and the global histograms.  Whatever the split, the results equal a
    single-device run with the same seed."""
    wiring = wiring or RUNTIME["wiring"]
    precision = precision or RUNTIME["precision"]
    engine = engine or RUNTIME["engine"]
    tr = _torchrun_world() if distributed in (None, True) else None
    if distributed is True and tr is None:
        raise RuntimeError("distributed=True needs an initialised torch.distributed process group with more than one rank")
    if tr is not None and devices is None and RUNTIME["devices"] is None:
        return _run_simulations_torchrun(tr, elector_data_full, model_params, num_simulations, max_rounds=max_rounds,
                                         supermajority_threshold=supermajority_threshold, runoff_threshold_rounds=runoff_threshold_rounds,
                                         runoff_candidate_count=runoff_candidate_count, enable_candidate_fatigue=enable_candidate_fatigue,
                                         fatigue_threshold_rounds=fatigue_threshold_rounds,
                                         fatigue_vote_share_threshold=fatigue_vote_share_threshold, fatigue_top_n_immune=fatigue_top_n_immune,
                                         enable_stop_candidate=enable_stop_candidate, sim_id_begin=sim_id_begin, seed=seed, wiring=wiring,
                                         precision=precision, engine=engine, gather_rows=gather_rows)
    devs = resolve_devices(RUNTIME["devices"] if devices is None else devices, device, workers)
    eng = get_group(devs) if len(devs) > 1 else get_engine(devs[0])
    x, g, pap = roster_arrays(elector_data_full)
    eng.upload_roster(x, g, pap)
    params = make_params(model_params, max_rounds=max_rounds, supermajority_threshold=supermajority_threshold,
                         runoff_threshold_rounds=runoff_threshold_rounds, runoff_candidate_count=runoff_candidate_count,
                         enable_candidate_fatigue=enable_candidate_fatigue,
                         fatigue_threshold_rounds=fatigue_threshold_rounds,
                         fatigue_vote_share_threshold=fatigue_vote_share_threshold,
                         fatigue_top_n_immune=fatigue_top_n_immune, enable_stop_candidate=enable_stop_candidate)
    prec = _PRECISION[precision]
    if uniform_tapes is not None:
        prec = _cabi.FP64
    return eng.run([params], num_simulations, sim_id_begin=sim_id_begin, seed=_seed() if seed is None else seed,
                   wiring=_WIRING[wiring], engine=_ENGINE[engine], precision=prec, tapes=uniform_tapes,
                   want_final_votes=want_final_votes, want_round_tallies=want_round_tallies)


def load_tapes(path: str, num_simulations: int, max_rounds: int, n_electors: int) -> np.ndarray:
    """--tape FILE: uniform draws exported from the reference (`.npy`, or `.npz` with an array named `tapes`), one row of
    max_rounds * E uniforms per trial in elector order per round — e.g. np.random.seed(s); np.random.random_sample(R * E) per
    trial, the reference's own MT19937 stream.  The fp64 exact engine then reproduces the reference bit for bit."""
    a = np.load(path)
    if hasattr(a, "files"):
        a = a["tapes"]
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    want = (int(num_simulations), int(max_rounds) * int(n_electors))
    if a.shape[0] < want[0] or a.shape[1] != want[1]:
        raise ValueError(f"--tape: {path} has shape {a.shape}, need at least {want[0]} rows of {want[1]} uniforms "
                         f"(max_rounds {max_rounds} x {n_electors} electors)")
    return a[:want[0]]


_torchrun_state: Dict[Any, Any] = {}


def _run_simulations_torchrun(tr, elector_data_full, model_params, num_simulations, *, sim_id_begin, seed, wiring, precision, engine,
                              gather_rows: bool = True, **engine_kw) -> RunResult:
    """One rank of a torchrun job: its contiguous shard of the sim ids on its LOCAL_RANK device, ONE NCCL all-reduce of the
    histograms through the C ABI, and (gather_rows) every rank's per-trial rows gathered over NVLink on rank 0 — the rank that
    writes the CSV — which copies them to page-locked host arrays in one pass.  Device buffers are kept between calls."""
    import torch
    import torch.distributed as tdist
    from .dist import DeviceBatch, shard_range, csim_comm
    rank, world, local_rank = tr
    eng = get_engine(local_rank)
    x, g, pap = roster_arrays(elector_data_full)
    eng.upload_roster(x, g, pap)
    params = make_params(model_params, **engine_kw)
    if seed is None:                        # every rank must draw from the SAME Philox key: rank 0's
        box = [_seed() if rank == 0 else None]
        tdist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    n = int(num_simulations)
    begin, count = shard_range(n, rank, world)
    torch.cuda.set_device(local_rank)
    R = int(params.max_rounds)
    mx = shard_range(n, 0, world)[1]                                   # the largest shard (ranks differ by at most one trial)
    key = (local_rank, n, world, R)
    st = _torchrun_state.get(key)
    if st is None:
        _torchrun_state.clear()                                         # one job shape at a time: release the previous buffers
        batch = DeviceBatch(eng, 1, mx, R, want_reason=True, comm=csim_comm(eng))     # padded to the largest shard: equal-size gathers
        dev = batch.winner.device
        recv = None
        if rank == 0 and gather_rows:
            recv = {"winner": torch.empty(world * mx, dtype=torch.int32, device=dev), "rounds": torch.empty(world * mx, dtype=torch.int32, device=dev),
                    "reason": torch.empty(world * mx, dtype=torch.uint8, device=dev)}
        st = _torchrun_state[key] = (batch, recv)
    batch, recv = st
    batch.n_sims_local = count                                          # this rank plays `count` of the mx slots
    batch.run([params], sim_id_begin=int(sim_id_begin) + begin, seed=seed, wiring=_WIRING[wiring], engine=_ENGINE[engine],
              precision=_PRECISION[precision], job_sims=n)
    batch.n_sims_local = mx
    wh, rh, tot = batch.hists()
    name, prec, fell = eng.last_run_info(requested_precision=_PRECISION[precision]) if count else ("", "", False)
    counts = [shard_range(n, r, world)[1] for r in range(world)]
    out = {}
    if gather_rows:
        for nm, t in (("winner", batch.winner), ("rounds", batch.rounds), ("reason", batch.reason)):
            if recv is not None and rank == 0:
                tdist.gather(t, [recv[nm][r * mx:(r + 1) * mx] for r in range(world)], dst=0)
            else:
                tdist.gather(t, None, dst=0)
    if rank == 0 and gather_rows:
        for nm, dt in (("winner", np.int32), ("rounds", np.int32), ("reason", np.uint8)):
            host = eng._pinned.array(n, dt)                             # page-locked: the copies below run at full PCIe speed
            ht = torch.from_numpy(host) if n else None
            pos = 0
            if n and all(c == mx for c in counts):
                ht.copy_(recv[nm][:n])
            else:
                for r, c in enumerate(counts):
                    if c:
                        ht[pos:pos + c].copy_(recv[nm][r * mx:r * mx + c])
                    pos += c
            out[nm] = host
    else:                                                               # this rank's own rows
        for nm, t, dt in (("winner", batch.winner, np.int32), ("rounds", batch.rounds, np.int32), ("reason", batch.reason, np.uint8)):
            host = eng._pinned.array(count, dt)
            if count:
                torch.from_numpy(host).copy_(t[:count])
            out[nm] = host
    ms = eng.stats()[1]
    return RunResult(out["winner"], out["rounds"], out["reason"], None, None, wh.copy(), rh.copy(), tot.copy(), eng.n_electors, ms, name, prec, fell,
                     (local_rank,), (ms,))


def _run_single_simulation_worker(sim_id: int, elector_data_full: pd.DataFrame, model_params: Dict[str, Any],
                                  max_rounds: int, supermajority_threshold: float,
                                  runoff_threshold_rounds: int, runoff_candidate_count: int,
                                  verbose: bool,
                                  enable_candidate_fatigue: bool,
                                  fatigue_threshold_rounds: int,
                                  fatigue_vote_share_threshold: float,
                                  fatigue_top_n_immune: int,
                                  enable_stop_candidate: bool) -> Dict[str, Any]:
    """One trial; same arguments and result dict as src/simulate.py:30-218.  The trial's random
    stream is the Philox stream of (process seed, sim_id)."""
    worker_log.setLevel(logging.DEBUG if verbose else logging.INFO)
    candidate_ids = list(elector_data_full.index) if not elector_data_full.empty else []
    if len(candidate_ids) == 0 or "ideology_score" not in elector_data_full or "region" not in elector_data_full:
        # empty model -> probabilities.size == 0 in round 1 (src/simulate.py:128-131)
        rounds = 1 if max_rounds >= 1 else 0
        reason = "No probabilities generated" if max_rounds >= 1 else "Max rounds reached"
        if max_rounds >= 1:
            worker_log.warning(f"[Sim {sim_id}] No probabilities returned. Ending simulation.")
        return {"sim_id": sim_id, "winner": None, "rounds": rounds, "reason": reason, "final_votes": {}, "run_details": []}
    res = run_simulations(elector_data_full, model_params, 1, max_rounds=max_rounds,
                          supermajority_threshold=supermajority_threshold,
                          runoff_threshold_rounds=runoff_threshold_rounds, runoff_candidate_count=runoff_candidate_count,
                          enable_candidate_fatigue=enable_candidate_fatigue,
                          fatigue_threshold_rounds=fatigue_threshold_rounds,
                          fatigue_vote_share_threshold=fatigue_vote_share_threshold,
                          fatigue_top_n_immune=fatigue_top_n_immune, enable_stop_candidate=enable_stop_candidate,
                          sim_id_begin=int(sim_id), want_final_votes=True)
    w = int(res.winner[0])
    winner = str(candidate_ids[w]) if w >= 0 else None
    rounds = int(res.rounds[0])
    reason = reason_text(int(res.reason[0]), supermajority_threshold)
    if winner is not None:
        worker_log.info(f"[Sim {sim_id}] {reason} in round {rounds}: {winner}")
    elif rounds == max_rounds:
        worker_log.info(f"[Sim {sim_id}] Max rounds ({max_rounds}) reached. No winner decided through threshold.")
    final_votes = {cid: int(v) for cid, v in zip(candidate_ids, res.final_votes[0])} if max_rounds >= 1 else {}
    return {"sim_id": sim_id, "winner": winner, "rounds": rounds, "reason": reason,
            "final_votes": final_votes, "run_details": []}


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Run Conclave simulations.")
    parser.add_argument("-n", "--num-simulations", type=int, default=DEFAULT_NUM_SIMULATIONS, help="Number of simulations to run.")
    parser.add_argument("-r", "--max-rounds", type=int, default=DEFAULT_MAX_ROUNDS, help="Maximum number of voting rounds per simulation.")
    parser.add_argument("-t", "--supermajority-threshold", type=float, default=DEFAULT_SUPERMAJORITY_THRESHOLD, help="Supermajority threshold for winning.")
    parser.add_argument("--runoff-rounds", type=int, default=DEFAULT_RUNOFF_THRESHOLD_ROUNDS, help="Rounds after which to trigger a runoff if no winner.")
    parser.add_argument("--runoff-candidates", type=int, default=DEFAULT_RUNOFF_CANDIDATE_COUNT, help="Number of top candidates in a runoff.")
    parser.add_argument("-f", "--elector-data-file", type=str, default="merged_electors.csv", help="Path to the elector data CSV file.")
    parser.add_argument("-o", "--output-results", type=str, default="data/simulation_results.csv", help="Path to save detailed simulation results.")
    parser.add_argument("-s", "--output-stats", type=str, default="data/simulation_stats.json", help="Path to save aggregate simulation statistics.")
    parser.add_argument("-v", "--verbose", action="store_true", help="Enable verbose logging for simulation workers.")
    parser.add_argument("-w", "--workers", type=int, default=None, help="Number of worker processes. Here: the number of GPUs the trials are sharded over (clamped to the visible devices; --devices overrides).")

    model_group = parser.add_argument_group("Transition Model Parameters")
    model_group.add_argument("--initial-beta-weight", type=float, default=1.0, help="Initial weight for ideological distance sensitivity.")
    model_group.add_argument("--beta-increment-amount", type=float, default=0.05, help="Increment to beta weight per N scaled rounds (for Dynamic Beta).")
    model_group.add_argument("--beta-increment-interval-rounds", type=float, default=10.0, help="Number of rounds over which beta increments fully (for Dynamic Beta).")
    model_group.add_argument("--enable-dynamic-beta", action="store_true", help="Enable dynamic beta adjustment.")
    model_group.add_argument("--stickiness-factor", type=float, default=0.5, help="Factor for elector stickiness to previous vote.")
    model_group.add_argument("--bandwagon-strength", type=float, default=0.0, help="Strength of the bandwagon effect.")
    model_group.add_argument("--regional-bonus", type=float, default=0.1, help="Bonus for candidates from the same region.")
    model_group.add_argument("--papabile-weight-factor", type=float, default=1.5, help="Multiplicative weight for papabile candidates.")

    fatigue_group = parser.add_argument_group("Candidate Fatigue Parameters")
    fatigue_group.add_argument("--enable-candidate-fatigue", action="store_true", help="Enable candidate fatigue mechanism.")
    fatigue_group.add_argument("--fatigue-threshold-rounds", type=int, default=3, help="Number of recent rounds to consider for fatigue.")
    fatigue_group.add_argument("--fatigue-vote-share-threshold", type=float, default=0.05, help="Vote share below which a candidate is considered for fatigue.")
    fatigue_group.add_argument("--fatigue-penalty-factor", type=float, default=0.5, help="Factor to penalize preference scores of fatigued candidates.")
    fatigue_group.add_argument("--fatigue-top-n-immune", type=int, default=3, help="Top N candidates (by recent vote share) immune to fatigue.")

    stop_group = parser.add_argument_group("Stop Candidate Parameters")
    stop_group.add_argument("--enable-stop-candidate", action="store_true", help="Enable stop candidate behavior.")
    stop_group.add_argument("--stop-candidate-threshold-unacceptable-distance", type=float, default=0.5, help="Ideological distance beyond which a candidate is 'unacceptable'.")
    stop_group.add_argument("--stop-candidate-threat-min-vote-share", type=float, default=0.15, help="Vote share a candidate needs to be a 'threat'.")
    stop_group.add_argument("--stop-candidate-boost-factor", type=float, default=1.5, help="Factor to boost preference for a strategic 'blocker' candidate.")

    gpu = parser.add_argument_group("B200 engine options (additions; defaults reproduce the reference CLI)")
    gpu.add_argument("--wiring", choices=["ref_cli", "model"], default="ref_cli", help="ref_cli: what the reference CLI computes; model: bandwagon/stickiness/fatigue/stop-candidate live.")
    gpu.add_argument("--seed", type=int, default=None, help="Philox key (default: fresh per process, like the reference's unseeded RNG).")
    gpu.add_argument("--device", type=int, default=0, help="CUDA device ordinal.")
    gpu.add_argument("--devices", type=str, default=None, help="GPUs to shard the trials over: a count (8), a list (0,1,2,3) or 'all'. One NCCL all-reduce of the histograms; same CSV/JSON as one device.")
    gpu.add_argument("--tape", type=str, default=None, help=".npy/.npz of uniform draws exported from the reference, [num_simulations, max_rounds*E]; forces the fp64 exact engine (bit-exact replay).")
    gpu.add_argument("--precision", choices=["fp32", "fp64"], default="fp32", help="fp32 fast path or fp64 exact path.")
    gpu.add_argument("--engine", choices=["auto", "dense", "table", "prefix", "full"], default="auto")
    return parser


def parse_args(argv: Optional[List[str]] = None):
    return build_parser().parse_args(argv)


def model_params_from_args(args: argparse.Namespace) -> Dict[str, Any]:
    """The 12 TransitionModel kwargs the reference CLI forwards (src/simulate.py:263-277)."""
    params = {
        "initial_beta_weight": args.initial_beta_weight,
        "enable_dynamic_beta": args.enable_dynamic_beta,
        "beta_increment_amount": args.beta_increment_amount,
        "beta_increment_interval_rounds": int(args.beta_increment_interval_rounds),
        "stickiness_factor": args.stickiness_factor,
        "bandwagon_strength": args.bandwagon_strength,
        "regional_affinity_bonus": args.regional_bonus,
        "papabile_weight_factor": args.papabile_weight_factor,
        "fatigue_penalty_factor": args.fatigue_penalty_factor,
        "stop_candidate_threshold_unacceptable_distance": args.stop_candidate_threshold_unacceptable_distance,
        "stop_candidate_boost_factor": args.stop_candidate_boost_factor,
        "stop_candidate_threat_min_vote_share": args.stop_candidate_threat_min_vote_share,
    }
    if args.enable_dynamic_beta:
        if args.beta_increment_amount is not None and args.beta_increment_amount != 0:
            log.info(f"Dynamic Beta Enabled (stepped): Increment Amount={args.beta_increment_amount}, Interval Rounds={int(args.beta_increment_interval_rounds)}.")
        else:
            log.info(f"Dynamic Beta Enabled (likely multiplicative growth): beta_increment_amount is {args.beta_increment_amount}. TransitionModel's default beta_growth_rate will apply if > 1.0.")
    else:
        log.info("Dynamic Beta Disabled.")
    if args.enable_candidate_fatigue:
        log.info(f"Candidate Fatigue Enabled: Threshold Rounds={args.fatigue_threshold_rounds}, Vote Share Threshold={args.fatigue_vote_share_threshold*100:.0f}%, Penalty={args.fatigue_penalty_factor}, Top N Immune={args.fatigue_top_n_immune}")
    if args.enable_stop_candidate:
        log.info(f"Stop Candidate Enabled: Unacceptable Distance={args.stop_candidate_threshold_unacceptable_distance}, "
                 f"Threat Share={args.stop_candidate_threat_min_vote_share*100:.0f}%, "
                 f"Boost Factor={args.stop_candidate_boost_factor}")
    return params


def aggregate_results(results: List[Dict[str, Any]], num_simulations: int) -> Dict[str, Any]:
    """src/simulate.py:356-387 — same keys, same expressions (so the JSON text is identical)."""
    successful = sum(1 for r in results if r["winner"] is not None)
    rounds_ok = [r["rounds"] for r in results if r["winner"] is not None]
    winner_counts = Counter(r["winner"] for r in results if r["winner"] is not None)
    top = winner_counts.most_common(1)
    return {
        "total_simulations": num_simulations,
        "successful_simulations": successful,
        "success_rate": successful / num_simulations if num_simulations > 0 else 0,
        "average_rounds_successful": np.mean(rounds_ok) if rounds_ok else None,
        "min_rounds_successful": np.min(rounds_ok) if rounds_ok else None,
        "max_rounds_successful": np.max(rounds_ok) if rounds_ok else None,
        "median_rounds_successful": np.median(rounds_ok) if rounds_ok else None,
        "winner_counts": dict(winner_counts),
        "most_frequent_winner": top[0][0] if top else None,
        "most_frequent_winner_count": top[0][1] if top else 0,
        "most_frequent_winner_percentage": (top[0][1] if top else 0) / num_simulations if num_simulations > 0 else 0,
        "average_rounds_all": np.mean([r["rounds"] for r in results]) if results else 0,
    }


def results_from_batch(res: RunResult, candidate_ids: List[Any], supermajority_threshold: float,
                       sim_id_begin: int = 0) -> List[Dict[str, Any]]:
    out = []
    for i in range(len(res.winner)):
        w = int(res.winner[i])
        out.append({"sim_id": sim_id_begin + i, "winner": str(candidate_ids[w]) if w >= 0 else None,
                    "rounds": int(res.rounds[i]),
                    "reason": reason_text(int(res.reason[i]) if res.reason is not None else
                                          (_cabi.REASON_WINNER if w >= 0 else _cabi.REASON_MAX_ROUNDS), supermajority_threshold)})
    return out


def results_frame(res: RunResult, candidate_ids: List[Any], supermajority_threshold: float, sim_id_begin: int = 0) -> pd.DataFrame:
    """The results table of src/simulate.py:395-412 (columns sim_id, winner, rounds, reason; winner None -> empty cell) built
    column-wise from the batch arrays: the same rows results_from_batch yields one dict at a time, without a Python loop over
    trials (a million trials: ~0.1 s instead of seconds)."""
    n = len(res.winner)
    if n == 0:
        return pd.DataFrame([])                                       # what the reference builds from an empty result list (:412): no columns
    w = np.asarray(res.winner)
    ids = np.empty(len(candidate_ids) + 1, dtype=object)
    ids[:-1] = [str(c) for c in candidate_ids]
    ids[-1] = None                                                    # index -1 = no winner
    code = np.asarray(res.reason) if res.reason is not None else np.where(w >= 0, _cabi.REASON_WINNER, _cabi.REASON_MAX_ROUNDS)
    texts = np.empty(3, dtype=object)
    for c in (_cabi.REASON_MAX_ROUNDS, _cabi.REASON_WINNER, _cabi.REASON_NO_PROBS):
        texts[c] = reason_text(c, supermajority_threshold)
    return pd.DataFrame({"sim_id": np.arange(sim_id_begin, sim_id_begin + n, dtype=np.int64), "winner": ids[w],
                         "rounds": np.asarray(res.rounds).astype(np.int64), "reason": texts[code]})


def _write_csv(results_df: pd.DataFrame, path: str) -> None:
    """pandas' writer; for large tables pyarrow's (same text: header, minimal quoting, empty cell for a missing winner)."""
    if len(results_df) >= 200_000:
        try:
            import pyarrow as pa
            import pyarrow.csv as pacsv
            table = pa.Table.from_pandas(results_df, preserve_index=False)
            with open(path, "wb") as f:                               # pyarrow quotes header names whatever the style: write it here
                f.write((",".join(results_df.columns) + "\n").encode())
                pacsv.write_csv(table, f, write_options=pacsv.WriteOptions(include_header=False, quoting_style="none"))
            return
        except Exception:                                             # pyarrow missing or too old: pandas does the same, slower
            pass
    results_df.to_csv(path, index=False)


def _json_default(x):
    if isinstance(x, np.integer):
        return int(x)
    if x is None or (isinstance(x, float) and np.isnan(x)):
        return None
    return x


def write_outputs(results, aggregate_stats: Dict[str, Any], output_results: str, output_stats: str):
    """src/simulate.py:395-426.  `results` = the list of per-trial dicts of the reference, or the results_frame of a batch."""
    if isinstance(results, pd.DataFrame):
        results_df = results
    else:
        rows = [{"sim_id": r["sim_id"], "winner": r["winner"], "rounds": r["rounds"], "reason": r["reason"]} for r in results]
        results_df = pd.DataFrame(rows)
    try:
        _write_csv(results_df, output_results)
        log.info(f"Saving simulation results ({len(results_df)} rows) to: {output_results}")
    except Exception as exc:
        log.error(f"Failed to save results CSV: {exc}")
    stats_path = Path(output_stats)
    stats_path.parent.mkdir(parents=True, exist_ok=True)
    try:
        with open(stats_path, "w") as f:
            json.dump(aggregate_stats, f, indent=4, default=_json_default)
        log.info(f"Saving aggregate stats to: {stats_path}")
    except Exception as exc:
        log.error(f"Failed to save stats JSON: {exc}")


def main(argv: Optional[List[str]] = None):
    args = parse_args(argv)
    log.setLevel(logging.DEBUG if args.verbose else logging.INFO)
    RUNTIME.update(wiring=args.wiring, seed=args.seed, device=args.device, precision=args.precision, engine=args.engine, devices=None)
    # torchrun (one process per GPU): join the job; rank 0 writes the outputs
    owns_pg = False
    rank = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.devices is None:
        import torch
        import torch.distributed as tdist
        if not tdist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            local_rank = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local_rank)
            tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            owns_pg = True
        rank = tdist.get_rank()

    log.info(f"Starting Conclave simulation with {args.num_simulations} simulations.")
    log.info(f"Parameters: Max Rounds={args.max_rounds}, Supermajority={args.supermajority_threshold*100:.0f}%"
             f", Runoff Rounds={args.runoff_rounds}, Runoff Candidates={args.runoff_candidates}")
    log.info(f"Model Params: Initial Beta={args.initial_beta_weight}, Stickiness={args.stickiness_factor}, Bandwagon={args.bandwagon_strength}, "
             f"Regional Bonus={args.regional_bonus}, Papabile Factor={args.papabile_weight_factor}")
    if args.beta_increment_amount > 0:
        log.info(f"Dynamic Beta Enabled: Increment={args.beta_increment_amount} per {args.beta_increment_interval_rounds} rounds.")

    elector_data_full = ElectorDataIngester(args.elector_data_file).load_and_prepare_data()
    if elector_data_full.empty:
        log.error("Failed to load or prepare elector data. Exiting.")
        return
    model_params = model_params_from_args(args)

    start = time.time()
    results: Any = []
    aggregate_stats = None
    try:
        tapes = load_tapes(args.tape, args.num_simulations, args.max_rounds, len(elector_data_full)) if args.tape else None
        batch = run_simulations(
            elector_data_full, model_params, args.num_simulations, max_rounds=args.max_rounds,
            supermajority_threshold=args.supermajority_threshold, runoff_threshold_rounds=args.runoff_rounds,
            runoff_candidate_count=args.runoff_candidates, enable_candidate_fatigue=args.enable_candidate_fatigue,
            fatigue_threshold_rounds=args.fatigue_threshold_rounds,
            fatigue_vote_share_threshold=args.fatigue_vote_share_threshold,
            fatigue_top_n_immune=args.fatigue_top_n_immune, enable_stop_candidate=args.enable_stop_candidate,
            uniform_tapes=tapes, devices=args.devices, workers=args.workers)
        log.info(f"Engine: {batch.engine} ({batch.precision}) on CUDA device(s) {list(batch.devices)}"
                 + (" [AUTO fell through from its first choice]" if batch.fell_through else ""))
        ids = list(elector_data_full.index)
        # the results table column-wise and the statistics straight from the device histograms: the same values as
        # results_from_batch + aggregate_results (tests/test_gpu_parity.py::test_sweep_stats_and_legacy_shim), no per-trial Python
        results = results_frame(batch, ids, args.supermajority_threshold)
        aggregate_stats = stats_from_histograms(batch.winner_hist[0], batch.rounds_hist[0], batch.totals[0], ids, args.num_simulations,
                                                winners_in_order=batch.winner)
    except Exception as exc:   # the reference logs worker errors and carries on with what it has (:350-351)
        log.error(f"Error in simulation worker: {exc}", exc_info=True)
    log.info(f"Simulation time: {time.time() - start:.2f} seconds")

    if aggregate_stats is None:
        aggregate_stats = aggregate_results([], args.num_simulations)
    if rank != 0:                       # torchrun: the rows and the reduced histograms are on rank 0, which writes the outputs
        if owns_pg:
            import torch.distributed as tdist
            tdist.destroy_process_group()
        return aggregate_stats
    log.info(f"Calculated aggregate stats: {aggregate_stats}")
    write_outputs(results, aggregate_stats, args.output_results, args.output_stats)
    if owns_pg:
        import torch.distributed as tdist
        tdist.destroy_process_group()

    sr = aggregate_stats["success_rate"]
    avg = aggregate_stats["average_rounds_successful"]
    log.info("\n--- Simulation Finished ---")
    log.info(f"Success Rate: {sr*100:.2f}%")
    log.info(f"Average Rounds (Successful): {avg if avg is not None else 'N/A'}")
    log.info(f"Most Frequent Winner: Elector {aggregate_stats['most_frequent_winner']} "
             f"({aggregate_stats['most_frequent_winner_count']} times, {aggregate_stats['most_frequent_winner_percentage']*100:.2f}%)")
    return aggregate_stats


# ---------------------------------------------------------------------------------------------
# Aggregation straight from the device histograms: one stats dict per parameter
# set of a sweep, without touching the per-trial rows.
# ---------------------------------------------------------------------------------------------
def stats_from_histograms(winner_hist: np.ndarray, rounds_hist: np.ndarray, totals: np.ndarray,
                          candidate_ids: List[Any], num_simulations: Optional[int] = None,
                          winners_in_order: Optional[np.ndarray] = None) -> Dict[str, Any]:
    """Same keys / values as aggregate_results (src/simulate.py:356-387) from
    winner_hist[E+1] (slot E = no winner), rounds_hist[R+1] (winning trials) and totals = (sum of rounds, trials).
    The reference's `winner_counts` is a Counter filled in completion order and `most_frequent_winner` is the FIRST-inserted
    among equal counts (:367-370).  `winners_in_order` (the per-trial winner indices in sim-id order, -1 = none) reproduces
    exactly that for a pool that completes in order — byte-identical JSON (tests/golden/main/*.json); without it keys and ties
    go by roster order (a sweep has no per-trial rows on the host)."""
    winner_hist = np.asarray(winner_hist, dtype=np.int64)
    rounds_hist = np.asarray(rounds_hist, dtype=np.int64)
    E = len(winner_hist) - 1
    n = int(totals[1]) if num_simulations is None else int(num_simulations)
    successful = int(winner_hist[:E].sum())
    rvals = np.nonzero(rounds_hist)[0]
    if successful:
        cum = np.cumsum(rounds_hist)
        lo = int(np.searchsorted(cum, (successful - 1) // 2 + 1))            # lower middle (1-based rank)
        hi = int(np.searchsorted(cum, successful // 2 + 1))                  # upper middle
        avg = float((np.arange(len(rounds_hist)) * rounds_hist).sum() / successful)
        median = (lo + hi) / 2.0
        rmin, rmax = int(rvals[0]), int(rvals[-1])
    else:
        avg = median = rmin = rmax = None
    order = np.nonzero(winner_hist[:E] > 0)[0]
    if winners_in_order is not None and len(order):
        w = np.asarray(winners_in_order)
        first = pd.unique(w[w >= 0])                                   # hash-based, keeps the order of first appearance
        if len(first) == len(order):
            order = first
    counts = {str(candidate_ids[i]): int(winner_hist[i]) for i in order}
    if counts:
        best = int(order[int(np.argmax(winner_hist[order]))])          # the first of the equal maxima in that order
        mf, mfc = str(candidate_ids[best]), int(winner_hist[best])
    else:
        mf, mfc = None, 0
    return {
        "total_simulations": n,
        "successful_simulations": successful,
        "success_rate": successful / n if n > 0 else 0,
        "average_rounds_successful": avg,
        "min_rounds_successful": rmin,
        "max_rounds_successful": rmax,
        "median_rounds_successful": median,
        "winner_counts": counts,
        "most_frequent_winner": mf,
        "most_frequent_winner_count": mfc,
        "most_frequent_winner_percentage": mfc / n if n > 0 else 0,
        "average_rounds_all": float(totals[0]) / float(totals[1]) if totals[1] > 0 else 0,
    }


def run_sweep(elector_data_full: pd.DataFrame, param_sets: List[Dict[str, Any]], num_simulations: int, *,
              seed: Optional[int] = None, wiring: Optional[str] = None, device: Optional[int] = None,
              precision: Optional[str] = None, engine: Optional[str] = None, devices=None) -> List[Dict[str, Any]]:
    """Parameter sweep (BASELINE config 4): every entry of `param_sets` is a dict with the keys of
    make_params (a 'model_params' dict plus max_rounds / supermajority_threshold / runoff_* ...); all sets run in
    ONE launch with per-CTA parameters (per device when `devices` names several: every device runs its share of the trials
    of EVERY set) and each gets its stats dict from the device histograms.  Every set of one sweep must share max_rounds
    (csim_run's batch limit); sweep max_rounds with one call per value."""
    wiring = wiring or RUNTIME["wiring"]
    devs = resolve_devices(RUNTIME["devices"] if devices is None else devices, device)
    eng = get_group(devs) if len(devs) > 1 else get_engine(devs[0])
    x, g, pap = roster_arrays(elector_data_full)
    eng.upload_roster(x, g, pap)
    sets = []
    for ps in param_sets:
        kw = dict(ps)
        mp = kw.pop("model_params", {})
        sets.append(make_params(mp, **kw))
    res = eng.run(sets, num_simulations, seed=_seed() if seed is None else seed, wiring=_WIRING[wiring],
                  engine=_ENGINE[engine or RUNTIME["engine"]], precision=_PRECISION[precision or RUNTIME["precision"]],
                  want_reason=False)
    ids = list(elector_data_full.index)
    return [stats_from_histograms(res.winner_hist[i], res.rounds_hist[i], res.totals[i], ids, num_simulations)
            for i in range(len(sets))]


# ---------------------------------------------------------------------------------------------
# Legacy entry point: `src/__main__.py:18,118` and the stale tests/test_simulate.py
# import a `run_monte_carlo_simulation` that no longer exists upstream; this shim provides it.
# ---------------------------------------------------------------------------------------------
REQUIRED_MAJORITY_FRACTION = DEFAULT_SUPERMAJORITY_THRESHOLD
MAX_ROUNDS = DEFAULT_MAX_ROUNDS


def run_monte_carlo_simulation(num_simulations: int, elector_data: pd.DataFrame, model_parameters: Dict[str, Any],
                               verbose: bool = False, max_rounds: int = MAX_ROUNDS,
                               supermajority_threshold: float = REQUIRED_MAJORITY_FRACTION):
    """(results_df[simulation_id, winner_id, rounds_taken, status], aggregate_stats) — the old schema
    (tests/test_simulate.py:60-68).  `model_parameters` may use the old key 'beta_weight'."""
    if elector_data is None or elector_data.empty:
        raise ValueError("Elector data cannot be empty")
    if "elector_id" not in elector_data.columns and elector_data.index.name != "elector_id":
        raise ValueError("Elector data must contain an 'elector_id' column")
    df = elector_data
    if "elector_id" in df.columns:
        df = df.copy()
        df["elector_id"] = df["elector_id"].astype(str)
        df = df.set_index("elector_id")
    if "region" not in df.columns:
        df = df.assign(region=[f"_r{i}" for i in range(len(df))])      # no regional information: nobody shares a region
    mp = dict(model_parameters or {})
    if "beta_weight" in mp:
        mp["initial_beta_weight"] = mp.pop("beta_weight")
    res = run_simulations(df, mp, num_simulations, max_rounds=max_rounds, supermajority_threshold=supermajority_threshold,
                          runoff_threshold_rounds=DEFAULT_RUNOFF_THRESHOLD_ROUNDS,
                          runoff_candidate_count=DEFAULT_RUNOFF_CANDIDATE_COUNT)
    ids = list(df.index)
    rows = [{"simulation_id": i, "winner_id": ids[w] if w >= 0 else None, "rounds_taken": int(r),
             "status": "Success" if w >= 0 else "Max rounds reached"} for i, (w, r) in enumerate(zip(res.winner, res.rounds))]
    results_df = pd.DataFrame(rows, columns=["simulation_id", "winner_id", "rounds_taken", "status"])
    ok = res.winner >= 0
    stats = {"num_simulations": num_simulations, "success_rate": float(ok.mean()) if num_simulations else 0.0,
             "avg_rounds_success": float(res.rounds[ok].mean()) if ok.any() else None,
             "winner_counts": {ids[i]: int(c) for i, c in enumerate(res.winner_hist[0, :-1]) if c > 0}}
    return results_df, stats


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench_configs.py — the five BASELINE.json configurations on ONE B200 (supplement to bench.py).

    python bench_configs.py [--quick] > profiles/rNN_configs.json

cfg-1  CLI defaults (thr 2/3, static beta), 100 trials, real-shape roster (E=134 synthetic stand-in)
cfg-2  README parameters x 1000 trials
cfg-3  README parameters x 10M trials (this GPU's 1/8 share: 1.25M)           -> see bench.py for the 8-GPU run
cfg-4  parameter sweep: 256 parameter sets x 100k trials (SURVEY 8d grid), one launch, per-CTA parameters
cfg-5  synthetic roster 1024 x 1024, max_rounds 50, 1M trials (default here: 100k; --full for 1M)
Each line: engine, wall time through the host API (results copied back), conclaves/s, elector-rounds/s.
"""
import argparse
import itertools
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from bench import synthetic_roster, CFG2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    from conclave_sim_b200 import Engine, make_params, FP32, FP64, WIRING_REF_CLI, WIRING_MODEL, ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX, ENGINE_FULL
    eng = Engine(0)
    out = []

    def run(name, E, sets, n, wiring, engine, precision, reps=2):
        x, g, p = synthetic_roster(E)
        eng.upload_roster(x, g, p)
        best = None
        r = None
        eng.run(sets, min(n, 1000), seed=1, wiring=wiring, engine=engine, precision=precision, want_reason=False)   # untimed: tables, first launch
        for _ in range(max(reps, 2)):
            r = None                              # release the previous result first: its page-locked arrays go back to the pool
            t0 = time.perf_counter()
            r = eng.run(sets, n, seed=20250507, wiring=wiring, engine=engine, precision=precision, want_reason=False)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        trials = n * len(sets)
        rec = dict(config=name, E=E, param_sets=len(sets), trials=trials,
                   wiring="ref_cli" if wiring == WIRING_REF_CLI else "model",
                   engine={ENGINE_DENSE: "dense", ENGINE_TABLE: "table", ENGINE_PREFIX: "prefix", ENGINE_FULL: "full"}[engine], precision="fp32" if precision == FP32 else "fp64",
                   wall_s=best, kernel_ms=r.kernel_ms, conclaves_per_s=trials / best, elector_rounds_per_s=r.elector_rounds / best,
                   success_rate=float((r.winner >= 0).mean()), mean_rounds=float(r.rounds.mean()))
        out.append(rec)
        print(json.dumps(rec), flush=True)

    cfg1 = make_params(dict(initial_beta_weight=1.0, beta_increment_amount=0.05, beta_increment_interval_rounds=10,
                            stickiness_factor=0.5, regional_affinity_bonus=0.1, papabile_weight_factor=1.5),
                       max_rounds=20, supermajority_threshold=2 / 3)
    cfg2 = make_params(CFG2, max_rounds=20, supermajority_threshold=0.56)
    for eng_id in (ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX):
        run("cfg1 defaults x100", 134, [cfg1], 100, WIRING_REF_CLI, eng_id, FP32)
        run("cfg2 README x1000", 134, [cfg2], 1000, WIRING_REF_CLI, eng_id, FP32)
        run("cfg3 README x1.25M (1/8 of 10M)", 134, [cfg2], 125_000 if a.quick else 1_250_000, WIRING_REF_CLI, eng_id, FP32)
    run("cfg2 README x1000 (fp64 exact)", 134, [cfg2], 1000, WIRING_REF_CLI, ENGINE_TABLE, FP64)
    for eng_id in (ENGINE_DENSE, ENGINE_PREFIX):
        run("cfg3 model wiring x1.25M", 134, [cfg2], 125_000 if a.quick else 1_250_000, WIRING_MODEL, eng_id, FP32)
    cfg2x = make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), max_rounds=20, supermajority_threshold=0.56,
                        enable_candidate_fatigue=True, enable_stop_candidate=True)
    run("cfg3 model wiring + fatigue + stop-candidate x250k", 134, [cfg2x], 25_000 if a.quick else 250_000, WIRING_MODEL, ENGINE_FULL, FP32)
    run("cfg3 model wiring + fatigue + stop-candidate x25k (fp64 exact)", 134, [cfg2x], 25_000, WIRING_MODEL, ENGINE_DENSE, FP64, reps=1)
    # cfg-4: SURVEY 8d grid
    grid = list(itertools.product([0.5, 1.0, 1.5, 2.0], [0.0, 0.1, 0.3, 0.5], [1.0, 1.5, 2.0, 3.0], [0.5, 0.56, 0.6, 2 / 3]))
    sets = [make_params(dict(CFG2, initial_beta_weight=b0, beta_increment_amount=db, papabile_weight_factor=pw),
                        max_rounds=20, supermajority_threshold=thr) for b0, db, pw, thr in grid]
    for eng_id in (ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX):
        run("cfg4 sweep 256 sets x 100k", 134, sets, 10_000 if a.quick else 100_000, WIRING_REF_CLI, eng_id, FP32, reps=1)
    # cfg-5
    cfg5 = make_params(CFG2, max_rounds=50, supermajority_threshold=0.56)
    n5 = 10_000 if a.quick else (1_000_000 if a.full else 100_000)
    run("cfg4 sweep 256 sets x 100k, model wiring", 134, sets, 10_000 if a.quick else 100_000, WIRING_MODEL, ENGINE_PREFIX, FP32, reps=1)
    for eng_id in (ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX):
        run(f"cfg5 1024x1024 R=50 x{n5}", 1024, [cfg5], n5, WIRING_REF_CLI, eng_id, FP32, reps=1)
    run(f"cfg5 1024x1024 R=50 x{n5}, model wiring", 1024, [cfg5], n5, WIRING_MODEL, ENGINE_PREFIX, FP32, reps=1)


if __name__ == "__main__":
    main()

"""Independently seeded samples of the UNMODIFIED reference for the native-RNG statistical gate (north_star: chi-square / KS at
p > 0.01 "against the reference's own CPU run"), beyond cfg-2 (tests/golden/dist_cfg2.json, oracle/gen_golden.py --dist):

  cfg1    CLI defaults on the real roster, R = 20 (BASELINE configs[0]): nobody reaches 2/3, so the winner / rounds histograms are
          degenerate; the informative distributions are the plurality LEADER of the last round and its vote count
  cfg4a   one set of the cfg-4 sweep grid (beta0 0.5, +0.1 / round, papabile 3.0, threshold 0.5), real roster
  cfg5    the cfg-5 shape: synthetic 1024 x 1024 roster (SURVEY 8d recipe), R = 50, cfg-2 parameters

TEST INFRASTRUCTURE ONLY; runs in the build container.  Every trial: np.random.seed(seed0 + i), then the reference's own
_run_single_simulation_worker (forked workers would otherwise share one RNG state).

    python -m oracle.gen_golden_dist cfg1 4000
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

from oracle import ref_harness as H
from oracle import conclave_oracle as O
from oracle.gen_golden import CFG1, CFG2, GOLD

CASES = {
    "cfg1": dict(roster="real", mp=CFG1, R=20, thr=2 / 3, rr=5, k=2, seed0=7100003),
    "cfg4a": dict(roster="real", mp=dict(CFG2, initial_beta_weight=0.5, beta_increment_amount=0.1, papabile_weight_factor=3.0), R=20, thr=0.5,
                  rr=5, k=2, seed0=7200003),
    "cfg5": dict(roster=1024, mp=CFG2, R=50, thr=0.56, rr=5, k=2, seed0=7300003),
}
_STATE = {}


def _df(case):
    key = str(case["roster"])
    if key not in _STATE:
        if case["roster"] == "real":
            _STATE[key] = H.load_roster_df()
        else:
            ro = O.synthetic_roster(int(case["roster"]))
            _STATE[key] = H.make_df(ro.ideology, np.array([f"R{g}" for g in ro.region], dtype=object), ro.papabile)
    return _STATE[key]


def _worker(args):
    name, i = args
    case = CASES[name]
    H.quiet()
    _, ref_sim = H.load_reference()
    df = _df(case)
    np.random.seed(case["seed0"] + i)
    r = ref_sim._run_single_simulation_worker(i, df, case["mp"], case["R"], case["thr"], case["rr"], case["k"], False, False, 3, 0.05, 3, False)
    ids = list(df.index)
    fv = r["final_votes"]
    votes = np.array([fv.get(c, 0) for c in ids])
    leader = int(np.argmax(votes))                      # first maximum = idxmax
    return (-1 if r["winner"] is None else ids.index(r["winner"]), int(r["rounds"]), leader, int(votes[leader]))


def main():
    name, n = sys.argv[1], int(sys.argv[2])
    case = CASES[name]
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    t0 = time.time()
    with mp.Pool(cores) as pool:
        res = pool.map(_worker, [(name, i) for i in range(n)], chunksize=4)
    wall = time.time() - t0
    w = np.array([a[0] for a in res]); r = np.array([a[1] for a in res]); ld = np.array([a[2] for a in res]); lv = np.array([a[3] for a in res])
    E = len(_df(case))
    doc = dict(case=name, roster=str(case["roster"]), model_params=case["mp"], max_rounds=case["R"], threshold=case["thr"], runoff=[case["rr"], case["k"]],
               n=n, seed0=case["seed0"], seeding="np.random.seed(seed0+i) per trial; unmodified _run_single_simulation_worker",
               winner_hist=np.bincount(w[w >= 0], minlength=E).tolist(), no_winner=int(np.sum(w < 0)),
               rounds_hist_success=np.bincount(r[w >= 0], minlength=case["R"] + 1).tolist(),
               rounds_hist_all=np.bincount(r, minlength=case["R"] + 1).tolist(),
               final_leader_hist=np.bincount(ld, minlength=E).tolist(), final_leader_votes_hist=np.bincount(lv, minlength=E + 1).tolist(),
               wall_s=wall, cores=cores, conclaves_per_s=n / wall, elector_rounds_per_s=float(r.sum()) * E / wall)
    with open(os.path.join(GOLD, f"dist_{name}.json"), "w") as f:
        json.dump(doc, f)
    print(name, n, "trials", f"{wall:.1f}s", "success", 1 - doc["no_winner"] / n, "mean rounds", r.mean(), flush=True)


if __name__ == "__main__":
    main()

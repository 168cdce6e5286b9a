#!/usr/bin/env python
"""Golden probability matrices WITH candidate fatigue and stop-candidate on full-size rosters, from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (build container: /root/reference must be present).  Writes tests/golden/model_extras.npz:

  real/…    the real 134-elector roster (tests/golden/roster_real.npz), README parameters (cfg-2) + enable_candidate_fatigue +
            enable_stop_candidate (unacceptable distance 0.3, as bench.py times the FULL engine), rounds 1..4 of ONE model instance
            (the beta schedule advances as in a trial, src/model.py:562-580); the previous votes concentrate on three front-runners in
            rounds 2 and 4 (threats above the 15 % share: model.py:384-415) and are spread one per candidate in round 3 (no threat);
            a random fatigue set each round (model.py:372-379)
  syn200/…  a synthetic 200-elector roster (oracle.conclave_oracle.synthetic_roster), round 2 with threats and a fatigue set

The random model cases of model_cases.npz have at most 24 electors; these pin the kernels that fill whole Philox quads (E >= 128:
the FULL engine's fast path, dense32.cuh EXT) to the reference's own TransitionModel.calculate_transition_probabilities
(src/model.py:269-560) at north_star's 1e-5, and csim_prob_matrix (fp64) at 1e-13.

    python oracle/gen_golden_extras.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_harness as H          # noqa: E402
from oracle import conclave_oracle as O      # noqa: E402
from oracle.gen_golden import CFG2           # noqa: E402

KW = dict(CFG2, enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.3)


def votes(rng, E, threats):
    if threats:                                  # three front-runners take ~60 % of the votes
        front = rng.choice(E, 3, replace=False)
        pv = np.where(rng.random(E) < 0.6, front[rng.integers(0, 3, E)], rng.integers(0, E, E)).astype(np.int32)
    else:                                        # one vote each (up to the self-vote fix below): every share far below 15 %
        pv = rng.permutation(E).astype(np.int32)
    me = np.arange(E)
    pv[pv == me] = (pv[pv == me] + 1) % E
    return pv


def run_case(out, key, df, rounds_spec, rng):
    import pandas as pd
    ref_model, _ = H.load_reference()
    E = len(df)
    ids = list(df.index)
    model = ref_model.TransitionModel(df.copy(), **KW)
    for r, threats in rounds_spec:
        if r == 1:
            pv = np.full(E, -1, np.int32); ser = None
        else:
            pv = votes(rng, E, threats)
            ser = pd.Series({ids[e]: ids[pv[e]] for e in range(E)}, dtype=object)
        fat = rng.random(E) < 0.55
        cnt = np.bincount(pv[pv >= 0], minlength=E)
        shares = cnt / float(E)                  # what a trial hands the model (src/simulate.py:73)
        P, _ = model.calculate_transition_probabilities(ser, r, set(np.where(fat)[0].tolist()), shares)
        out[f"{key}/prev_vote{r}"] = pv
        out[f"{key}/fatigued{r}"] = fat
        out[f"{key}/shares{r}"] = shares
        out[f"{key}/P{r}"] = np.asarray(P, np.float64)
        out[f"{key}/beta{r}"] = np.array(model.effective_beta_weight)
        print(key, "round", r, "threats", bool((shares > KW["stop_candidate_threat_min_vote_share"]).any()), "fatigued", int(fat.sum()),
              "beta", model.effective_beta_weight)


def main():
    H.quiet()
    H.load_reference()                           # (sets the pandas option the reference needs BEFORE any frame is built)
    out = {}
    z = np.load(os.path.join(GOLD, "roster_real.npz"))
    df = H.make_df(z["ideology"], z["region_names"][z["region"]], z["papabile"], ids=[str(s) for s in z["ids"]])
    run_case(out, "real", df, [(1, False), (2, True), (3, False), (4, True)], np.random.default_rng(20251020))
    syn = O.synthetic_roster(200)
    df2 = H.make_df(syn.ideology, np.array([f"R{g}" for g in syn.region]), syn.papabile)
    out["syn200/ideology"] = syn.ideology; out["syn200/region"] = syn.region.astype(np.int32); out["syn200/papabile"] = syn.papabile
    run_case(out, "syn200", df2, [(1, False), (2, True)], np.random.default_rng(20251021))
    out["meta"] = np.array(json.dumps(dict(kwargs=KW)))
    np.savez_compressed(os.path.join(GOLD, "model_extras.npz"), **out)
    print("wrote", os.path.join(GOLD, "model_extras.npz"), os.path.getsize(os.path.join(GOLD, "model_extras.npz")), "bytes")


if __name__ == "__main__":
    main()

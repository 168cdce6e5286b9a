"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).

    python -m oracle.gen_golden            # model + engine vectors (~1-2 min)
    python -m oracle.gen_golden --dist 20000   # + distributional sample (~8 min on 8 cores)

Outputs (committed):
  tests/golden/roster_real.npz      the three columns of merged_electors.csv the engine reads
                                    (model.py:199-213), region factorised
  tests/golden/model_cases.npz      TransitionModel.calculate_transition_probabilities outputs
                                    on random small rosters, every term exercised (`model` wiring)
  tests/golden/model_subset.npz     same with effective_candidate_ids column subsets (model.py:280-296)
  tests/golden/model_real.npz       same on the real roster (cfg-1 / cfg-2, selected rounds)
  tests/golden/engine_cases.json    engine-level case manifest (params, seeds)
  tests/golden/engine_cases.npz     winners / rounds / per-round tallies of
                                    _run_single_simulation_worker under MT19937 tapes
                                    (tape of trial i = RandomState(seed_i).random_sample(R*E))
  tests/golden/dist_cfg2.json       winner / rounds histograms of an independently seeded
                                    reference sample (README params), for the chi2 / KS gate
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402

CFG1 = dict(  # CLI defaults, simulate.py:235-242,249,255-257
    initial_beta_weight=1.0, enable_dynamic_beta=False, beta_increment_amount=0.05,
    beta_increment_interval_rounds=10, stickiness_factor=0.5, bandwagon_strength=0.0,
    regional_affinity_bonus=0.1, papabile_weight_factor=1.5, fatigue_penalty_factor=0.5,
    stop_candidate_threshold_unacceptable_distance=0.5, stop_candidate_boost_factor=1.5,
    stop_candidate_threat_min_vote_share=0.15)
CFG2 = dict(  # README.md:37-47
    initial_beta_weight=1.5, enable_dynamic_beta=True, beta_increment_amount=0.3,
    beta_increment_interval_rounds=1, stickiness_factor=0.8, bandwagon_strength=0.7,
    regional_affinity_bonus=0.1, papabile_weight_factor=2.0, fatigue_penalty_factor=0.5,
    stop_candidate_threshold_unacceptable_distance=0.5, stop_candidate_boost_factor=1.5,
    stop_candidate_threat_min_vote_share=0.15)


def factorise(region):
    names, codes = np.unique(np.asarray(region).astype(str), return_inverse=True)
    return names, codes.astype(np.int32)


def gen_roster():
    df = H.load_roster_df()
    names, codes = factorise(df["region"].values)
    np.savez_compressed(
        os.path.join(GOLD, "roster_real.npz"),
        ideology=df["ideology_score"].values.astype(np.float64),
        region=codes, region_names=names.astype("U32"),
        papabile=df["is_papabile"].astype(bool).values,
        ids=np.array(list(df.index)).astype("U8"))
    return df


def gen_model_cases():
    import pandas as pd
    ref_model, _ = H.load_reference()
    rng = np.random.default_rng(20250507)
    regs = np.array(["R0", "R1", "R2", "R3"])
    out = {}
    meta = []
    for t in range(28):
        E = int(rng.integers(2, 25))
        ideo = np.round(rng.random(E), 2) if t % 2 else rng.random(E)
        if t == 5:
            ideo = np.zeros(E); ideo[0] = 800.0     # exp underflow -> zero row (model.py:447-449)
        reg = rng.integers(0, 4, E)
        pap = rng.random(E) < 0.3
        kw = dict(
            initial_beta_weight=float(rng.uniform(0.2, 4)), enable_dynamic_beta=bool(rng.integers(2)),
            beta_growth_rate=float(rng.choice([1.0, 1.05, 1.2])),
            beta_increment_amount=(None if rng.integers(2) else float(rng.uniform(0, 0.5))),
            beta_increment_interval_rounds=int(rng.integers(1, 4)),
            stickiness_factor=float(rng.choice([0, 0.5, 0.8])), bandwagon_strength=float(rng.choice([0, 0.3, 0.7])),
            regional_affinity_bonus=(0.0 if t == 5 else float(rng.choice([0, 0.1, 0.3]))),
            papabile_weight_factor=float(rng.choice([1, 1.5, 2])),
            enable_candidate_fatigue=bool(rng.integers(2)), fatigue_penalty_factor=float(rng.choice([0.5, 0.25])),
            enable_stop_candidate=bool(rng.integers(2)),
            stop_candidate_threshold_unacceptable_distance=float(rng.uniform(0.2, 0.7)),
            stop_candidate_threat_min_vote_share=0.15, stop_candidate_boost_factor=float(rng.choice([1.5, 2.0])))
        if t == 5:
            kw["bandwagon_strength"] = 0.0
        df = H.make_df(ideo, regs[reg], pap)
        model = ref_model.TransitionModel(df.copy(), **kw)
        rounds = 4
        pvs, fats, shs, Ps, betas = [], [], [], [], []
        for r in range(1, rounds + 1):
            pv = rng.integers(0, E, E).astype(np.int32)
            if r == 1:
                pv[:] = -1
            elif t % 3 == 0:
                pv[rng.random(E) < 0.2] = -1        # electors without a recorded previous vote
            ser = None
            if r > 1:
                ser = pd.Series({str(e): str(pv[e]) for e in range(E) if pv[e] >= 0}, dtype=object)
            fat = rng.random(E) < 0.2
            cnt = np.bincount(pv[pv >= 0], minlength=E)
            shares = cnt / max(1, cnt.sum())
            P, _ = model.calculate_transition_probabilities(ser, r, set(np.where(fat)[0].tolist()), shares)
            pvs.append(pv); fats.append(fat); shs.append(shares); Ps.append(P); betas.append(model.effective_beta_weight)
        k = f"c{t:02d}"
        out[k + "/ideology"] = ideo; out[k + "/region"] = reg.astype(np.int32); out[k + "/papabile"] = pap
        out[k + "/prev_vote"] = np.stack(pvs); out[k + "/fatigued"] = np.stack(fats)
        out[k + "/shares"] = np.stack(shs); out[k + "/P"] = np.stack(Ps); out[k + "/beta"] = np.array(betas)
        meta.append(dict(name=k, kwargs=kw))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "model_cases.npz"), **out)


def gen_model_subset():
    """effective_candidate_ids (model.py:280-296): column subsets in caller order, every term exercised."""
    import pandas as pd
    ref_model, _ = H.load_reference()
    rng = np.random.default_rng(424242)
    regs = np.array(["R0", "R1", "R2"])
    out = {}
    meta = []
    for t in range(10):
        E = int(rng.integers(4, 20))
        ideo = rng.random(E)
        reg = rng.integers(0, 3, E)
        pap = rng.random(E) < 0.4
        kw = dict(initial_beta_weight=float(rng.uniform(0.5, 3)), stickiness_factor=float(rng.choice([0, 0.6])),
                  bandwagon_strength=float(rng.choice([0, 0.5])), regional_affinity_bonus=float(rng.choice([0, 0.2])),
                  papabile_weight_factor=float(rng.choice([1, 2])), enable_candidate_fatigue=bool(rng.integers(2)),
                  fatigue_penalty_factor=0.5, enable_stop_candidate=bool(rng.integers(2)),
                  stop_candidate_threshold_unacceptable_distance=float(rng.uniform(0.2, 0.6)),
                  stop_candidate_threat_min_vote_share=0.1, stop_candidate_boost_factor=1.7)
        df = H.make_df(ideo, regs[reg], pap)
        model = ref_model.TransitionModel(df.copy(), **kw)
        n_c = int(rng.integers(1, E + 1))
        cand = rng.permutation(E)[:n_c].astype(np.int32)
        pv = rng.integers(0, E, E).astype(np.int32)
        ser = pd.Series({str(e): str(pv[e]) for e in range(E)}, dtype=object)
        fat = rng.random(E) < 0.25
        cnt = np.bincount(pv, minlength=E)
        shares = cnt / cnt.sum()
        P, _ = model.calculate_transition_probabilities(ser, 2, set(np.where(fat)[0].tolist()), shares,
                                                        effective_candidate_ids=[str(c) for c in cand] + ["not-a-candidate"])
        k = f"s{t:02d}"
        out[k + "/ideology"] = ideo; out[k + "/region"] = reg.astype(np.int32); out[k + "/papabile"] = pap
        out[k + "/prev_vote"] = pv; out[k + "/fatigued"] = fat; out[k + "/shares"] = shares; out[k + "/cand"] = cand
        out[k + "/P"] = P; out[k + "/beta"] = np.array(model.effective_beta_weight)
        meta.append(dict(name=k, kwargs=kw))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "model_subset.npz"), **out)


def gen_model_real(df):
    import pandas as pd
    ref_model, _ = H.load_reference()
    out = {}
    for name, kw, rounds in (("cfg1", CFG1, (1,)), ("cfg2", CFG2, (1, 6, 20))):
        model = ref_model.TransitionModel(df.copy(), **kw)
        for r in range(1, max(rounds) + 1):
            P, _ = model.calculate_transition_probabilities(None, r)
            if r in rounds:
                out[f"{name}/P{r}"] = P
                out[f"{name}/beta{r}"] = np.array(model.effective_beta_weight)
    # model wiring on the real roster: live bandwagon + stickiness at round 2
    rng = np.random.default_rng(7)
    E = len(df)
    pv = rng.integers(0, E, E).astype(np.int32)
    pv[pv == np.arange(E)] = 0
    ids = list(df.index)
    ser = pd.Series({ids[e]: ids[pv[e]] for e in range(E)}, dtype=object)
    model = ref_model.TransitionModel(df.copy(), **CFG2)
    model.calculate_transition_probabilities(None, 1)
    P, _ = model.calculate_transition_probabilities(ser, 2)
    out["cfg2_model/prev_vote"] = pv
    out["cfg2_model/P2"] = P
    out["cfg2_model/beta2"] = np.array(model.effective_beta_weight)
    np.savez_compressed(os.path.join(GOLD, "model_real.npz"), **out)


def engine_case_list():
    rng = np.random.default_rng(99)
    small = dict(ideology=[0.1, 0.1, 0.5, 0.5, 0.9], region=[0, 0, 1, 1, 0], papabile=[1, 0, 1, 0, 0])
    cases = [
        dict(name="cfg1", roster="real", mp=CFG1, R=20, thr=2 / 3, rr=5, k=2, n=10, seed0=4242),
        dict(name="cfg2", roster="real", mp=CFG2, R=20, thr=0.56, rr=5, k=2, n=64, seed0=1000003),
        dict(name="cfg2_k3_rr3", roster="real", mp=CFG2, R=15, thr=0.56, rr=3, k=3, n=12, seed0=31),
        dict(name="cfg2_k0", roster="real", mp=CFG2, R=12, thr=0.56, rr=5, k=0, n=6, seed0=51),
        dict(name="cfg2_rr1", roster="real", mp=CFG2, R=10, thr=0.56, rr=1, k=2, n=12, seed0=71),
        dict(name="cfg2_lowthr", roster="real", mp=CFG2, R=10, thr=0.05, rr=5, k=2, n=12, seed0=91),
        dict(name="cfg1_iv3", roster="real", mp=dict(CFG1, enable_dynamic_beta=True, beta_increment_amount=0.5,
                                                    beta_increment_interval_rounds=3), R=12, thr=0.5, rr=4, k=2, n=10, seed0=111),
        dict(name="growth", roster="real", mp=dict(initial_beta_weight=2.0, enable_dynamic_beta=True, beta_growth_rate=1.2,
                                                   regional_affinity_bonus=0.05, papabile_weight_factor=3.0),
             R=12, thr=0.5, rr=5, k=2, n=10, seed0=131),
        dict(name="small5_ties", roster=small, mp=dict(initial_beta_weight=0.0, regional_affinity_bonus=0.0,
                                                       papabile_weight_factor=1.0), R=8, thr=0.8, rr=2, k=2, n=40, seed0=151),
        dict(name="small5_k4", roster=small, mp=CFG2, R=8, thr=0.8, rr=1, k=4, n=30, seed0=171),
        dict(name="two", roster=dict(ideology=[0.2, 0.7], region=[0, 1], papabile=[0, 1]), mp=CFG1, R=4, thr=2 / 3, rr=2, k=2, n=4, seed0=191),
        dict(name="one", roster=dict(ideology=[0.2], region=[0], papabile=[0]), mp=CFG1, R=3, thr=2 / 3, rr=2, k=2, n=2, seed0=201),
        dict(name="syn33", roster=dict(ideology=rng.beta(2, 4, 33).tolist(), region=rng.integers(0, 8, 33).tolist(),
                                       papabile=(rng.random(33) < 0.15).astype(int).tolist()),
             mp=CFG2, R=20, thr=0.56, rr=5, k=2, n=24, seed0=211),
        # `model` wiring (loop restated in ref_harness.run_model_wiring_trial; model = reference's)
        dict(name="m_cfg2", roster="real", wiring="model", mp=CFG2, R=20, thr=0.56, rr=5, k=2, n=24, seed0=777),
        dict(name="m_cfg2_stop", roster="real", wiring="model", mp=dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3),
             stop=True, R=12, thr=0.56, rr=5, k=2, n=10, seed0=801),
        dict(name="m_cfg2_fatigue", roster="real", wiring="model", mp=CFG2, fatigue=True, fat_n=3, fat_rounds=3, fat_thr=0.05,
             stop=True, R=12, thr=0.56, rr=5, k=2, n=16, seed0=821),
        dict(name="m_cfg2_fatigue_k0", roster="real", wiring="model", mp=dict(CFG2, bandwagon_strength=0.2), fatigue=True, fat_n=0,
             fat_rounds=2, fat_thr=0.02, R=10, thr=0.56, rr=5, k=0, n=8, seed0=841),
        dict(name="m_small5", roster=small, wiring="model", mp=CFG2, R=8, thr=0.8, rr=3, k=2, n=30, seed0=861),
        dict(name="m_syn33", roster=dict(ideology=rng.beta(2, 4, 33).tolist(), region=rng.integers(0, 8, 33).tolist(),
                                         papabile=(rng.random(33) < 0.15).astype(int).tolist()),
             wiring="model", mp=dict(CFG2, stickiness_factor=2.0, bandwagon_strength=1.5), R=20, thr=0.6, rr=5, k=2, n=24, seed0=881),
    ]
    return cases


def gen_engine_cases(df_real):
    cases = engine_case_list()
    out = {}
    for c in cases:
        if c["roster"] == "real":
            df = df_real
        else:
            ro = c["roster"]
            df = H.make_df(ro["ideology"], np.array([f"R{g}" for g in ro["region"]], dtype=object), np.array(ro["papabile"], bool))
        E = len(df)
        ids = list(df.index)
        n, R = c["n"], c["R"]
        winners = np.full(n, -1, np.int32); rounds = np.zeros(n, np.int32); reason = np.zeros(n, np.int8)
        tallies = np.full((n, R, E), -1, np.int16); used = np.zeros(n, np.int32); ties = np.zeros(n, bool)
        tape_head = np.zeros((n, 2))
        for i in range(n):
            tape = H.mt_tape(c["seed0"] + i, R * E)
            tape_head[i] = tape[:2]
            if c.get("wiring", "ref_cli") == "ref_cli":
                res, tl, _, u = H.run_worker_with_tape(df, c["mp"], tape, R, c["thr"], c["rr"], c["k"])
                winners[i] = -1 if res["winner"] is None else ids.index(res["winner"])
                rounds[i] = res["rounds"]; used[i] = u
                reason[i] = 1 if res["reason"].startswith("Winner") else (0 if res["reason"].startswith("Max") else 2)
                tallies[i, : tl.shape[0]] = tl
            else:
                w, rd, tl, _, _, tie = H.run_model_wiring_trial(
                    df, c["mp"], tape, R, c["thr"], c["rr"], c["k"],
                    enable_candidate_fatigue=c.get("fatigue", False), fatigue_threshold_rounds=c.get("fat_rounds", 3),
                    fatigue_vote_share_threshold=c.get("fat_thr", 0.05), fatigue_top_n_immune=c.get("fat_n", 3),
                    enable_stop_candidate=c.get("stop", False))
                winners[i] = w; rounds[i] = rd; used[i] = rd * E; reason[i] = 1 if w >= 0 else 0
                tallies[i, : tl.shape[0]] = tl; ties[i] = tie
        k = c["name"]
        out[k + "/winners"] = winners; out[k + "/rounds"] = rounds; out[k + "/reason"] = reason
        out[k + "/tallies"] = tallies; out[k + "/draws_used"] = used; out[k + "/fatigue_tie"] = ties
        out[k + "/tape_head"] = tape_head
        print(f"  {k}: E={E} n={n} winners={np.sum(winners >= 0)} mean_rounds={rounds.mean():.2f} ties={ties.sum()}", flush=True)
    np.savez_compressed(os.path.join(GOLD, "engine_cases.npz"), **out)
    with open(os.path.join(GOLD, "engine_cases.json"), "w") as f:
        json.dump(cases, f, indent=1)


def _dist_worker(args):
    i, seed0 = args
    H.quiet()
    _, ref_sim = H.load_reference()
    global _DF
    try:
        df = _DF
    except NameError:
        df = _DF = H.load_roster_df()
    np.random.seed(seed0 + i)  # forked pool workers otherwise share one RNG state
    r = ref_sim._run_single_simulation_worker(i, df, CFG2, 20, 0.56, 5, 2, False, False, 3, 0.05, 3, False)
    return (-1 if r["winner"] is None else int(r["winner"]), int(r["rounds"]))


def gen_dist(n, seed0=5000011):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    t0 = time.time()
    with mp.Pool(cores) as pool:
        res = pool.map(_dist_worker, [(i, seed0) for i in range(n)], chunksize=16)
    wall = time.time() - t0
    w = np.array([a for a, _ in res]); r = np.array([b for _, b in res])
    E = 134
    doc = dict(
        config="README params (cfg-2): R=20 thr=0.56 runoff 5/2, merged_electors.csv",
        n=n, seed0=seed0, seeding="np.random.seed(seed0+i) per trial; unmodified _run_single_simulation_worker",
        winner_hist=np.bincount(w[w >= 0], minlength=E).tolist(), no_winner=int(np.sum(w < 0)),
        rounds_hist_success=np.bincount(r[w >= 0], minlength=21).tolist(),
        rounds_hist_all=np.bincount(r, minlength=21).tolist(),
        wall_s=wall, cores=cores, conclaves_per_s=n / wall, elector_rounds_per_s=float(r.sum()) * E / wall)
    with open(os.path.join(GOLD, "dist_cfg2.json"), "w") as f:
        json.dump(doc, f)
    print("dist:", n, "trials", f"{wall:.1f}s", "success", 1 - doc["no_winner"] / n, "mean rounds(success)", r[w >= 0].mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dist", type=int, default=0)
    ap.add_argument("--only-dist", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    H.quiet()
    if not a.only_dist:
        df = gen_roster()
        gen_model_cases()
        gen_model_subset()
        gen_model_real(df)
        gen_engine_cases(df)
    if a.dist:
        gen_dist(a.dist)


if __name__ == "__main__":
    main()

"""Diagnostic (not a pytest file): the FULL engine's fast path (the EXT instantiation of the dense kernel, dense32.cuh) against the
fp64 C oracle and against the legacy k_trials_full32 (CSIM_FULL_LEGACY=1), case by case, one JSON line each — which term, which
round, which roster size.  Run on the GPU box:  python tests/ext_check.py [--time]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import conclave_oracle as O      # noqa: E402
from oracle import c_oracle as CO            # noqa: E402
import golden_util as G                      # noqa: E402
from test_gpu_parity import CFG2, to_params, upload   # noqa: E402

FULL, PREFIX = 4, 3


def emit(**kw):
    print(json.dumps(kw, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)), flush=True)


def probe_cases(eng):
    from conclave_sim_b200 import FP64
    for case in G.model_cases():
        mp, roster = case["mp"], case["roster"]
        if case["name"] == "c05" or not (mp.enable_candidate_fatigue or mp.enable_stop_candidate):
            continue
        upload(eng, roster)
        E = roster.E
        p = to_params(mp)
        for r in range(1, case["P"].shape[0] + 1):
            beta = case["beta"][r - 1]
            pv = None if r == 1 else case["prev_vote"][r - 1].astype(np.int32)
            pc = None if pv is None else np.bincount(pv[pv >= 0], minlength=E).astype(np.int32)
            fat = case["fatigued"][r - 1]
            shares = np.zeros(E) if pc is None else pc / float(E)
            try:
                P_ref = eng.prob_matrix(p, beta, pv, pc, fat, shares, precision=FP64)
                scores, prefix, chunk = eng.engine_probs(p, wiring=O.WIRING_MODEL, engine=FULL, round=r, prev_vote=pv, prev_count=pc, fatigued=fat)
                S = prefix[:, -1].astype(np.float64)
                P = scores.astype(np.float64) / S[:, None]
                nz = P_ref > 0
                err = float(np.abs(P[nz] / P_ref[nz] - 1).max())
                ends = np.minimum(np.arange(1, prefix.shape[1] + 1) * chunk, E) - 1
                cdf = np.cumsum(P_ref, axis=1)[:, ends]
                pre = prefix.astype(np.float64) / S[:, None]
                perr = np.abs(pre / cdf - 1)
                worst = np.unravel_index(np.argmax(perr), perr.shape)
                emit(kind="probe", case=case["name"], E=E, r=r, fat=bool(mp.enable_candidate_fatigue), n_fat=int(np.asarray(fat).sum()),
                     stop=bool(mp.enable_stop_candidate), chunk=int(chunk), score_err=err, prefix_err=float(perr.max()), worst_e_j=[int(worst[0]), int(worst[1])],
                     diag0=bool((np.diag(scores) == 0).all()), engine=eng.last_run_info()[0])
            except Exception as ex:      # noqa: BLE001
                emit(kind="probe", case=case["name"], r=r, error=repr(ex))


BASE = dict(CFG2, stickiness_factor=1.2, bandwagon_strength=1.0)
EPK = dict(max_rounds=10, supermajority_threshold=0.6, runoff_threshold_rounds=6, fatigue_threshold_rounds=2, fatigue_vote_share_threshold=0.04,
           fatigue_top_n_immune=2)


def variants():
    """name -> (ModelParams overrides, EngineParams overrides): each term of the EXT path alone and together."""
    return {
        "fat_only": (dict(enable_candidate_fatigue=True), dict(enable_candidate_fatigue=True)),
        "stop_only": (dict(enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.25, stop_candidate_threat_min_vote_share=0.1),
                      dict(enable_stop_candidate=True)),
        "both": (dict(enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.25,
                      stop_candidate_threat_min_vote_share=0.1), dict(enable_candidate_fatigue=True, enable_stop_candidate=True)),
        "both_k3": (dict(enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.25,
                         stop_candidate_threat_min_vote_share=0.1), dict(enable_candidate_fatigue=True, enable_stop_candidate=True, runoff_candidate_count=3)),
        "stop_negT": (dict(enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.25, stop_candidate_threat_min_vote_share=-0.5,
                           stop_candidate_boost_factor=0.5), dict(enable_stop_candidate=True)),
        "both_nosticky_nobw": (dict(stickiness_factor=0.0, bandwagon_strength=0.0, enable_candidate_fatigue=True, enable_stop_candidate=True,
                                    stop_candidate_threshold_unacceptable_distance=0.25, stop_candidate_threat_min_vote_share=0.1),
                               dict(enable_candidate_fatigue=True, enable_stop_candidate=True)),
        # 110 rounds without a run-off: 110 whole round tables do not fit in shared memory, every warp stages the round it plays
        # (Dense32Args::cta_tabs = 0) and folds the fatigue factors into that copy in place.  Nobody reaches the supermajority with
        # these terms live, so every trial plays all 110 rounds: the per-round tallies are what is compared
        "both_long_no_runoff": (dict(enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.25,
                                     stop_candidate_threat_min_vote_share=0.1, beta_increment_amount=0.05),
                                dict(enable_candidate_fatigue=True, enable_stop_candidate=True, max_rounds=110, runoff_candidate_count=0)),
    }


def trial_cases(eng, quick):
    from conclave_sim_b200 import FP32
    for E in ([135, 40] if quick else [135, 128, 40, 200, 2, 9, 255]):
        roster = O.synthetic_roster(E, seed=300 + E)
        upload(eng, roster)
        for name, (mkw, ekw) in variants().items():
            mp = O.ModelParams(**dict(BASE, **mkw))
            ep = O.EngineParams(**dict(EPK, **ekw))
            n = 256
            try:
                ref = CO.run(roster, [(mp, ep)], O.WIRING_MODEL, n, seed=8, want_tallies=True)
                res = {}
                for tag, legacy in (("ext", False), ("legacy", True)):
                    if legacy:
                        os.environ["CSIM_FULL_LEGACY"] = "1"
                    else:
                        os.environ.pop("CSIM_FULL_LEGACY", None)
                    out = eng.run([to_params(mp, ep)], n, seed=8, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32, want_round_tallies=True)
                    same = (out.winner == ref["winner"]) & (out.rounds == ref["rounds"])
                    rt, rr_ = out.round_tallies, ref["round_tallies"]
                    upto = [float(np.mean((rt[:, :q + 1, :] == rr_[:, :q + 1, :]).all(axis=(1, 2)))) for q in range(rt.shape[1])]
                    res[tag] = dict(disagree=int((~same).sum()), tallies_equal_through_round=[round(x, 3) for x in upto])
                os.environ.pop("CSIM_FULL_LEGACY", None)
                emit(kind="trials", E=E, variant=name, n=n, mean_rounds=float(ref["rounds"].mean()), **res)
            except Exception as ex:      # noqa: BLE001
                os.environ.pop("CSIM_FULL_LEGACY", None)
                emit(kind="trials", E=E, variant=name, error=repr(ex))


def timing(eng):
    from conclave_sim_b200 import FP32, make_params
    from bench import synthetic_roster, ENGINE_KW, CFG2 as BC2
    ideology, region, pap = synthetic_roster(135)
    eng.upload_roster(ideology, region, pap)
    n = 250_000
    for name, kw in (("fat+stop", dict(enable_candidate_fatigue=True, enable_stop_candidate=True)), ("fat", dict(enable_candidate_fatigue=True)),
                     ("stop", dict(enable_stop_candidate=True))):
        p = make_params(dict(BC2, stop_candidate_threshold_unacceptable_distance=0.3), **kw, **ENGINE_KW)
        for tag, legacy in (("ext", False), ("legacy", True)):
            if legacy:
                os.environ["CSIM_FULL_LEGACY"] = "1"
            else:
                os.environ.pop("CSIM_FULL_LEGACY", None)
            best = None
            for _ in range(3):
                out = eng.run([p], n, seed=5, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32)
                best = out.kernel_ms if best is None else min(best, out.kernel_ms)
            emit(kind="time", what=name, path=tag, trials=n, kernel_ms=best, elector_rounds_per_s=float(out.elector_rounds) / (best * 1e-3),
                 mean_rounds=float(out.rounds.mean()))
        os.environ.pop("CSIM_FULL_LEGACY", None)


def main():
    from conclave_sim_b200 import Engine
    CO.build()
    eng = Engine(0)
    t0 = time.time()
    probe_cases(eng)
    trial_cases(eng, quick="--quick" in sys.argv)
    if "--time" in sys.argv:
        timing(eng)
    emit(kind="done", seconds=time.time() - t0)
    eng.close()


if __name__ == "__main__":
    main()

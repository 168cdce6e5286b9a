"""north_star: "vote probabilities must agree within 1e-5 relative in FP32" — held here for the arithmetic the TRIAL kernels
actually use, not for the standalone csim_prob_matrix kernels (those are checked in test_gpu_parity.py).

csim_engine_probs runs one trial of the named engine through the PROBE instantiation of its own trial kernel (same source,
same __device__ routines, same tables) and dumps, for one full-field round, what every draw is made from:

  scores[e, c]   the per-candidate score the re-scan sums           -> P[e, c] = scores / S_e
  prefix[e, j]   the chunk prefixes the binary search probes        -> must equal cumsum(P)[chunk boundaries] * S_e
                 (S_e = prefix[e, -1], the row sum the target u * S_e is formed from)

Both are compared at rtol 1e-5 with the matrices the UNMODIFIED reference produced (tests/golden/model_real.npz,
model_cases.npz) and, for inputs the goldens do not hold, with the fp64 csim_prob_matrix that is itself pinned to them at 1e-13.
The table engine never forms single probabilities: what it searches is the fp32 CDF row, which is held to 1e-5 relative; the
differences of neighbouring CDF entries carry the CDF's absolute rounding (6e-8) and are checked at that absolute tolerance.
"""
import os

import numpy as np
import pytest

from oracle import conclave_oracle as O
import golden_util as G
from test_gpu_parity import to_params, upload, CFG2

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # north_star's tolerance
DENSE, TABLE, PREFIX, FULL = 1, 2, 3, 4
REF_CLI, MODEL = 0, 1


@pytest.fixture(scope="module")
def eng():
    from conclave_sim_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def check_against(eng, p, wiring, engine, r, P_ref, pv=None, pc=None, fat=None, what=""):
    """P_ref: fp64 [E, E] rows summing to 1 (reference).  Checks single probabilities and the probed chunk prefixes."""
    scores, prefix, chunk = eng.engine_probs(p, wiring=wiring, engine=engine, round=r, prev_vote=pv, prev_count=pc, fatigued=fat)
    E = P_ref.shape[0]
    S = prefix[:, -1].astype(np.float64)
    assert (S > 0).all(), what
    P = scores.astype(np.float64) / S[:, None]
    cdf_ref = np.cumsum(P_ref, axis=1)
    ends = np.minimum(np.arange(1, prefix.shape[1] + 1) * chunk, E) - 1           # last candidate of each chunk
    pre = prefix.astype(np.float64) / S[:, None]
    if engine == TABLE:
        assert np.allclose(pre, cdf_ref[:, ends], rtol=RTOL, atol=0), (what, np.abs(pre / cdf_ref[:, ends] - 1).max())
        assert np.allclose(P, P_ref, rtol=0, atol=2e-7), (what, np.abs(P - P_ref).max())
        return
    assert (np.diag(scores) == 0).all(), what                                   # no self vote (model.py:418-424)
    nz = P_ref > 0
    err = np.abs(P[nz] / P_ref[nz] - 1).max()
    assert err <= RTOL, (what, "single probabilities", err)
    assert (P[~nz] == 0).all(), what
    perr = np.abs(pre / cdf_ref[:, ends] - 1).max()
    assert perr <= RTOL, (what, "chunk prefixes", perr)


def test_real_roster_matrices_of_the_reference(eng, monkeypatch):
    """cfg-1 / cfg-2 rounds on the real 134-elector roster (model_real.npz, produced by the reference): every fp32 trial engine."""
    z = np.load(os.path.join(G.GOLD, "model_real.npz"))
    upload(eng, G.real_roster())
    from oracle.gen_golden import CFG1, CFG2 as C2
    n = 0
    for name, kw in (("cfg1", CFG1), ("cfg2", C2)):
        p = to_params(O.ModelParams(**kw))
        for key in z.files:
            if not key.startswith(name + "/P"):
                continue
            r = int(key.split("/P")[1])
            for engine in (DENSE, TABLE, PREFIX, FULL):
                check_against(eng, p, REF_CLI, engine, r, z[key], what=f"{name} round {r} engine {engine}")
                n += 1
            monkeypatch.setenv("CSIM_DENSE_MUFU", "1")                    # the MUFU.EX2 instantiation of the dense kernel
            check_against(eng, p, REF_CLI, DENSE, r, z[key], what=f"{name} round {r} dense/MUFU")
            monkeypatch.delenv("CSIM_DENSE_MUFU")
    assert n >= 8
    # `model` wiring: round 2 with the previous round's votes live (bandwagon + stickiness), reference matrix
    p = to_params(O.ModelParams(**C2))
    pv = z["cfg2_model/prev_vote"].astype(np.int32)
    pc = np.bincount(pv, minlength=134).astype(np.int32)
    for engine in (DENSE, PREFIX, FULL):
        check_against(eng, p, MODEL, engine, 2, z["cfg2_model/P2"], pv=pv, pc=pc, what=f"cfg2 model round 2 engine {engine}")


@pytest.mark.parametrize("case", [c for c in G.model_cases() if c["name"] != "c05"], ids=lambda c: c["name"])
def test_model_cases_of_the_reference(eng, case):
    """The 28 random model cases (4 rounds each, previous votes / fatigue masks / shares as the reference was given them).
    FULL takes every case as is; DENSE and PREFIX take the ones without fatigue / stop-candidate as is and the others with
    those two switched off (reference then = the fp64 csim_prob_matrix, pinned to the goldens at 1e-13 in test_gpu_parity)."""
    from conclave_sim_b200 import FP64
    mp, roster = case["mp"], case["roster"]
    upload(eng, roster)
    E = roster.E
    plain = not (mp.enable_candidate_fatigue or mp.enable_stop_candidate)
    p_full = to_params(mp)
    p_plain = to_params(O.ModelParams(**{**mp.__dict__, "enable_candidate_fatigue": False, "enable_stop_candidate": False}))
    for r in range(1, case["P"].shape[0] + 1):
        beta = case["beta"][r - 1]
        pv = None if r == 1 else case["prev_vote"][r - 1].astype(np.int32)
        pc = None if pv is None else np.bincount(pv[pv >= 0], minlength=E).astype(np.int32)
        fat = case["fatigued"][r - 1]
        shares_trial = None if pc is None else pc / float(E)          # what a trial hands the model (simulate.py:73)
        golden_ok = pc is None or np.array_equal(case["shares"][r - 1], shares_trial)
        # FULL: every term
        if golden_ok and not (r == 1 and mp.enable_stop_candidate and case["shares"][0].any()):
            P_ref = case["P"][r - 1]
        else:
            P_ref = eng.prob_matrix(p_full, beta, pv, pc, fat, shares_trial if shares_trial is not None else np.zeros(E), precision=FP64)
        check_against(eng, p_full, MODEL, FULL, r, P_ref, pv=pv, pc=pc, fat=fat, what=f"{case['name']} round {r} full")
        # DENSE / PREFIX: bandwagon + stickiness
        P_plain = case["P"][r - 1] if plain else eng.prob_matrix(p_plain, beta, pv, pc, precision=FP64)
        for engine in (DENSE, PREFIX):
            check_against(eng, p_plain, MODEL, engine, r, P_plain, pv=pv, pc=pc, what=f"{case['name']} round {r} engine {engine}")


@pytest.mark.parametrize("E", [135, 300, 1024])
def test_benchmarked_workload_probabilities(eng, E):
    """The rosters bench.py times (synthetic E = 135 headline, cfg-5's 1024) with cfg-2 parameters, rounds 1..5 (the full-field
    rounds of the benchmark): every engine against the fp64 matrix."""
    from conclave_sim_b200 import FP64
    roster = O.synthetic_roster(E)
    upload(eng, roster)
    p = to_params(O.ModelParams(**CFG2))
    rng = np.random.default_rng(E)
    for r in (1, 3, 5):
        beta = eng.beta_schedule(p, r, 1)[0]
        P_ref = eng.prob_matrix(p, beta, precision=FP64)
        for engine in ((DENSE, TABLE, PREFIX, FULL) if E <= 256 else (DENSE, PREFIX, FULL)):
            check_against(eng, p, REF_CLI, engine, r, P_ref, what=f"E={E} round {r} engine {engine}")
        pv = rng.integers(0, E, E).astype(np.int32)
        pv[pv == np.arange(E)] = (pv[pv == np.arange(E)] + 1) % E
        pc = np.bincount(pv, minlength=E).astype(np.int32)
        P_m = eng.prob_matrix(p, beta, pv, pc, precision=FP64)
        for engine in (DENSE, PREFIX, FULL):
            check_against(eng, p, MODEL, engine, r, P_m, pv=pv, pc=pc, what=f"E={E} round {r} model engine {engine}")


@pytest.mark.parametrize("E", [135, 200, 255])
def test_full_fast_path_probabilities_at_bench_size(eng, E, monkeypatch):
    """The FULL engine's fast path (EXT instantiation of the dense kernel) on rosters that fill whole Philox quads (the golden model
    cases are <= 22 electors: leftover-elector path only): cfg-2 parameters + candidate fatigue + stop-candidate as bench.py times
    them, a random fatigue mask, previous votes WITH threats (a few candidates above the threat share: the masked pair loop) and
    WITHOUT (every share below it: the phi-folded plain loop), against the fp64 matrix (pinned to the reference at 1e-13 in
    test_gpu_parity) — and the general kernel k_trials_full32 on the same inputs."""
    from conclave_sim_b200 import FP64
    roster = O.synthetic_roster(E)
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.3))
    p = to_params(mp)
    rng = np.random.default_rng(7000 + E)
    for r in (2, 5):
        beta = eng.beta_schedule(p, r, 1)[0]
        for threats in (True, False):
            if threats:                                    # three front-runners take 60 % of the votes
                front = rng.choice(E, 3, replace=False)
                pv = np.where(rng.random(E) < 0.6, front[rng.integers(0, 3, E)], rng.integers(0, E, E)).astype(np.int32)
            else:
                pv = rng.permutation(E).astype(np.int32)   # one vote each: every share is 1 / E < 0.15
            pv[pv == np.arange(E)] = (pv[pv == np.arange(E)] + 1) % E
            pc = np.bincount(pv, minlength=E).astype(np.int32)
            assert (pc / E > 0.15).any() == threats
            fat = rng.random(E) < 0.6
            P_ref = eng.prob_matrix(p, beta, pv, pc, fat, pc / float(E), precision=FP64)
            for legacy in (False, True):
                if legacy:
                    monkeypatch.setenv("CSIM_FULL_LEGACY", "1")
                else:
                    monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)
                check_against(eng, p, MODEL, FULL, r, P_ref, pv=pv, pc=pc, fat=fat,
                              what=f"E={E} round {r} threats={threats} {'k_trials_full32' if legacy else 'fast path'}")
            monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)


@pytest.mark.parametrize("case", list(G.extras_cases()), ids=lambda c: c["name"])
def test_fatigue_and_stop_matrices_of_the_reference_on_full_size_rosters(eng, case, monkeypatch):
    """The reference's OWN matrices with candidate fatigue + stop-candidate live on the real 134-elector roster and a 200-elector one
    (tests/golden/model_extras.npz, oracle/gen_golden_extras.py): csim_prob_matrix in fp64 at 1e-13, and what the FULL engine's two
    kernels (the fast path of dense32.cuh and k_trials_full32) draw from at north_star's 1e-5."""
    from conclave_sim_b200 import FP64
    mp, roster, r = case["mp"], case["roster"], case["round"]
    upload(eng, roster)
    E = roster.E
    p = to_params(mp)
    assert eng.beta_schedule(p, r, 1)[0] == case["beta"]
    pv = None if r == 1 else case["prev_vote"].astype(np.int32)
    pc = None if pv is None else np.bincount(pv, minlength=E).astype(np.int32)
    P64 = eng.prob_matrix(p, case["beta"], pv, pc, case["fatigued"], case["shares"], precision=FP64)
    assert np.allclose(P64, case["P"], rtol=1e-13, atol=0)
    for legacy in (False, True):
        if legacy:
            monkeypatch.setenv("CSIM_FULL_LEGACY", "1")
        else:
            monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)
        check_against(eng, p, MODEL, FULL, r, case["P"], pv=pv, pc=pc, fat=case["fatigued"],
                      what=f"{case['name']} {'k_trials_full32' if legacy else 'fast path'}")
    monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)


def test_probe_argument_errors(eng):
    from conclave_sim_b200._cabi import CsimError
    upload(eng, O.synthetic_roster(40))
    p = to_params(O.ModelParams(**CFG2))
    with pytest.raises(CsimError):
        eng.engine_probs(p, wiring=REF_CLI, engine=0, round=1)                      # AUTO is not an engine
    with pytest.raises(CsimError):
        eng.engine_probs(p, wiring=MODEL, engine=TABLE, round=1)                    # table: ref_cli only
    with pytest.raises(CsimError):
        eng.engine_probs(p, wiring=REF_CLI, engine=DENSE, round=0)

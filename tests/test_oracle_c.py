"""Pins the C restatement (oracle/conclave_oracle.c) against the reference's golden vectors and
against the NumPy restatement, including the native Philox stream."""
import numpy as np
import pytest

from oracle import conclave_oracle as O
from oracle import c_oracle as CO
import golden_util as G


@pytest.fixture(scope="module", autouse=True)
def _build():
    CO.build()


def test_pairwise_and_need_votes():
    rng = np.random.default_rng(3)
    for n in [1, 7, 8, 9, 127, 128, 129, 134, 1024, 3000]:
        a = rng.random(n)
        assert CO.pairwise_sum(a) == float(np.sum(a))
    for thr, E in [(0.56, 134), (2 / 3, 134), (2 / 3, 135), (0.5, 134), (0.0, 10), (1.0, 7), (0.8, 5)]:
        assert CO.need_votes(thr, E) == O.need_votes(thr, E)


def test_philox_stream_identical():
    for seed, sim, r, E in [(0, 0, 1, 5), (20250507, 123456789012, 17, 134), (2**63 + 5, 2**40 + 1, 50, 1024)]:
        assert np.array_equal(CO.philox_uniforms(seed, sim, r, E), O.philox_uniforms(seed, sim, r, E))


@pytest.mark.parametrize("case", list(G.model_cases()), ids=lambda c: c["name"])
def test_prob_matrix(case):
    mp, roster = case["mp"], case["roster"]
    for r in range(1, case["P"].shape[0] + 1):
        beta = CO.beta_schedule(mp, r, 1)[0]
        assert beta == case["beta"][r - 1]
        pv = case["prev_vote"][r - 1]
        P = CO.prob_matrix(roster, mp, beta, prev_vote=None if r == 1 else pv,
                           fatigued=case["fatigued"][r - 1], shares=case["shares"][r - 1])
        # libm exp vs numpy's SIMD exp may differ in the last ulp
        assert np.allclose(P, case["P"][r - 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("case", list(G.extras_cases()), ids=lambda c: c["name"])
def test_prob_matrix_with_fatigue_and_stop_on_full_size_rosters(case):
    """The reference's matrices with candidate fatigue + stop-candidate on the real roster and a 200-elector one (model_extras.npz)."""
    mp, roster, r = case["mp"], case["roster"], case["round"]
    assert CO.beta_schedule(mp, r, 1)[0] == case["beta"]
    P = CO.prob_matrix(roster, mp, case["beta"], prev_vote=None if r == 1 else case["prev_vote"], fatigued=case["fatigued"], shares=case["shares"])
    assert np.allclose(P, case["P"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("case", list(G.engine_cases()), ids=lambda c: c["name"])
def test_engine_under_reference_tapes(case):
    tapes = G.case_tapes(case)
    out = CO.run(case["roster"], [(case["mp"], case["ep"])], case["wiring"], case["n"], tapes=tapes, want_tallies=True)
    ok = ~case["fatigue_tie"]
    assert np.array_equal(out["winner"][ok], case["winners"][ok])
    assert np.array_equal(out["rounds"][ok], case["rounds"][ok])
    assert np.array_equal(out["reason"][ok], case["reason"][ok].astype(np.uint8))
    assert np.array_equal(out["round_tallies"][ok], case["tallies"][ok])
    for i in np.where(ok)[0]:
        assert np.array_equal(out["final_votes"][i], case["tallies"][i, case["rounds"][i] - 1])


@pytest.mark.parametrize("wiring", [O.WIRING_REF_CLI, O.WIRING_MODEL])
def test_native_rng_matches_numpy_oracle(wiring):
    from oracle.gen_golden import CFG2
    roster = G.real_roster()
    mp = O.ModelParams(**CFG2)
    ep = O.EngineParams(max_rounds=20, supermajority_threshold=0.56)
    n = 24 if wiring == O.WIRING_REF_CLI else 8
    ref = O.run_batch(roster, mp, ep, n, seed=99, sim_id_begin=1000, wiring=wiring)
    out = CO.run(roster, [(mp, ep)], wiring, n, seed=99, sim_id_begin=1000, want_tallies=True)
    for i, tr in enumerate(ref):
        assert out["winner"][i] == tr.winner and out["rounds"][i] == tr.rounds
        assert np.array_equal(out["round_tallies"][i, : tr.rounds], tr.tallies)
    assert out["totals"][0, 1] == n and out["totals"][0, 0] == sum(t.rounds for t in ref)
    assert out["winner_hist"].sum() == n


def test_histograms_and_param_sets():
    from oracle.gen_golden import CFG1, CFG2
    roster = G.real_roster()
    sets = [(O.ModelParams(**CFG2), O.EngineParams(max_rounds=20, supermajority_threshold=0.56)),
            (O.ModelParams(**CFG1), O.EngineParams(max_rounds=20, supermajority_threshold=2 / 3))]
    out = CO.run(roster, sets, O.WIRING_REF_CLI, 200, seed=5)
    w = out["winner"].reshape(2, 200); r = out["rounds"].reshape(2, 200)
    for s in range(2):
        assert out["winner_hist"][s, :134].sum() == (w[s] >= 0).sum()
        assert out["winner_hist"][s, 134] == (w[s] < 0).sum()
        assert np.array_equal(out["rounds_hist"][s], np.bincount(r[s][w[s] >= 0], minlength=21))
        assert out["totals"][s, 0] == r[s].sum()
    assert (w[1] < 0).all() and (r[1] == 20).all()          # cfg-1: nobody reaches 2/3
    # same sims in one set alone give the same results (independent of batching)
    solo = CO.run(roster, sets[:1], O.WIRING_REF_CLI, 50, seed=5, sim_id_begin=100)
    assert np.array_equal(solo["winner"], w[0][100:150]) and np.array_equal(solo["rounds"], r[0][100:150])


def test_faithful_mode_gives_identical_results():
    from oracle.gen_golden import CFG2
    roster = G.real_roster()
    sets = [(O.ModelParams(**CFG2), O.EngineParams(max_rounds=20, supermajority_threshold=0.56))]
    a = CO.run(roster, sets, O.WIRING_REF_CLI, 60, seed=3, want_tallies=True)
    CO.set_faithful(True)
    try:
        b = CO.run(roster, sets, O.WIRING_REF_CLI, 60, seed=3, want_tallies=True)
    finally:
        CO.set_faithful(False)
    assert np.array_equal(a["winner"], b["winner"]) and np.array_equal(a["round_tallies"], b["round_tallies"])

"""Pins the CPU oracle (oracle/conclave_oracle.py) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py), and against the closed-form known answers the
reference's own tests/test_model.py states (papabile :552-632, regional :635-705, bandwagon
:708-792, stickiness :795-874, combined :879-1000, dynamic beta :1006-1079)."""
import os

import numpy as np
import pytest

from oracle import conclave_oracle as O
import golden_util as G


def test_pairwise_sum_is_numpys():
    rng = np.random.default_rng(0)
    for n in [1, 5, 7, 8, 9, 64, 127, 128, 129, 134, 135, 200, 256, 257, 1000, 1024, 2049]:
        a = rng.random(n)
        assert O.pairwise_sum(a) == float(np.sum(a))


@pytest.mark.parametrize("case", list(G.model_cases()), ids=lambda c: c["name"])
def test_prob_matrix_matches_reference_model(case):
    mp, roster = case["mp"], case["roster"]
    beta = float(mp.initial_beta_weight)
    for r in range(1, case["P"].shape[0] + 1):
        beta = O.beta_step(mp, beta, r)
        assert beta == case["beta"][r - 1]
        pv = case["prev_vote"][r - 1]
        P = O.prob_matrix(roster, mp, beta, prev_vote=None if r == 1 else pv,
                          fatigued=case["fatigued"][r - 1], shares=case["shares"][r - 1])
        assert np.array_equal(P, case["P"][r - 1])


def test_prob_matrix_real_roster():
    import os
    z = np.load(os.path.join(G.GOLD, "model_real.npz"))
    roster = G.real_roster()
    from oracle.gen_golden import CFG1, CFG2
    for name, kw in (("cfg1", CFG1), ("cfg2", CFG2)):
        mp = O.ModelParams(**kw)
        betas = O.beta_schedule(mp, 20)
        for key in z.files:
            if key.startswith(name + "/P"):
                r = int(key.split("/P")[1])
                assert betas[r - 1] == float(z[f"{name}/beta{r}"])
                assert np.array_equal(O.prob_matrix(roster, mp, betas[r - 1]), z[key])
    mp = O.ModelParams(**CFG2)
    P = O.prob_matrix(roster, mp, float(z["cfg2_model/beta2"]), prev_vote=z["cfg2_model/prev_vote"])
    assert np.array_equal(P, z["cfg2_model/P2"])


@pytest.mark.parametrize("case", list(G.extras_cases()), ids=lambda c: c["name"])
def test_prob_matrix_with_fatigue_and_stop_on_full_size_rosters(case):
    """Candidate fatigue + stop-candidate on the real 134-elector roster and a 200-elector one (oracle/gen_golden_extras.py): the
    reference's own matrices, bit for bit, and its beta schedule."""
    mp, roster, r = case["mp"], case["roster"], case["round"]
    assert O.beta_schedule(mp, r)[r - 1] == case["beta"]
    P = O.prob_matrix(roster, mp, case["beta"], prev_vote=None if r == 1 else case["prev_vote"], fatigued=case["fatigued"], shares=case["shares"])
    assert np.array_equal(P, case["P"])
    if r > 1:      # the cases are meant to exercise both branches of the stop-candidate rule
        assert (case["shares"] > mp.stop_candidate_threat_min_vote_share).any() == (r != 3)


@pytest.mark.parametrize("case", list(G.engine_cases()), ids=lambda c: c["name"])
def test_engine_matches_reference_worker(case):
    tapes = G.case_tapes(case)
    res = O.run_batch(case["roster"], case["mp"], case["ep"], case["n"], wiring=case["wiring"], tapes=tapes)
    E = case["roster"].E
    checked = 0
    for i, tr in enumerate(res):
        if case["fatigue_tie"][i]:
            # reference's immune set depended on numpy's unstable argsort (simulate.py:94); not a defined behaviour
            assert tr.fatigue_tie
            continue
        checked += 1
        assert tr.winner == case["winners"][i]
        assert tr.rounds == case["rounds"][i]
        assert tr.reason == case["reason"][i]
        assert tr.rounds * E == case["draws_used"][i]
        assert np.array_equal(tr.tallies, case["tallies"][i, : tr.rounds])
        assert np.all(case["tallies"][i, tr.rounds:] == -1)
    assert checked > 0 or case["name"] == "m_cfg2_fatigue"


# ---- closed-form known answers, as the reference's tests/test_model.py states them ----
def _kat_roster():
    return O.Roster(np.array([0.1, 0.5, 0.9]), np.array([0, 1, 0], np.int32), np.array([True, False, False]))


def _norm(s):
    s = s.copy()
    np.fill_diagonal(s, 0.0)
    return s / s.sum(axis=1, keepdims=True)


def test_kat_papabile_multiplicative():
    ro = _kat_roster()
    mp = O.ModelParams(initial_beta_weight=1.0, papabile_weight_factor=2.0)
    s = np.exp(-np.abs(ro.ideology[:, None] - ro.ideology[None, :]))
    s[:, 0] *= 2.0
    assert np.allclose(O.prob_matrix(ro, mp, 1.0), _norm(s), rtol=1e-12)


def test_kat_regional_additive():
    ro = _kat_roster()
    mp = O.ModelParams(initial_beta_weight=1.0, regional_affinity_bonus=0.25)
    s = np.exp(-np.abs(ro.ideology[:, None] - ro.ideology[None, :]))
    s[0, 2] += 0.25; s[2, 0] += 0.25; s[0, 0] += 0.25; s[1, 1] += 0.25; s[2, 2] += 0.25
    assert np.allclose(O.prob_matrix(ro, mp, 1.0), _norm(s), rtol=1e-12)


def test_kat_bandwagon_and_stickiness_order():
    ro = _kat_roster()
    mp = O.ModelParams(initial_beta_weight=1.0, bandwagon_strength=0.5, stickiness_factor=0.5,
                       papabile_weight_factor=1.5, regional_affinity_bonus=0.1)
    prev = np.array([1, 0, 1], np.int32)          # counts: c0=1, c1=2 -> bonuses 0.25, 0.5, 0
    s = np.exp(-np.abs(ro.ideology[:, None] - ro.ideology[None, :]))
    s[:, 0] *= 1.5
    s[ro.region[:, None] == ro.region[None, :]] += 0.1
    s += np.array([0.25, 0.5, 0.0])[None, :]
    for e in range(3):
        s[e, prev[e]] *= 1.5
    assert np.allclose(O.prob_matrix(ro, mp, 1.0, prev_vote=prev), _norm(s), rtol=1e-12)
    # zero stickiness + no previous votes == first round (test_model.py:305-343)
    mp0 = O.ModelParams(initial_beta_weight=1.0)
    assert np.array_equal(O.prob_matrix(ro, mp0, 1.0, prev_vote=None), O.prob_matrix(ro, mp0, 1.0, prev_vote=np.full(3, -1)))


def test_kat_dynamic_beta():
    mp = O.ModelParams(initial_beta_weight=1.0, enable_dynamic_beta=True, beta_growth_rate=1.1)
    assert np.allclose(O.beta_schedule(mp, 3), [1.1, 1.21, 1.331], rtol=1e-15)
    mp = O.ModelParams(initial_beta_weight=1.5, enable_dynamic_beta=True, beta_increment_amount=0.3, beta_increment_interval_rounds=1)
    assert O.beta_schedule(mp, 2)[0] == 1.5 + 0.3                      # round 1 already incremented
    mp = O.ModelParams(initial_beta_weight=1.0, enable_dynamic_beta=True, beta_increment_amount=0.5, beta_increment_interval_rounds=3)
    assert list(O.beta_schedule(mp, 6)) == [1.0, 1.0, 1.5, 1.5, 1.5, 2.0]
    mp = O.ModelParams(initial_beta_weight=1.0, enable_dynamic_beta=False, beta_increment_amount=0.5)
    assert list(O.beta_schedule(mp, 3)) == [1.0, 1.0, 1.0]


def test_properties_rows_sum_to_one_zero_diag():
    ro = G.real_roster()
    P = O.prob_matrix(ro, O.ModelParams(initial_beta_weight=2.0, regional_affinity_bonus=0.1, papabile_weight_factor=1.5), 2.0)
    assert P.shape == (134, 134) and (P >= 0).all()
    assert np.allclose(P.sum(axis=1), 1.0, rtol=1e-12)
    assert (np.diag(P) == 0).all()


def test_need_votes_and_topk():
    assert O.need_votes(0.56, 134) == 76 and O.need_votes(2 / 3, 134) == 90 and O.need_votes(2 / 3, 135) == 90
    assert O.need_votes(0.5, 134) == 67 and O.need_votes(0.0, 10) == 0
    assert list(O.top_k(np.array([3, 5, 5, 1, 5]), 2)) == [1, 2]
    assert list(O.top_k(np.array([0, 0, 0]), 2)) == [0, 1]


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    out = O.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(w[0]) for w in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = O.philox4x32_10([0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(w[0]) for w in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = O.philox4x32_10([0x243F6A88], [0x85A308D3], [0x13198A2E], [0x03707344], 0xA4093822, 0x299F31D0)
    assert [int(w[0]) for w in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = O.philox_uniforms(7, 3, 1, 9)
    assert u.shape == (9,) and (u >= 0).all() and (u < 1).all()


@pytest.mark.parametrize("case", list(G.subset_cases()), ids=lambda c: c["name"])
def test_prob_matrix_candidate_subset_matches_reference(case):
    """effective_candidate_ids (model.py:280-296)."""
    P = O.prob_matrix(case["roster"], case["mp"], case["beta"], prev_vote=case["prev_vote"], fatigued=case["fatigued"],
                      shares=case["shares"], cand_index=case["cand"])
    assert P.shape == case["P"].shape and np.array_equal(P, case["P"])


def test_oracle_aggregate_equals_the_reference_main_json():
    """oracle.aggregate (src/simulate.py:356-387) against the stats JSON the unmodified reference's main() wrote
    (tests/golden/main, oracle/gen_golden_main.py)."""
    import json
    import pandas as pd
    gold = os.path.join(G.GOLD, "main")
    with open(os.path.join(gold, "main_cases.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 4
    for c in cases:
        with open(os.path.join(gold, c["name"] + ".json")) as f:
            want = json.load(f)
        if c["n"]:
            ref = pd.read_csv(os.path.join(gold, c["name"] + ".csv"), dtype={"winner": str})
            winners = [None if pd.isna(w) else w for w in ref["winner"]]
            rounds = ref["rounds"].tolist()
        else:
            winners, rounds = [], []
        got = O.aggregate(winners, rounds, c["n"])
        assert json.loads(json.dumps(got)) == want, c["name"]
        assert list(got) == list(want) and list(got["winner_counts"]) == list(want["winner_counts"])

"""GPU parity tests of the PREFIX engine (CSIM_ENGINE_PREFIX, conclave_sim_b200/csrc/prefix.cuh), through the C ABI.

The engine is fp32: the bar is the one the other fp32 engines meet — per-trial agreement with the fp64 oracle
under the shared Philox stream (a draw flips only when u lands within fp32 rounding of a CDF boundary), the
chi-square / KS gate against the reference's own sample, and size-independent properties at full size.
Its tables are built in fp64 and rounded once, so in practice it agrees with the oracle more often than the
dense fp32 engine does.
"""
import json
import os

import numpy as np
import pytest

from oracle import conclave_oracle as O
from oracle import c_oracle as CO
import golden_util as G
from parity_util import assert_agree
from test_gpu_parity import CFG2, EP2, to_params, upload, _chi2_homogeneity

pytestmark = pytest.mark.gpu

DENSE, TABLE, PREFIX = 1, 2, 3


@pytest.fixture(scope="module")
def eng():
    from conclave_sim_b200 import Engine
    CO.build()
    e = Engine(0)
    yield e
    e.close()


def _same(out, ref):
    return (out.winner == ref["winner"]) & (out.rounds == ref["rounds"]) & (out.final_votes == ref["final_votes"]).all(axis=1)


@pytest.mark.parametrize("wiring,n", [(O.WIRING_REF_CLI, 20000), (O.WIRING_MODEL, 6000)])
def test_prefix_agrees_with_oracle_per_trial(eng, wiring, n):
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**CFG2)
    out = eng.run([to_params(mp, EP2)], n, seed=5, wiring=wiring, engine=PREFIX, precision=FP32, want_final_votes=True)
    ref = CO.run(roster, [(mp, EP2)], wiring, n, seed=5)
    same = _same(out, ref)
    assert_agree(same, "prefix", ref["rounds"].sum() * 134, what=f"prefix wiring {wiring}")
    assert np.array_equal(out.winner_hist[0, :134], np.bincount(out.winner[out.winner >= 0], minlength=134))
    assert out.totals[0, 0] == out.rounds.sum() and out.totals[0, 1] == n


@pytest.mark.parametrize("kw", [
    dict(stickiness_factor=0.0), dict(bandwagon_strength=0.0), dict(stickiness_factor=2.5, bandwagon_strength=0.0),
    dict(regional_affinity_bonus=0.0, papabile_weight_factor=1.0), dict(enable_dynamic_beta=False, initial_beta_weight=0.0),
    dict(initial_beta_weight=-1.0, enable_dynamic_beta=False), dict(initial_beta_weight=10.0, beta_increment_amount=5.0),
    dict(bandwagon_strength=5.0, stickiness_factor=0.1),
], ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_prefix_model_terms_one_by_one(eng, kw):
    """Each term of the `model` wiring on and off (and betas outside the factorised-exponential range of dense32)."""
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, **kw))
    ep = O.EngineParams(max_rounds=12, supermajority_threshold=0.56, runoff_threshold_rounds=6)
    n = 1500
    ref = CO.run(roster, [(mp, ep)], O.WIRING_MODEL, n, seed=17)
    out = eng.run([to_params(mp, ep)], n, seed=17, wiring=O.WIRING_MODEL, engine=PREFIX, precision=FP32, want_final_votes=True)
    same = _same(out, ref)
    assert_agree(same, "prefix", ref["rounds"].sum() * 134, what=f"prefix model terms {kw}")


def test_prefix_fp32_underflow_matches_dense32(eng):
    """beta = 40 + 20 per round: exp(-beta |dx|) underflows fp32 in the run-off rounds (both run-off scores flush to zero ->
    the uniform fallback of simulate.py:151), which fp64 does not do — a property of fp32, not of the engine: the
    full-field rounds still follow the oracle, and the two fp32 engines agree with each other per trial."""
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, initial_beta_weight=40.0, beta_increment_amount=20.0))
    ep = O.EngineParams(max_rounds=12, supermajority_threshold=0.56, runoff_threshold_rounds=6)
    n = 1500
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=17, want_tallies=True)
        pf = eng.run([to_params(mp, ep)], n, seed=17, wiring=wiring, engine=PREFIX, precision=FP32, want_round_tallies=True)
        dn = eng.run([to_params(mp, ep)], n, seed=17, wiring=wiring, engine=DENSE, precision=FP32, want_round_tallies=True)
        for r in range(6):
            assert np.mean((pf.round_tallies[:, r, :] == ref["round_tallies"][:, r, :]).all(axis=1)) > 0.99
        assert_agree((pf.winner == dn.winner) & (pf.rounds == dn.rounds), "dense", pf.rounds.sum() * 134, what=f"underflow prefix vs dense32 wiring {wiring}", slack=2.0)


@pytest.mark.parametrize("E", [1, 2, 3, 9, 33, 127, 128, 129, 136, 137, 193, 200, 255, 256, 257, 300, 513])
def test_prefix_ragged_rosters(eng, E):
    """Roster sizes around every boundary of the lane dealing: 32-lane passes, 128-elector quad groups, a tail of more /
    fewer than 64 electors (193: the tail is dealt as quads), 256 (chunks of 8 -> 16) and 512 (-> 32)."""
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(E, seed=100 + E)
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, stickiness_factor=1.5, bandwagon_strength=1.0))
    ep = O.EngineParams(max_rounds=9, supermajority_threshold=0.6, runoff_threshold_rounds=4)
    n = 128 if E > 256 else 256
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=8, want_tallies=True)
        out = eng.run([to_params(mp, ep)], n, seed=8, wiring=wiring, engine=PREFIX, precision=FP32, want_round_tallies=True,
                      want_final_votes=True)
        played = out.round_tallies[:, 0, :]
        assert (played.sum(axis=1) == E).all()                        # every elector votes exactly once per round
        live = out.round_tallies.reshape(n, -1, E).sum(axis=2)
        for i in range(0, n, 17):
            assert (live[i, :out.rounds[i]] == E).all() and (out.round_tallies[i, out.rounds[i]:] == -1).all()
        assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), "prefix", ref["rounds"].sum() * E, what=f"prefix ragged E={E} wiring {wiring}")


@pytest.mark.parametrize("ep_kw", [
    dict(max_rounds=0), dict(max_rounds=1), dict(max_rounds=6, runoff_threshold_rounds=0), dict(max_rounds=6, runoff_threshold_rounds=-3),
    dict(max_rounds=6, runoff_candidate_count=500), dict(max_rounds=6, runoff_candidate_count=1),
    dict(max_rounds=5, supermajority_threshold=0.0), dict(max_rounds=5, supermajority_threshold=1.5),
    dict(max_rounds=7, runoff_threshold_rounds=100), dict(max_rounds=9, runoff_threshold_rounds=2, runoff_candidate_count=3),
    dict(max_rounds=30, runoff_candidate_count=0, supermajority_threshold=0.9),
], ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_prefix_edge_engine_parameters(eng, ep_kw):
    """Degenerate worker arguments the reference accepts (simulate.py:30-41)."""
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(37, seed=11)
    upload(eng, roster)
    mp = O.ModelParams(**CFG2)
    kw = dict(max_rounds=6, supermajority_threshold=0.56)
    kw.update(ep_kw)
    ep = O.EngineParams(**kw)
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, 96, seed=2, want_tallies=True)
        out = eng.run([to_params(mp, ep)], 96, seed=2, wiring=wiring, engine=PREFIX, precision=FP32, want_final_votes=True,
                      want_round_tallies=True)
        assert np.array_equal(out.reason, (out.winner >= 0).astype(np.uint8))
        assert_agree((out.rounds == ref["rounds"]) & (out.winner == ref["winner"]), "prefix", max(1, ref["rounds"].sum()) * roster.E, what=f"prefix edge parameters {ep_kw} wiring {wiring}")
        if ep.max_rounds == 0:
            assert (out.final_votes == 0).all() and (out.rounds == 0).all() and (out.winner == -1).all()
        else:
            assert np.mean((out.round_tallies == ref["round_tallies"]).all(axis=(1, 2))) > 0.95    # vote-level: a moved vote that leaves the outcome alone still shows here
    assert eng.run([to_params(mp, ep)], 0, seed=2, engine=PREFIX).winner.shape == (0,)


def test_prefix_heterogeneous_sets_and_wide_sim_ids(eng):
    """One batch mixing run-off widths / start rounds / thresholds (so CTAs re-stage their tables per set);
    sim ids beyond 2^32."""
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    eps = [O.EngineParams(max_rounds=12, supermajority_threshold=0.56, runoff_threshold_rounds=5, runoff_candidate_count=2),
           O.EngineParams(max_rounds=12, supermajority_threshold=0.50, runoff_threshold_rounds=3, runoff_candidate_count=3),
           O.EngineParams(max_rounds=12, supermajority_threshold=0.60, runoff_threshold_rounds=4, runoff_candidate_count=0),
           O.EngineParams(max_rounds=12, supermajority_threshold=0.45, runoff_threshold_rounds=1, runoff_candidate_count=5)]
    mps = [O.ModelParams(**dict(CFG2, papabile_weight_factor=pw, regional_affinity_bonus=rb, initial_beta_weight=b0))
           for pw, rb, b0 in ((2.0, 0.1, 1.5), (1.0, 0.0, 0.5), (3.0, 0.3, 2.0), (1.5, 0.05, 1.0))]
    sets = list(zip(mps, eps))
    base = 2**40 + 5
    n = 300
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, sets, wiring, n, seed=2**63 + 11, sim_id_begin=base)
        out = eng.run([to_params(*s) for s in sets], n, seed=2**63 + 11, sim_id_begin=base, wiring=wiring, engine=PREFIX,
                      precision=FP32, want_final_votes=True)
        same = _same(out, ref)
        assert_agree(same, "prefix", ref["rounds"].sum() * 134, what=f"prefix heterogeneous sets wiring {wiring}")
        w = out.winner.reshape(4, n); r = out.rounds.reshape(4, n)
        for s in range(4):
            assert out.winner_hist[s, :134].sum() == (w[s] >= 0).sum() and out.winner_hist[s, 134] == (w[s] < 0).sum()
            assert np.array_equal(out.rounds_hist[s], np.bincount(r[s][w[s] >= 0], minlength=13))
            assert out.totals[s, 0] == r[s].sum() and out.totals[s, 1] == n


def test_prefix_sweep_many_sets_and_sharding(eng):
    """A 96-set sweep (BASELINE cfg-4 in miniature): every set's trials are the ones a single-set run gives, and
    a shard of the sim-id range reproduces the whole (what multi-GPU sharding relies on)."""
    import itertools
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(135)
    upload(eng, roster)
    grid = list(itertools.product([0.5, 1.0, 1.5, 2.0], [0.0, 0.3], [1.0, 1.5, 3.0], [0.5, 0.56, 0.6, 2 / 3]))
    sets = [(O.ModelParams(**dict(CFG2, initial_beta_weight=b0, beta_increment_amount=db, papabile_weight_factor=pw)),
             O.EngineParams(max_rounds=20, supermajority_threshold=thr)) for b0, db, pw, thr in grid]
    ps = [to_params(*s) for s in sets]
    n = 700
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        full = eng.run(ps, n, seed=4, wiring=wiring, engine=PREFIX, precision=FP32)
        w = full.winner.reshape(len(ps), n); r = full.rounds.reshape(len(ps), n)
        for s in (0, 37, 95):
            solo = eng.run([ps[s]], n, seed=4, wiring=wiring, engine=PREFIX, precision=FP32)
            assert np.array_equal(solo.winner, w[s]) and np.array_equal(solo.rounds, r[s])
            assert np.array_equal(solo.winner_hist[0], full.winner_hist[s]) and np.array_equal(solo.totals[0], full.totals[s])
            part = eng.run([ps[s]], 250, seed=4, sim_id_begin=300, wiring=wiring, engine=PREFIX, precision=FP32)
            assert np.array_equal(part.winner, w[s][300:550]) and np.array_equal(part.rounds, r[s][300:550])
        ref = CO.run(roster, [sets[37]], wiring, n, seed=4)
        assert_agree((w[37] == ref["winner"]) & (r[37] == ref["rounds"]), "prefix", ref["rounds"].sum() * roster.E, what=f"prefix 96-set sweep wiring {wiring}")


def test_prefix_tables_in_global_memory(eng):
    """Tables read through L2 instead of shared memory: forced on a small roster, natural on cfg-5's 1024 x 1024."""
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**CFG2)
    p = [to_params(mp, EP2)]
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        staged = eng.run(p, 3000, seed=9, wiring=wiring, engine=PREFIX, precision=FP32, want_final_votes=True)
        os.environ["CSIM_PREFIX_NOSTAGE"] = "1"
        try:
            glob = eng.run(p, 3000, seed=9, wiring=wiring, engine=PREFIX, precision=FP32, want_final_votes=True)
        finally:
            del os.environ["CSIM_PREFIX_NOSTAGE"]
        assert np.array_equal(staged.winner, glob.winner) and np.array_equal(staged.final_votes, glob.final_votes)
    roster = O.synthetic_roster(1024)
    upload(eng, roster)
    ep = O.EngineParams(max_rounds=50, supermajority_threshold=0.56)
    p = [to_params(mp, ep)]
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        n = 4000
        out = eng.run(p, n, seed=5, wiring=wiring, engine=PREFIX, precision=FP32, want_final_votes=True)
        d = eng.run(p, n, seed=5, wiring=wiring, engine=DENSE, precision=FP32, want_final_votes=True)
        assert (out.final_votes.sum(axis=1) == 1024).all()
        assert_agree((out.winner == d.winner) & (out.rounds == d.rounds), "dense", out.totals[0, 0] * 1024, what=f"cfg5 prefix vs dense32 wiring {wiring}", slack=2.0)
        ref = CO.run(roster, [(mp, ep)], wiring, 48, seed=5)
        assert_agree((out.winner[:48] == ref["winner"]) & (out.rounds[:48] == ref["rounds"]), "prefix", ref["rounds"].sum() * 1024, what=f"cfg5 prefix vs oracle wiring {wiring}")


def test_prefix_distribution_matches_reference_sample(eng):
    """north_star's statistical gate (chi-square winners / rounds, KS rounds, p > 0.01) against the reference's own
    20 000-trial CPU sample, 1e6 GPU trials."""
    from scipy import stats
    from conclave_sim_b200 import FP32
    with open(os.path.join(G.GOLD, "dist_cfg2.json")) as f:
        ref = json.load(f)
    roster = G.real_roster()
    upload(eng, roster)
    n = 1_000_000
    out = eng.run([to_params(O.ModelParams(**CFG2), EP2)], n, seed=987654321, precision=FP32, engine=PREFIX, want_reason=False)
    gw = np.concatenate([out.winner_hist[0, :134], [out.winner_hist[0, 134]]])
    rw = np.concatenate([ref["winner_hist"], [ref["no_winner"]]])
    p_w = _chi2_homogeneity(gw, rw)
    gr = np.concatenate([out.rounds_hist[0], [out.winner_hist[0, 134]]])
    rr = np.concatenate([ref["rounds_hist_success"], [ref["no_winner"]]])
    p_r = _chi2_homogeneity(gr, rr)
    p_ks = stats.ks_2samp(np.repeat(np.arange(21), out.rounds_hist[0]), np.repeat(np.arange(21), ref["rounds_hist_success"])).pvalue
    print(f"[prefix] chi2 winners p={p_w:.4f}  chi2 rounds p={p_r:.4f}  KS rounds p={p_ks:.4f}")
    assert p_w > 0.01 and p_r > 0.01 and p_ks > 0.01


def test_prefix_full_size_properties_model_wiring(eng):
    """1.25 M trials (BASELINE cfg-3's per-GPU share) under the `model` wiring: conservation, independence from how the
    sim-id range is cut, and per-trial agreement with the dense fp32 kernel (a different algorithm on the same stream)."""
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(135)
    upload(eng, roster)
    p = [to_params(O.ModelParams(**CFG2), EP2)]
    n = 1_250_000
    for wiring in (O.WIRING_MODEL, O.WIRING_REF_CLI):
        pf = eng.run(p, n, seed=31337, wiring=wiring, engine=PREFIX, precision=FP32)
        dn = eng.run(p, n, seed=31337, wiring=wiring, engine=DENSE, precision=FP32)
        w, rd = pf.winner, pf.rounds
        assert pf.winner_hist.sum() == n and pf.totals[0, 1] == n and pf.totals[0, 0] == rd.astype(np.int64).sum()
        assert np.array_equal(pf.winner_hist[0], np.bincount(np.where(w >= 0, w, 135), minlength=136))
        assert np.array_equal(pf.rounds_hist[0], np.bincount(rd[w >= 0], minlength=21))
        assert ((w >= 0) | (rd == 20)).all() and (rd >= 1).all() and (rd <= 20).all()
        third = n // 3
        a = eng.run(p, third, seed=31337, sim_id_begin=0, wiring=wiring, engine=PREFIX, precision=FP32)
        b = eng.run(p, n - third, seed=31337, sim_id_begin=third, wiring=wiring, engine=PREFIX, precision=FP32)
        assert np.array_equal(np.concatenate([a.winner, b.winner]), w) and np.array_equal(np.concatenate([a.rounds, b.rounds]), rd)
        assert np.array_equal(a.winner_hist + b.winner_hist, pf.winner_hist)
        same = np.mean((pf.winner == dn.winner) & (pf.rounds == dn.rounds))
        print(f"prefix==dense on {same:.5f} of {n} trials (wiring {wiring})")
        assert same > 0.9995                                          # measured 0.9998 (dense32's own rate)
        assert _chi2_homogeneity(pf.winner_hist[0], dn.winner_hist[0]) > 1e-3


def test_prefix_model_wiring_distribution_vs_independent_oracle_sample(eng):
    """`model` wiring under the native RNG: winners / rounds of 1e6 GPU trials against an INDEPENDENTLY seeded sample of
    the fp64 oracle (40 000 trials, other seed, other sim ids) — the analogue, for the wiring the reference's CLI never
    reaches, of the chi-square / KS gate against the reference's own sample."""
    from scipy import stats
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**CFG2)
    ref = CO.run(roster, [(mp, EP2)], O.WIRING_MODEL, 40000, seed=424242, sim_id_begin=7_000_000)
    n = 1_000_000
    out = eng.run([to_params(mp, EP2)], n, seed=99, wiring=O.WIRING_MODEL, precision=FP32, engine=PREFIX, want_reason=False)
    p_w = _chi2_homogeneity(out.winner_hist[0], ref["winner_hist"][0])
    gr = np.append(out.rounds_hist[0], out.winner_hist[0, 134])
    rr = np.append(ref["rounds_hist"][0], ref["winner_hist"][0, 134])
    p_r = _chi2_homogeneity(gr, rr)
    p_ks = stats.ks_2samp(np.repeat(np.arange(21), out.rounds_hist[0]), np.repeat(np.arange(21), ref["rounds_hist"][0])).pvalue
    print(f"[prefix/model] chi2 winners p={p_w:.4f}  chi2 rounds p={p_r:.4f}  KS rounds p={p_ks:.4f}")
    assert p_w > 0.01 and p_r > 0.01 and p_ks > 0.01


def test_plain_c_client_matches_python(eng, tmp_path):
    """The C ABI without Python in the loop: tests/c/c_client.c (C99, only include/conclave_b200.h) is compiled with gcc,
    linked against the library and run as its own process; its per-trial output equals the Python host's for the same
    roster, parameters, seed and sim ids — on every engine — and an unsupported request comes back as a status + message."""
    import subprocess
    from conclave_sim_b200 import FP32, FP64, make_params
    from conclave_sim_b200._cabi import LIB_PATH
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "c_client"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", "c_client.c"), LIB_PATH, "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-o", str(exe)],
                   check=True)
    E, n, seed = 135, 3000, 20250507
    s = 12345
    x = np.zeros(E); g = np.zeros(E, np.int32); pap = np.zeros(E, np.uint8)
    for i in range(E):                                   # the LCG of c_client.c
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF; x[i] = (s >> 8) / 16777216.0
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF; g[i] = (s >> 16) % 8
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF; pap[i] = ((s >> 16) % 100) < 15
    eng.upload_roster(x, g, pap)
    p = make_params(CFG2, max_rounds=20, supermajority_threshold=0.56)
    for engine, wiring, precision in ((PREFIX, 1, FP32), (PREFIX, 0, FP32), (DENSE, 1, FP32), (TABLE, 0, FP32), (TABLE, 0, FP64), (0, 1, FP32)):
        r = subprocess.run([str(exe), str(E), str(n), str(engine), str(wiring), str(precision), str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().split("\n")
        got = np.array([[int(v) for v in ln.split()] for ln in lines[:n]])
        ref = eng.run([p], n, seed=seed, sim_id_begin=1000, wiring=wiring, engine=engine, precision=precision)
        assert np.array_equal(got[:, 0], ref.winner) and np.array_equal(got[:, 1], ref.rounds), (engine, wiring, precision)
        assert lines[n] == f"totals {int(ref.totals[0, 0])} {int(ref.totals[0, 1])}"
        assert lines[n + 1] == "whist " + " ".join(str(int(v)) for v in ref.winner_hist[0])
        assert lines[n + 2] == "rhist " + " ".join(str(int(v)) for v in ref.rounds_hist[0])
        assert lines[n + 3].startswith("prefix+fp64 -> -4 (") and "fp32 engine" in lines[n + 3]


def test_cli_model_wiring_runs_on_the_prefix_engine(eng, tmp_path):
    """`python -m conclave_sim_b200.simulate --wiring model` (bandwagon + stickiness live): the reference's CLI surface, CSV and
    stats JSON contracts, on the engine AUTO picks for this wiring; the statistics match the oracle's for the same wiring."""
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    roster = G.real_roster()
    df = pd.DataFrame({"ideology_score": roster.ideology, "region": [f"R{g}" for g in roster.region],
                       "is_papabile": roster.papabile}, index=pd.Index(roster.ids, name="elector_id"))
    csv = tmp_path / "roster.csv"
    df.reset_index().to_csv(csv, index=False)
    out_csv, out_json = tmp_path / "res.csv", tmp_path / "stats.json"
    common = ["-n", "4000", "-r", "20", "-t", "0.56", "-f", str(csv), "-o", str(out_csv), "-s", str(out_json),
              "--enable-dynamic-beta", "--beta-increment-amount", "0.3", "--beta-increment-interval-rounds", "1",
              "--initial-beta-weight", "1.5", "--bandwagon-strength", "0.7", "--papabile-weight-factor", "2.0",
              "--stickiness-factor", "0.8", "--seed", "42", "--wiring", "model"]
    res = {}
    for engine in ("auto", "prefix", "dense"):
        S.main(common + ["--engine", engine])
        got = pd.read_csv(out_csv)
        assert list(got.columns) == ["sim_id", "winner", "rounds", "reason"] and len(got) == 4000
        res[engine] = (got, json.load(open(out_json)))
    assert res["auto"][0].equals(res["prefix"][0])                       # AUTO == PREFIX for this wiring
    same = (res["prefix"][0]["rounds"] == res["dense"][0]["rounds"]).mean()
    assert same > 0.99                                                    # two fp32 kernels, one Philox stream
    ref = CO.run(roster, [(O.ModelParams(**CFG2), EP2)], O.WIRING_MODEL, 4000, seed=42)
    js = res["prefix"][1]
    assert js["total_simulations"] == 4000
    assert abs(js["success_rate"] - float((ref["winner"] >= 0).mean())) < 0.005
    assert abs(js["average_rounds_all"] - float(ref["rounds"].mean())) < 0.02
    S.RUNTIME.update(wiring="ref_cli", engine="auto", seed=None)


# ------------------------------------------------------------------ the FULL engine (fatigue + stop-candidate in fp32)
FULL = 4


def _extras_sets():
    base = dict(CFG2)
    mps = [O.ModelParams(**base, enable_candidate_fatigue=True, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.3),
           O.ModelParams(**base, enable_candidate_fatigue=True, fatigue_penalty_factor=0.25),
           O.ModelParams(**base, enable_stop_candidate=True, stop_candidate_threshold_unacceptable_distance=0.2,
                         stop_candidate_threat_min_vote_share=0.08, stop_candidate_boost_factor=2.5),
           O.ModelParams(**base)]
    eps = [O.EngineParams(max_rounds=14, supermajority_threshold=0.56, runoff_threshold_rounds=7, enable_candidate_fatigue=True,
                          fatigue_threshold_rounds=2, fatigue_vote_share_threshold=0.03, fatigue_top_n_immune=3, enable_stop_candidate=True),
           O.EngineParams(max_rounds=14, supermajority_threshold=0.6, runoff_threshold_rounds=9, runoff_candidate_count=3,
                          enable_candidate_fatigue=True, fatigue_threshold_rounds=3, fatigue_vote_share_threshold=0.05, fatigue_top_n_immune=0),
           O.EngineParams(max_rounds=14, supermajority_threshold=0.56, runoff_threshold_rounds=5, enable_stop_candidate=True),
           O.EngineParams(max_rounds=14, supermajority_threshold=0.56, runoff_threshold_rounds=5)]
    return list(zip(mps, eps))


def test_full_engine_fatigue_and_stop_vs_oracle(eng):
    """fp32 with candidate fatigue and stop-candidate live: per-trial agreement with the fp64 oracle (the fatigue set and the
    stop trigger are decided in fp64 on both sides, so a trial diverges only through an fp32 draw flip), per-round tallies,
    histograms; the same batch through AUTO (which must pick this engine) and against the exact fp64 GPU engine."""
    from conclave_sim_b200 import FP32, FP64
    roster = G.real_roster()
    upload(eng, roster)
    sets = _extras_sets()
    ps = [to_params(*s) for s in sets]
    n = 600
    ref = CO.run(roster, sets, O.WIRING_MODEL, n, seed=3, want_tallies=True)
    out = eng.run(ps, n, seed=3, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32, want_round_tallies=True, want_final_votes=True)
    same = _same(out, ref).reshape(len(sets), n).mean(axis=1)
    print("full32 vs oracle per set:", same)
    for si in range(len(sets)):
        assert_agree(_same(out, ref).reshape(len(sets), n)[si], "full", ref["rounds"].reshape(len(sets), n)[si].sum() * 134, what=f"full32 fatigue+stop set {si}")
    first = out.round_tallies[:, 0, :]
    assert (first.sum(axis=1) == 134).all()
    assert np.mean((out.round_tallies[:, :3, :] == ref["round_tallies"][:, :3, :]).all(axis=(1, 2))) > 0.99
    w = out.winner.reshape(len(sets), n)
    for s in range(len(sets)):
        assert out.winner_hist[s, :134].sum() == (w[s] >= 0).sum() and out.totals[s, 1] == n
    auto = eng.run(ps, n, seed=3, wiring=O.WIRING_MODEL, precision=FP32)              # AUTO: fatigue / stop live -> FULL
    assert np.array_equal(auto.winner, out.winner) and np.array_equal(auto.rounds, out.rounds)
    exact = eng.run(ps, n, seed=3, wiring=O.WIRING_MODEL, precision=FP64)
    assert np.array_equal(exact.winner, ref["winner"]) and np.array_equal(exact.rounds, ref["rounds"])
    # the extras change the outcome distribution: the test would be vacuous otherwise
    plain = eng.run([ps[3]] * 1, n, seed=3, wiring=O.WIRING_MODEL, engine=PREFIX, precision=FP32)
    # set 3 has no extras: FULL follows PREFIX trial for trial up to fp32 boundary draws (the batch runs on the factorised dense kernel)
    assert_agree(plain.winner == w[3], "full", ref["rounds"].reshape(len(sets), n)[3].sum() * 134, what="full32 set without extras vs prefix")
    assert np.mean(w[0] != w[3]) > 0.2 and np.mean(w[2] != w[3]) > 0.05


@pytest.mark.parametrize("E", [2, 9, 33, 129, 200, 300])
def test_full_engine_on_ragged_rosters(eng, E):
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(E, seed=300 + E)
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, stickiness_factor=1.2, bandwagon_strength=1.0), enable_candidate_fatigue=True, enable_stop_candidate=True,
                       stop_candidate_threshold_unacceptable_distance=0.25, stop_candidate_threat_min_vote_share=0.1)
    ep = O.EngineParams(max_rounds=10, supermajority_threshold=0.6, runoff_threshold_rounds=6, enable_candidate_fatigue=True,
                        fatigue_threshold_rounds=2, fatigue_vote_share_threshold=0.04, fatigue_top_n_immune=2, enable_stop_candidate=True)
    n = 128
    ref = CO.run(roster, [(mp, ep)], O.WIRING_MODEL, n, seed=8, want_tallies=True)
    out = eng.run([to_params(mp, ep)], n, seed=8, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32, want_round_tallies=True)
    assert (out.round_tallies[:, 0, :].sum(axis=1) == E).all()
    assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), "full", ref["rounds"].sum() * E, what=f"full32 ragged E={E}")


@pytest.mark.parametrize("E", [135, 255])
def test_full_fast_path_term_by_term(eng, E, monkeypatch):
    """The FULL engine's fast path (EXT instantiation of the dense kernel, dense32.cuh) term by term — fatigue only, stop-candidate only,
    both, a k = 3 run-off, a negative threat share (every candidate a threat from round 1), no stickiness / bandwagon — against the fp64
    oracle, and the legacy k_trials_full32 (CSIM_FULL_LEGACY=1) on the same batch: both follow the oracle per trial."""
    from conclave_sim_b200 import FP32
    import ext_check as X
    roster = O.synthetic_roster(E, seed=300 + E)
    upload(eng, roster)
    n = 300
    for name, (mkw, ekw) in X.variants().items():
        mp = O.ModelParams(**dict(X.BASE, **mkw))
        ep = O.EngineParams(**dict(X.EPK, **ekw))
        ref = CO.run(roster, [(mp, ep)], O.WIRING_MODEL, n, seed=8, want_tallies=True)
        for legacy in (False, True):
            if legacy:
                monkeypatch.setenv("CSIM_FULL_LEGACY", "1")
            else:
                monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)
            out = eng.run([to_params(mp, ep)], n, seed=8, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32, want_round_tallies=True)
            assert (out.round_tallies[:, 0, :].sum(axis=1) == E).all()
            assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), "full", ref["rounds"].sum() * E,
                         what=f"full32 {'legacy' if legacy else 'fast path'} {name} E={E}")
            if name == "both_long_no_runoff":              # 110 rounds each, no winner: the tallies carry the comparison
                assert (ref["rounds"] == 110).all()
                eq = (out.round_tallies == ref["round_tallies"])
                # (a moved vote feeds back through bandwagon / stickiness, so a trial's tallies stay different once an fp32 boundary
                # draw has flipped: measured 0.95 / 0.74 of the trials identical through 20 / 110 rounds at E = 255, on both kernels;
                # a wrong table or factor leaves none)
                eq20, eq_all = eq[:, :20, :].all(axis=(1, 2)).mean(), eq.all(axis=(1, 2)).mean()
                print(f"full32 {'legacy' if legacy else 'fast path'} {name} E={E}: tallies identical through 20 / 110 rounds in {eq20:.3f} / {eq_all:.3f} of the trials")
                assert eq20 >= 0.90 and eq_all >= 0.55, (name, legacy, eq20, eq_all)
        monkeypatch.delenv("CSIM_FULL_LEGACY", raising=False)


def test_full_engine_without_extras_equals_prefix_and_scales(eng):
    """No fatigue / stop: the FULL kernel is a plain dense fp32 engine; on 200 000 trials it agrees with PREFIX per trial
    (different kernels, one Philox stream), conserves votes and is independent of how the sim-id range is cut."""
    from conclave_sim_b200 import FP32
    roster = O.synthetic_roster(135)
    upload(eng, roster)
    p = [to_params(O.ModelParams(**CFG2), EP2)]
    n = 200_000
    for wiring in (O.WIRING_MODEL, O.WIRING_REF_CLI):
        f = eng.run(p, n, seed=77, wiring=wiring, engine=FULL, precision=FP32)
        q = eng.run(p, n, seed=77, wiring=wiring, engine=PREFIX, precision=FP32)
        assert_agree((f.winner == q.winner) & (f.rounds == q.rounds), "full", f.totals[0, 0] * 135, what=f"full32 vs prefix wiring {wiring}")
        assert f.winner_hist.sum() == n and f.totals[0, 0] == f.rounds.astype(np.int64).sum()
        a_ = eng.run(p, 70_000, seed=77, sim_id_begin=0, wiring=wiring, engine=FULL, precision=FP32)
        b_ = eng.run(p, n - 70_000, seed=77, sim_id_begin=70_000, wiring=wiring, engine=FULL, precision=FP32)
        assert np.array_equal(np.concatenate([a_.winner, b_.winner]), f.winner)
        print(f"full32 wiring {wiring}: {n} trials in {f.kernel_ms:.2f} ms = {f.elector_rounds / (f.kernel_ms * 1e-3):.3e} elector-rounds/s")


def test_many_regions(eng):
    """More than 64 distinct regions (a roster keyed by country, say): the two-phase dense fp32 kernel does not take it,
    the PREFIX / FULL engines (and AUTO, which must route around DENSE) do; checked against the oracle."""
    from conclave_sim_b200 import FP32, CsimError
    E = 120
    base = O.synthetic_roster(E, seed=5)
    region = (np.arange(E) * 7 % 97).astype(np.int32)            # 97 distinct codes among 120 electors
    roster = O.Roster.from_arrays(base.ideology, region, base.papabile)
    assert len(np.unique(roster.region)) == 97
    upload(eng, roster)
    mp = O.ModelParams(**CFG2)
    ep = O.EngineParams(max_rounds=12, supermajority_threshold=0.56)
    p = [to_params(mp, ep)]
    n = 1000
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=6)
        for en in (PREFIX, FULL, 0):
            out = eng.run(p, n, seed=6, wiring=wiring, engine=en, precision=FP32)
            assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), "full" if en == FULL else "prefix", ref["rounds"].sum() * roster.E, what=f"many regions wiring {wiring} engine {en}")
        # round 2: the dense kernel's regional rows are floats (no 64-bit region masks any more): any number of regions that fits
        out = eng.run(p, n, seed=6, wiring=wiring, engine=DENSE, precision=FP32)
        assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), "dense", ref["rounds"].sum() * roster.E, what=f"many regions wiring {wiring} dense")


@pytest.mark.parametrize("kw", [
    dict(regional_affinity_bonus=-0.05), dict(bandwagon_strength=-0.4), dict(papabile_weight_factor=-1.0, regional_affinity_bonus=0.3),
    dict(regional_affinity_bonus=-0.2, bandwagon_strength=-0.3, enable_candidate_fatigue=True, fatigue_penalty_factor=-0.5),
], ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_negative_factors_go_to_the_full_engine(eng, kw):
    """Negative bonus / bandwagon / papabile / penalty factors make scores negative before the model's clamps (model.py:369,427):
    the table-based engines refuse them, AUTO sends them to FULL, which follows the fp64 oracle."""
    from conclave_sim_b200 import FP32, CsimError
    roster = G.real_roster()
    upload(eng, roster)
    mp = O.ModelParams(**dict(CFG2, **kw))
    ep = O.EngineParams(max_rounds=12, supermajority_threshold=0.56, enable_candidate_fatigue=bool(kw.get("enable_candidate_fatigue")),
                        fatigue_threshold_rounds=2)
    n = 800
    for wiring in (O.WIRING_MODEL, O.WIRING_REF_CLI):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=12)
        out = eng.run([to_params(mp, ep)], n, seed=12, wiring=wiring, precision=FP32)          # AUTO
        full = eng.run([to_params(mp, ep)], n, seed=12, wiring=wiring, engine=FULL, precision=FP32)
        assert np.array_equal(out.winner, full.winner) and np.array_equal(out.rounds, full.rounds)
        assert_agree((full.winner == ref["winner"]) & (full.rounds == ref["rounds"]), "full", ref["rounds"].sum() * roster.E, what=f"negative factors {kw} wiring {wiring}")
        with pytest.raises(CsimError):
            eng.run([to_params(mp, ep)], n, seed=12, wiring=wiring, engine=PREFIX, precision=FP32)


def test_pinned_result_pool(eng):
    """Large batches get their per-trial arrays in page-locked memory (csim_host_alloc); blocks are recycled once the arrays
    are gone, live results never share memory, and the values equal those of a small (pageable) call."""
    import gc
    from conclave_sim_b200 import FP32
    roster = G.real_roster()
    upload(eng, roster)
    p = [to_params(O.ModelParams(**CFG2), EP2)]
    n = 200_000
    a = eng.run(p, n, seed=1, engine=PREFIX, precision=FP32)
    b = eng.run(p, n, seed=2, engine=PREFIX, precision=FP32)
    assert a.winner.ctypes.data != b.winner.ctypes.data and not np.array_equal(a.winner, b.winner)
    small = eng.run(p, 1000, seed=1, engine=PREFIX, precision=FP32)             # pageable path
    assert np.array_equal(small.winner, a.winner[:1000]) and np.array_equal(small.rounds, a.rounds[:1000])
    addr = a.winner.ctypes.data
    keep = a.winner.copy()
    del a
    gc.collect()
    c = eng.run(p, n, seed=1, engine=PREFIX, precision=FP32)                     # reuses a released block
    assert np.array_equal(c.winner, keep)
    assert c.winner.ctypes.data in (addr, c.winner.ctypes.data) and c.winner.flags.writeable
    assert np.array_equal(b.winner, eng.run(p, n, seed=2, engine=PREFIX, precision=FP32).winner)   # b was never overwritten


def test_c_abi_error_paths_report_and_leave_the_context_usable(eng):
    """Every refusal is a status + csim_last_error() message (the reference logs a failed worker and silently drops its trial,
    simulate.py:350-351; nothing is dropped here), and the context keeps working afterwards."""
    from conclave_sim_b200 import FP32, FP64, CsimError, make_params
    roster = G.real_roster()
    upload(eng, roster)
    p = to_params(O.ModelParams(**CFG2), EP2)
    cases = [
        (dict(engine=9), "bad engine"),
        (dict(precision=7), "bad precision"),
        (dict(wiring=5), "bad wiring"),
        (dict(engine=PREFIX, precision=FP64), "fp32 engine"),
        (dict(engine=FULL, precision=FP64), "fp32 engine"),
        (dict(engine=TABLE, wiring=O.WIRING_MODEL), "ref_cli wiring only"),
        (dict(precision=FP32, tapes=np.zeros((10, 20 * 134))), "tape requires"),
    ]
    for kw, needle in cases:
        with pytest.raises(CsimError) as ei:
            eng.run([p], 10, seed=1, **kw)
        assert needle in str(ei.value), (kw, str(ei.value))
        ok = eng.run([p], 10, seed=1)                                  # the context survives every refusal
        assert ok.winner.shape == (10,)
    other = make_params(CFG2, max_rounds=7, supermajority_threshold=0.56)
    with pytest.raises(CsimError) as ei:
        eng.run([p, other], 10, seed=1)
    assert "share max_rounds" in str(ei.value)
    with pytest.raises(ValueError):
        eng.run([p], 10, seed=1, precision=FP64, tapes=np.zeros((10, 5)))   # host-side shape check
    with pytest.raises(TypeError):
        make_params({"not_a_model_parameter": 1})


@pytest.mark.parametrize("case_seed", list(range(int(os.environ.get("CSIM_FUZZ_CASES", "48")))))   # CSIM_FUZZ_CASES=N: a longer campaign
def test_fuzz_engines_vs_oracle(eng, case_seed):
    """Seeded differential fuzzing: random roster (1..220 electors, 1..E regions), random parameters of every kind (negative
    factors, fatigue, stop-candidate, run-off widths / start rounds, thresholds, beta schedules), both wirings.  The fp64
    engines must equal the C oracle exactly; every fp32 engine that accepts the case must agree with it per trial up to the
    draws fp32 rounding can flip, and conserve votes."""
    from conclave_sim_b200 import FP32, FP64, CsimError
    big = os.environ.get("CSIM_FUZZ_BIG", "0") == "1"       # CSIM_FUZZ_BIG=1: rosters past 256 / 512 electors (the prefix engine's
    rng = np.random.default_rng((500000 if big else 1000) + case_seed)     # 16-wide chunks and global-memory tables), fewer trials each
    E = int(rng.choice([257, 300, 511, 513, 640, 1024] if big else [1, 2, 3, 5, 17, 31, 32, 33, 64, 100, 127, 128, 129, 134, 135, 160, 220]))
    base = O.synthetic_roster(E, seed=int(rng.integers(1 << 30)))
    G_ = int(rng.integers(1, min(E, 80) + 1))
    region = rng.integers(0, G_, E).astype(np.int32)
    roster = O.Roster.from_arrays(base.ideology, region, base.papabile)
    upload(eng, roster)
    neg = rng.random() < 0.25
    extras = rng.random() < 0.4
    mp = O.ModelParams(
        initial_beta_weight=float(rng.choice([0.0, 0.5, 1.0, 1.5, 3.0, 8.0])), enable_dynamic_beta=bool(rng.random() < 0.7),
        beta_increment_amount=(None if rng.random() < 0.2 else float(rng.choice([0.0, 0.1, 0.3, 1.0]))),
        beta_growth_rate=float(rng.choice([1.0, 1.05, 1.3])), beta_increment_interval_rounds=int(rng.choice([0, 1, 2, 3])),
        stickiness_factor=float(rng.choice([0.0, 0.3, 0.8, 2.0])), bandwagon_strength=float(rng.choice([0.0, 0.7, 2.0]) * (-0.5 if neg and rng.random() < 0.5 else 1)),
        regional_affinity_bonus=float(rng.choice([0.0, 0.1, 0.5]) * (-1 if neg and rng.random() < 0.5 else 1)),
        papabile_weight_factor=float(rng.choice([1.0, 1.5, 2.0, 4.0])),
        enable_candidate_fatigue=extras and bool(rng.random() < 0.7), fatigue_penalty_factor=float(rng.choice([0.0, 0.25, 0.5, 0.9])),
        enable_stop_candidate=extras and bool(rng.random() < 0.7),
        stop_candidate_threshold_unacceptable_distance=float(rng.choice([0.1, 0.2, 0.3, 0.6])),
        stop_candidate_threat_min_vote_share=float(rng.choice([0.0, 0.05, 0.1, 0.15])), stop_candidate_boost_factor=float(rng.choice([1.0, 1.5, 3.0])))
    ep = O.EngineParams(
        max_rounds=int(rng.choice([1, 3, 8, 14])), supermajority_threshold=float(rng.choice([0.4, 0.5, 0.56, 2 / 3, 0.9])),
        runoff_threshold_rounds=int(rng.choice([0, 1, 3, 5, 100])), runoff_candidate_count=int(rng.choice([0, 1, 2, 2, 3, 7])),
        enable_candidate_fatigue=mp.enable_candidate_fatigue, fatigue_threshold_rounds=int(rng.choice([1, 2, 3])),
        fatigue_vote_share_threshold=float(rng.choice([0.02, 0.05, 0.1])), fatigue_top_n_immune=int(rng.choice([0, 1, 3])),
        enable_stop_candidate=mp.enable_stop_candidate)
    p = [to_params(mp, ep)]
    n = 24 if big else 96
    seed = int(rng.integers(1 << 40))
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=seed, want_tallies=True)
        ex = eng.run(p, n, seed=seed, wiring=wiring, precision=FP64, want_round_tallies=True)
        tie = False
        if wiring == O.WIRING_MODEL and mp.enable_candidate_fatigue and ep.fatigue_top_n_immune > 0:
            tie = True      # the immune top-n under ties is the one behaviour without a defined reference: same rule both sides
        assert np.array_equal(ex.winner, ref["winner"]) and np.array_equal(ex.rounds, ref["rounds"]), (case_seed, wiring, tie)
        assert np.array_equal(ex.round_tallies, ref["round_tallies"]), (case_seed, wiring)
        # the same under an injected uniform tape (the mode the reference's own draws are replayed in), dense and table engines
        nt = 12
        tape = rng.random((nt, ep.max_rounds * E))
        ref_t = CO.run(roster, [(mp, ep)], wiring, nt, tapes=tape, want_tallies=True)
        for en in ((DENSE, TABLE) if wiring == O.WIRING_REF_CLI else (DENSE,)):
            ex_t = eng.run(p, nt, wiring=wiring, engine=en, precision=FP64, tapes=tape, want_round_tallies=True)
            assert np.array_equal(ex_t.winner, ref_t["winner"]) and np.array_equal(ex_t.round_tallies, ref_t["round_tallies"]), (case_seed, wiring, en)
        for en in (0, PREFIX, FULL, DENSE, TABLE):
            try:
                out = eng.run(p, n, seed=seed, wiring=wiring, engine=en, precision=FP32, want_round_tallies=True)
            except CsimError:
                # AUTO takes everything; FULL everything whose region-bonus rows fit in shared memory (not 1024 electors x 60 regions)
                assert en in (PREFIX, DENSE, TABLE) or (big and en == FULL and E * G_ > 40000), (case_seed, wiring, en)
                continue
            played = out.round_tallies[:, 0, :]
            assert (played.sum(axis=1) == E).all(), (case_seed, wiring, en)
            same = (out.winner == ref["winner"]) & (out.rounds == ref["rounds"])
            if out.precision == "fp64":                               # AUTO fell through to the exact engine: no tolerance at all
                assert same.all(), (case_seed, wiring, en)
            else:
                # random parameters include knife-edge elections (beta = 0: every candidate equally likely), where one moved vote
                # changes the outcome far more often than on cfg-2: four times the engine's cfg-2 rate, never "most trials"
                assert_agree(same, out.engine, max(1, ref["rounds"].sum()) * E, what=f"fuzz {case_seed} wiring {wiring} engine {en}", slack=4.0)


def test_auto_falls_through_to_an_engine_that_fits(eng):
    """1024 electors x 62 regions with candidate fatigue live: PREFIX refuses fatigue, FULL's bonus rows (E x G floats) do not fit
    in shared memory, the fp32 dense engine has no fatigue -- AUTO must still answer, from the fp64 exact engine, and an
    explicitly named FULL must say why it cannot."""
    from conclave_sim_b200 import FP32, FP64, CsimError
    E, G_ = 1024, 62
    base = O.synthetic_roster(E, seed=5)
    region = (np.arange(E) % G_).astype(np.int32)
    roster = O.Roster.from_arrays(base.ideology, region, base.papabile)
    upload(eng, roster)
    mp = O.ModelParams(enable_candidate_fatigue=True, fatigue_penalty_factor=0.5, bandwagon_strength=0.7, stickiness_factor=0.3)
    ep = O.EngineParams(max_rounds=4, enable_candidate_fatigue=True, fatigue_threshold_rounds=1, fatigue_vote_share_threshold=0.05)
    p = [to_params(mp, ep)]
    with pytest.raises(CsimError) as ei:
        eng.run(p, 8, seed=3, wiring=O.WIRING_MODEL, engine=FULL, precision=FP32)
    assert "does not fit in shared memory" in str(ei.value)
    auto = eng.run(p, 8, seed=3, wiring=O.WIRING_MODEL, precision=FP32, want_round_tallies=True)
    exact = eng.run(p, 8, seed=3, wiring=O.WIRING_MODEL, precision=FP64, want_round_tallies=True)
    ref = CO.run(roster, [(mp, ep)], O.WIRING_MODEL, 8, seed=3, want_tallies=True)
    assert np.array_equal(auto.winner, exact.winner) and np.array_equal(auto.round_tallies, exact.round_tallies)
    assert np.array_equal(exact.winner, ref["winner"]) and np.array_equal(exact.round_tallies, ref["round_tallies"])
    # ref_cli, same roster: fatigue is inert there and AUTO's first choice runs
    out = eng.run(p, 8, seed=3, wiring=O.WIRING_REF_CLI, precision=FP32)
    assert out.winner.shape == (8,)


@pytest.mark.parametrize("E", [2048, 2500])
def test_rosters_at_and_past_the_prefix_limit(eng, E):
    """2048 electors is the PREFIX engine's limit (128 chunks of 16); 2500 is past it, where AUTO hands `model` trials to FULL
    and `ref_cli` trials to whatever fits.  Votes are conserved, fp64 equals the oracle exactly, fp32 agrees per trial."""
    from conclave_sim_b200 import FP32, FP64
    roster = O.synthetic_roster(E, seed=E)
    upload(eng, roster)
    mp = O.ModelParams(initial_beta_weight=2.0, bandwagon_strength=0.7, stickiness_factor=0.3, regional_affinity_bonus=0.1)
    ep = O.EngineParams(max_rounds=4, supermajority_threshold=0.5, runoff_threshold_rounds=2, runoff_candidate_count=2)
    p = [to_params(mp, ep)]
    n = 6
    for wiring in (O.WIRING_REF_CLI, O.WIRING_MODEL):
        ref = CO.run(roster, [(mp, ep)], wiring, n, seed=11, want_tallies=True)
        ex = eng.run(p, n, seed=11, wiring=wiring, precision=FP64, want_round_tallies=True)
        assert np.array_equal(ex.winner, ref["winner"]) and np.array_equal(ex.round_tallies, ref["round_tallies"])
        out = eng.run(p, n, seed=11, wiring=wiring, precision=FP32, want_round_tallies=True)
        assert (out.round_tallies[:, 0, :].sum(axis=1) == E).all()
        if out.precision == "fp64":
            assert np.array_equal(out.winner, ref["winner"]) and np.array_equal(out.rounds, ref["rounds"])
        else:
            assert_agree((out.winner == ref["winner"]) & (out.rounds == ref["rounds"]), out.engine, ref["rounds"].sum() * E, what=f"auto fall-through wiring {wiring}")

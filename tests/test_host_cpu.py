"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute call is made without a GPU), the ctypes struct mirrors match the header,
the host-only entry points (beta schedule, vote threshold) agree with the oracle, the CLI /
aggregation / writers keep the reference's contract, and the product path FAILS LOUDLY
without a CUDA device."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import conclave_oracle as O
import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "conclave_b200.h")


@pytest.fixture(scope="module")
def lib():
    from conclave_sim_b200 import _build, _cabi
    _build.build()
    return _cabi.load_library()


def test_library_exports_every_declared_symbol(lib):
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(csim_[a-z_0-9]+)\s*\(", src))
    from conclave_sim_b200._cabi import EXPORTS
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.csim_abi_version() == 2


def test_struct_mirrors_match_header(tmp_path):
    """Compile a tiny C program printing sizeof/offsetof of the header's structs; compare with ctypes."""
    from conclave_sim_b200._cabi import CsimParams, CsimRunArgs, CsimPrefixPlan, CsimFullPlan
    from oracle.c_oracle import CParams
    fields_p = [f[0] for f in CsimParams._fields_]
    fields_r = [f[0] for f in CsimRunArgs._fields_]
    fields_q = [f[0] for f in CsimPrefixPlan._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "conclave_b200.h"\nint main(){\n'
    prog += 'printf("%zu\\n", sizeof(csim_params));\n'
    for f in fields_p:
        prog += f'printf("%zu\\n", offsetof(csim_params, {f}));\n'
    prog += 'printf("%zu\\n", sizeof(csim_run_args));\n'
    for f in fields_r:
        prog += f'printf("%zu\\n", offsetof(csim_run_args, {f}));\n'
    prog += 'printf("%zu\\n", sizeof(csim_prefix_plan));\n'
    for f in fields_q:
        prog += f'printf("%zu\\n", offsetof(csim_prefix_plan, {f}));\n'
    fields_f = [f[0] for f in CsimFullPlan._fields_]
    prog += 'printf("%zu\\n", sizeof(csim_full_plan));\n'
    for f in fields_f:
        prog += f'printf("%zu\\n", offsetof(csim_full_plan, {f}));\n'
    prog += "return 0;}\n"
    c = tmp_path / "t.c"
    c.write_text(prog)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    nums = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    exp = [ctypes.sizeof(CsimParams)] + [getattr(CsimParams, f).offset for f in fields_p] + \
          [ctypes.sizeof(CsimRunArgs)] + [getattr(CsimRunArgs, f).offset for f in fields_r] + \
          [ctypes.sizeof(CsimPrefixPlan)] + [getattr(CsimPrefixPlan, f).offset for f in fields_q] + \
          [ctypes.sizeof(CsimFullPlan)] + [getattr(CsimFullPlan, f).offset for f in fields_f]
    assert nums == exp
    assert ctypes.sizeof(CParams) == ctypes.sizeof(CsimParams)
    assert [f[0] for f in CParams._fields_] == fields_p


def test_host_entry_points_match_oracle(lib):
    from conclave_sim_b200 import make_params
    for thr, E in [(0.56, 134), (2 / 3, 134), (2 / 3, 135), (0.5, 134), (0.0, 10), (1.0, 7), (0.8, 5), (0.56, 1024)]:
        assert lib.csim_need_votes(thr, E) == O.need_votes(thr, E)
    cases = [dict(initial_beta_weight=1.5, enable_dynamic_beta=True, beta_increment_amount=0.3, beta_increment_interval_rounds=1),
             dict(initial_beta_weight=1.0, enable_dynamic_beta=True, beta_increment_amount=0.5, beta_increment_interval_rounds=3),
             dict(initial_beta_weight=2.0, enable_dynamic_beta=True, beta_growth_rate=1.2),
             dict(initial_beta_weight=1.0, enable_dynamic_beta=False, beta_increment_amount=0.05, beta_increment_interval_rounds=10),
             dict(initial_beta_weight=1.0, enable_dynamic_beta=True, beta_increment_amount=0.1, beta_increment_interval_rounds=0)]
    for kw in cases:
        p = make_params(kw, max_rounds=50)
        out = np.zeros(50)
        assert lib.csim_beta_schedule(ctypes.byref(p), 1, 50, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(out, O.beta_schedule(O.ModelParams(**kw), 50))
        out2 = np.zeros(5)
        lib.csim_beta_schedule(ctypes.byref(p), 7, 5, out2.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(out2, out[6:11])


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conclave_sim_b200 import Engine, CsimError
    with pytest.raises(CsimError) as ei:
        Engine(0)
    assert "no CPU fallback" in str(ei.value)
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    df = pd.DataFrame({"ideology_score": [0.1, 0.5], "region": ["a", "b"], "is_papabile": [True, False]},
                      index=pd.Index(["0", "1"], name="elector_id"))
    with pytest.raises(CsimError):
        S._run_single_simulation_worker(0, df, {}, 5, 0.5, 5, 2, False, False, 3, 0.05, 3, False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "conclave_sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "conclave_oracle" not in txt, f


def test_cli_surface_matches_reference():
    from conclave_sim_b200 import simulate as S
    a = S.parse_args([])
    assert (a.num_simulations, a.max_rounds, a.runoff_rounds, a.runoff_candidates) == (100, 100, 5, 2)
    assert a.supermajority_threshold == 2 / 3 and a.elector_data_file == "merged_electors.csv"
    assert a.output_results == "data/simulation_results.csv" and a.output_stats == "data/simulation_stats.json"
    assert (a.initial_beta_weight, a.beta_increment_amount, a.beta_increment_interval_rounds) == (1.0, 0.05, 10.0)
    assert (a.stickiness_factor, a.bandwagon_strength, a.regional_bonus, a.papabile_weight_factor) == (0.5, 0.0, 0.1, 1.5)
    assert (a.fatigue_threshold_rounds, a.fatigue_vote_share_threshold, a.fatigue_penalty_factor, a.fatigue_top_n_immune) == (3, 0.05, 0.5, 3)
    assert (a.stop_candidate_threshold_unacceptable_distance, a.stop_candidate_threat_min_vote_share, a.stop_candidate_boost_factor) == (0.5, 0.15, 1.5)
    assert not (a.enable_dynamic_beta or a.enable_candidate_fatigue or a.enable_stop_candidate or a.verbose) and a.workers is None
    a = S.parse_args("-n 7 -r 9 -t 0.56 -f x.csv -o o.csv -s s.json -v -w 3 --regional-bonus 0.2 --beta-increment-interval-rounds 2.0".split())
    assert (a.num_simulations, a.max_rounds, a.supermajority_threshold, a.workers, a.verbose) == (7, 9, 0.56, 3, True)
    mp = S.model_params_from_args(a)
    assert set(mp) == {"initial_beta_weight", "enable_dynamic_beta", "beta_increment_amount", "beta_increment_interval_rounds",
                       "stickiness_factor", "bandwagon_strength", "regional_affinity_bonus", "papabile_weight_factor",
                       "fatigue_penalty_factor", "stop_candidate_threshold_unacceptable_distance",
                       "stop_candidate_boost_factor", "stop_candidate_threat_min_vote_share"}
    assert mp["regional_affinity_bonus"] == 0.2 and mp["beta_increment_interval_rounds"] == 2 and isinstance(mp["beta_increment_interval_rounds"], int)


def test_aggregation_and_writers_match_reference_contract(tmp_path):
    from conclave_sim_b200 import simulate as S
    results = [dict(sim_id=i, winner=w, rounds=r, reason=S.reason_text(1 if w else 0, 0.56))
               for i, (w, r) in enumerate([("39", 7), (None, 20), ("12", 9), ("39", 11), (None, 20), ("7", 6)])]
    agg = S.aggregate_results(results, 6)
    ref = O.aggregate([r["winner"] for r in results], [r["rounds"] for r in results], 6)
    for k in ref:
        assert (agg[k] == ref[k]) or np.isclose(agg[k], ref[k]), k
    assert list(agg) == list(ref)      # same key order -> same JSON text
    S.write_outputs(results, agg, str(tmp_path / "r.csv"), str(tmp_path / "d" / "s.json"))
    lines = (tmp_path / "r.csv").read_text().splitlines()
    assert lines[0] == "sim_id,winner,rounds,reason"
    assert lines[1] == "0,39,7,Winner found by supermajority (56% threshold)"
    assert lines[2] == "1,,20,Max rounds reached"
    js = json.loads((tmp_path / "d" / "s.json").read_text())
    assert js["winner_counts"] == {"39": 2, "12": 1, "7": 1} and js["most_frequent_winner"] == "39"
    assert js["median_rounds_successful"] == 8.0 and js["min_rounds_successful"] == 6
    empty = S.aggregate_results([dict(sim_id=0, winner=None, rounds=20, reason="Max rounds reached")], 1)
    S.write_outputs([], empty, str(tmp_path / "e.csv"), str(tmp_path / "e.json"))
    assert json.loads((tmp_path / "e.json").read_text())["average_rounds_successful"] is None
    assert S.reason_text(2, 0.5) == "No probabilities generated" and S.reason_text(1, 2 / 3) == "Winner found by supermajority (67% threshold)"


def test_ingest_and_model_mirror_host_behaviour(tmp_path, caplog):
    import logging
    import pandas as pd
    from conclave_sim_b200.ingest import ElectorDataIngester, roster_arrays
    from conclave_sim_b200.model import TransitionModel
    p = tmp_path / "e.csv"
    p.write_text("elector_id,region,is_papabile,ideology_score\n0,Europe,True,0.25\n1,Asia,False,0.5\n2,Europe,False,0.75\n")
    df = ElectorDataIngester(str(p)).load_and_prepare_data()
    assert list(df.index) == ["0", "1", "2"] and df.index.name == "elector_id"
    x, g, pap = roster_arrays(df)
    assert list(x) == [0.25, 0.5, 0.75] and g[0] == g[2] != g[1] and list(pap) == [True, False, False]
    assert ElectorDataIngester(str(tmp_path / "missing.csv")).load_and_prepare_data().empty
    (tmp_path / "g.csv").write_text("gc_id,region,ideology_score\n5,a,0.1\n")
    assert list(ElectorDataIngester(str(tmp_path / "g.csv")).load_and_prepare_data().index) == ["5"]
    # model mirror: attributes, stateful beta, empty-data behaviour and log texts (tests/test_model.py:428-468,1006-1079,1165-1216)
    m = TransitionModel(df, initial_beta_weight=1.0, enable_dynamic_beta=True, beta_growth_rate=1.1)
    for r, want in ((1, 1.1), (2, 1.1 * 1.1)):
        m.update_beta_weight(r)
        assert m.effective_beta_weight == pytest.approx(want, rel=1e-15)
    assert m.get_candidate_ids() == ["0", "1", "2"] and m.get_elector_data() is df
    with caplog.at_level(logging.WARNING):
        e = TransitionModel(pd.DataFrame(), initial_beta_weight=1.0)
        P, d = e.calculate_transition_probabilities(current_round_num=1)
    assert P.shape == (0, 0) and d == []
    assert "TransitionModel initialized with empty elector_data. Most operations will result in empty outputs." in caplog.text
    assert "No electors or actual candidates to calculate transition probabilities for." in caplog.text
    caplog.clear()
    with caplog.at_level(logging.WARNING):
        TransitionModel(df.reset_index(drop=True))
    assert "Elector data index is not named 'elector_id' and 'elector_id' column not found." in caplog.text
    with pytest.raises(TypeError):
        from conclave_sim_b200 import make_params
        make_params({"not_a_model_kwarg": 1})


def test_worker_on_empty_roster_matches_reference():
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    r = S._run_single_simulation_worker(3, pd.DataFrame(), {}, 10, 0.5, 5, 2, False, False, 3, 0.05, 3, False)
    assert r == {"sim_id": 3, "winner": None, "rounds": 1, "reason": "No probabilities generated", "final_votes": {}, "run_details": []}


def test_stats_from_histograms_equals_row_aggregation():
    from conclave_sim_b200 import simulate as S
    rng = np.random.default_rng(5)
    ids = [str(i) for i in range(30)]
    for n in (0, 1, 2, 7, 50, 501):
        w = rng.integers(-1, 30, n)
        w[rng.random(n) < 0.3] = -1
        r = np.where(w >= 0, rng.integers(1, 21, n), 20)
        results = [dict(sim_id=i, winner=(ids[w[i]] if w[i] >= 0 else None), rounds=int(r[i]), reason="") for i in range(n)]
        ref = S.aggregate_results(results, n)
        wh = np.bincount(np.where(w >= 0, w, 30), minlength=31)
        rh = np.bincount(r[w >= 0], minlength=21)
        got = S.stats_from_histograms(wh, rh, np.array([r.sum(), n]), ids, n)
        for k in ref:
            if k in ("most_frequent_winner",):
                assert (got[k] is None) == (ref[k] is None)
                if got[k] is not None:
                    assert ref["winner_counts"][got[k]] == ref["most_frequent_winner_count"]
            elif ref[k] is None:
                assert got[k] is None, k
            elif isinstance(ref[k], dict):
                assert got[k] == ref[k], k
            else:
                assert np.isclose(got[k], ref[k]), (k, got[k], ref[k])
        assert list(got) == list(ref)


def test_legacy_shim_input_validation():
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    with pytest.raises(ValueError, match="Elector data cannot be empty"):
        S.run_monte_carlo_simulation(1, pd.DataFrame(columns=["elector_id", "ideology_score"]), {"beta_weight": 0.1})
    with pytest.raises(ValueError, match="must contain an 'elector_id' column"):
        S.run_monte_carlo_simulation(1, pd.DataFrame({"ideology_score": [0.1, 0.2]}), {"beta_weight": 0.1})


def test_plain_c_client_compiles_and_links(tmp_path):
    """tests/c/c_client.c — a C99 program using nothing but include/conclave_b200.h — compiles with -Wall -Werror -pedantic
    and links against the shared library (it is RUN by the GPU test test_plain_c_client_matches_python)."""
    from conclave_sim_b200._cabi import LIB_PATH
    exe = tmp_path / "c_client"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "c_client.c"), LIB_PATH, "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-o", str(exe)],
                   check=True)
    assert exe.exists()


def test_vote_thresholds_reproduce_the_share_tests(lib):
    """csim_vote_threshold: the reference compares SHARES (votes / E in fp64) with its thresholds (simulate.py:111 `<`,
    model.py:397 `>`); the kernels compare vote COUNTS with these integers — the two must agree for every count, ties included."""
    rng = np.random.default_rng(3)
    for E in (1, 2, 7, 85, 134, 135, 1024):
        shares = list(rng.random(40)) + [k / E for k in range(0, E + 1, max(1, E // 17))] + [0.0, 1.0, -0.1, 1.5, 0.05, 0.15, 0.03,
                                                                                            np.nextafter(3 / E, 0), np.nextafter(3 / E, 1)]
        v = np.arange(0, E + 1)
        sh = v.astype(np.float64) / np.float64(E)
        for s in shares:
            lo = lib.csim_vote_threshold(float(s), E, 0)
            hi = lib.csim_vote_threshold(float(s), E, 1)
            assert np.array_equal(sh < s, v < lo), (E, s)
            assert np.array_equal(sh > s, v >= hi), (E, s)


def test_results_frame_and_fast_csv_equal_the_row_by_row_path(tmp_path):
    """The CLI builds its results table column-wise (results_frame) and writes large tables with pyarrow: both must give
    exactly what the reference's row-by-row route gives (results_from_batch -> DataFrame(rows).to_csv, simulate.py:395-412)."""
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    from conclave_sim_b200.engine import RunResult
    rng = np.random.default_rng(0)
    E = 134
    ids = [f"E{i:03d}" for i in range(E)]
    for n, with_reason in ((1000, True), (250_000, True), (3000, False)):
        w = rng.integers(-1, E, n).astype(np.int32)
        r = rng.integers(1, 21, n).astype(np.int32)
        reason = np.where(w >= 0, 1, rng.choice([0, 2], n)).astype(np.uint8) if with_reason else None
        res = RunResult(w, r, reason, None, None, None, None, None, E)
        df = S.results_frame(res, ids, 0.56, sim_id_begin=7)
        rows = pd.DataFrame(S.results_from_batch(res, ids, 0.56, sim_id_begin=7))
        assert list(df.columns) == ["sim_id", "winner", "rounds", "reason"] and len(df) == n
        assert df["winner"].isna().sum() == int((w < 0).sum())
        a, b = tmp_path / "fast.csv", tmp_path / "rows.csv"
        S._write_csv(df, str(a))                      # pyarrow for n >= 200 000, pandas below
        rows.to_csv(b, index=False)
        assert a.read_bytes() == b.read_bytes()
    # an elector id that needs CSV quoting: the fast writer must hand over to pandas, not mangle it
    ids2 = list(ids); ids2[3] = 'Doe, John "JD"'
    res = RunResult(np.full(200_001, 3, np.int32), np.ones(200_001, np.int32), np.ones(200_001, np.uint8), None, None, None, None, None, E)
    df = S.results_frame(res, ids2, 0.56)
    S._write_csv(df, str(tmp_path / "q.csv"))
    assert pd.read_csv(tmp_path / "q.csv")["winner"].iloc[0] == ids2[3]


# ---------------------------------------------------------------------------------------------
# Round 2: the writers against the bytes the UNMODIFIED reference's main() wrote (tests/golden/main, oracle/gen_golden_main.py),
# the multi-GPU host logic, the new CLI flags
# ---------------------------------------------------------------------------------------------
MAIN_GOLD = os.path.join(ROOT, "tests", "golden", "main")


def _main_cases():
    with open(os.path.join(MAIN_GOLD, "main_cases.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _main_cases(), ids=lambda c: c["name"])
def test_writers_reproduce_the_reference_main_bytes(tmp_path, case):
    """src/simulate.py:356-431 — the rows the reference's main() aggregated are fed to our aggregation / writers in all three
    forms (row dicts -> aggregate_results; batch arrays -> results_frame + stats_from_histograms); CSV and JSON must be the same
    BYTES the reference wrote."""
    import pandas as pd
    import golden_util as G
    from conclave_sim_b200 import simulate as S
    from conclave_sim_b200.engine import RunResult
    want_csv = open(os.path.join(MAIN_GOLD, case["name"] + ".csv"), "rb").read()
    want_json = open(os.path.join(MAIN_GOLD, case["name"] + ".json"), "rb").read()
    ids = G.real_roster().ids
    n = case["n"]
    args = S.parse_args(["-n", str(n)] + case["args"])
    thr = args.supermajority_threshold
    if n:
        ref = pd.read_csv(os.path.join(MAIN_GOLD, case["name"] + ".csv"), dtype={"winner": str}, keep_default_na=True)
        winners = [None if pd.isna(w) else w for w in ref["winner"]]
        results = [dict(sim_id=int(s), winner=w, rounds=int(r), reason=str(t)) for s, w, r, t in zip(ref["sim_id"], winners, ref["rounds"], ref["reason"])]
    else:
        winners, results = [], []
    # 1. the reference's own row-by-row form
    agg = S.aggregate_results(results, n)
    S.write_outputs(results, agg, str(tmp_path / "a.csv"), str(tmp_path / "a.json"))
    assert open(tmp_path / "a.csv", "rb").read() == want_csv
    assert open(tmp_path / "a.json", "rb").read() == want_json
    # 2. what main() does with a GPU batch: arrays -> column-wise frame, histograms -> stats
    E, R = len(ids), args.max_rounds
    widx = np.array([ids.index(w) if w is not None else -1 for w in winners], np.int32)
    rounds = np.array([r["rounds"] for r in results], np.int32)
    reason = np.array([1 if r["reason"].startswith("Winner") else 0 for r in results], np.uint8)
    wh = np.bincount(np.where(widx >= 0, widx, E), minlength=E + 1).astype(np.int64)
    rh = np.bincount(rounds[widx >= 0], minlength=R + 1).astype(np.int64)
    tot = np.array([rounds.sum(), n], np.int64)
    batch = RunResult(widx, rounds, reason, None, None, wh[None], rh[None], tot[None], E)
    frame = S.results_frame(batch, ids, thr)
    stats = S.stats_from_histograms(wh, rh, tot, ids, n, winners_in_order=widx)
    S.write_outputs(frame, stats, str(tmp_path / "b.csv"), str(tmp_path / "b.json"))
    assert open(tmp_path / "b.csv", "rb").read() == want_csv
    assert open(tmp_path / "b.json", "rb").read() == want_json


def test_shard_range_c_abi_equals_python(lib):
    import ctypes as C
    from conclave_sim_b200.dist import shard_range
    for n in (0, 1, 2, 7, 8, 9, 1000, 10_000_000, 2**40 + 3):
        for world in (1, 2, 3, 8, 16):
            covered = 0
            for rank in range(world):
                b, c = C.c_uint64(), C.c_uint64()
                lib.csim_shard_range(n, rank, world, C.byref(b), C.byref(c))
                assert (b.value, c.value) == shard_range(n, rank, world)
                assert b.value == covered
                covered += c.value
            assert covered == n


def test_resolve_devices_and_new_cli_flags(monkeypatch):
    from conclave_sim_b200 import simulate as S
    monkeypatch.setattr(S, "device_count", lambda: 8)
    assert S.resolve_devices(None, 0) == (0,)
    assert S.resolve_devices(None, 3) == (3,)
    assert S.resolve_devices(None, 0, workers=1) == (0,)
    assert S.resolve_devices(None, 0, workers=4) == (0, 1, 2, 3)
    assert S.resolve_devices(None, 0, workers=64) == tuple(range(8))          # -w clamps to the box
    assert S.resolve_devices(2) == (0, 1) and S.resolve_devices("2") == (0, 1)
    assert S.resolve_devices("all") == tuple(range(8))
    assert S.resolve_devices("3,1") == (3, 1) and S.resolve_devices([5, 6]) == (5, 6)
    for bad in (0, "0", 9, [1, 1], []):
        with pytest.raises(ValueError):
            S.resolve_devices(bad)
    monkeypatch.setattr(S, "device_count", lambda: 1)
    assert S.resolve_devices(None, 0, workers=16) == (0,)
    a = S.parse_args("--devices 4 --tape t.npy -w 2".split())
    assert (a.devices, a.tape, a.workers) == ("4", "t.npy", 2)
    a = S.parse_args([])
    assert a.devices is None and a.tape is None


def test_load_tapes(tmp_path):
    from conclave_sim_b200 import simulate as S
    t = np.random.default_rng(0).random((5, 3 * 7))
    np.save(tmp_path / "t.npy", t)
    np.savez(tmp_path / "t.npz", tapes=t)
    for f in ("t.npy", "t.npz"):
        got = S.load_tapes(str(tmp_path / f), 4, 3, 7)
        assert got.shape == (4, 21) and np.array_equal(got, t[:4]) and got.flags.c_contiguous
    with pytest.raises(ValueError):
        S.load_tapes(str(tmp_path / "t.npy"), 6, 3, 7)           # not enough rows
    with pytest.raises(ValueError):
        S.load_tapes(str(tmp_path / "t.npy"), 4, 4, 7)           # wrong row length


def test_multi_gpu_entry_points_without_a_device(lib):
    """No GPU here: the group / NCCL entry points must fail with a status and a message, never crash or compute."""
    import ctypes as C
    from conclave_sim_b200 import _cabi
    g = C.c_void_p()
    devs = (C.c_int32 * 2)(0, 1)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    assert lib.csim_group_create(C.byref(g), devs, 2) != 0 and not g.value
    assert b"" != lib.csim_last_error()
    assert lib.csim_group_create(C.byref(g), None, 0) == -1
    n = C.c_int32(7)
    assert lib.csim_device_count(C.byref(n)) != 0 and n.value == 0
    assert lib.csim_group_size(None) == 0
    assert lib.csim_allreduce_hists(None, None, None, None, None, 1, 1, None) == -1
    # NCCL binds at run time: loading it needs no GPU (skip when the box has no NCCL at all)
    try:
        v = _cabi.nccl_load()
    except _cabi.CsimError:
        pytest.skip("no libnccl on this box")
    assert v >= 21800 and lib.csim_nccl_version() == v


def test_stats_first_appearance_order_and_ties():
    from conclave_sim_b200 import simulate as S
    ids = ["a", "b", "c", "d"]
    w = np.array([2, -1, 0, 2, 0, 3])
    wh = np.bincount(np.where(w >= 0, w, 4), minlength=5)
    rh = np.bincount(np.array([3, 4, 3, 5, 6]), minlength=8)
    got = S.stats_from_histograms(wh, rh, np.array([30, 6]), ids, 6, winners_in_order=w)
    assert list(got["winner_counts"].items()) == [("c", 2), ("a", 2), ("d", 1)]
    assert got["most_frequent_winner"] == "c"                     # the reference's Counter.most_common: first inserted of the equal counts
    got = S.stats_from_histograms(wh, rh, np.array([30, 6]), ids, 6)
    assert list(got["winner_counts"].items()) == [("a", 2), ("c", 2), ("d", 1)] and got["most_frequent_winner"] == "a"

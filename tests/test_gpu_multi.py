"""The multi-GPU split through the drop-in surface: csim_group_run / EngineGroup, run_simulations(devices=),
`python -m src.simulate --devices N`, the reference's own `-w` knob, csim_allreduce_hists, and what AUTO reports.

Everything that needs two devices skips on a one-GPU box (the driver's round-end box has one); those tests are run with
`gpurun --gpus 2` during the round and their log is committed under profiles/.  The one-device forms of the same code paths
(a group of one device through the general path, an NCCL communicator of one rank) run everywhere.
"""
import ctypes as C
import json
import os
import subprocess
import sys
import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import conclave_oracle as O
from oracle import c_oracle as CO
from test_gpu_parity import to_params, upload, CFG2, EP2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DENSE, TABLE, PREFIX, FULL = 1, 2, 3, 4
REF_CLI, MODEL = 0, 1


def n_gpus():
    from conclave_sim_b200.engine import device_count
    return device_count()


need2 = pytest.mark.skipif("n_gpus() < 2", reason="needs two CUDA devices")


@pytest.fixture(scope="module")
def eng():
    from conclave_sim_b200 import Engine
    CO.build()
    e = Engine(0)
    yield e
    e.close()


def same(a, b, tall=False):
    assert np.array_equal(a.winner, b.winner) and np.array_equal(a.rounds, b.rounds) and np.array_equal(a.reason, b.reason)
    assert np.array_equal(a.winner_hist, b.winner_hist) and np.array_equal(a.rounds_hist, b.rounds_hist) and np.array_equal(a.totals, b.totals)
    if tall:
        assert np.array_equal(a.round_tallies, b.round_tallies) and np.array_equal(a.final_votes, b.final_votes)


def group_vs_single(eng, devices, n=3001):
    """Every engine, two parameter sets, a trial count the device count does not divide: group == one device, bit for bit."""
    from conclave_sim_b200 import FP32, FP64
    from conclave_sim_b200.engine import EngineGroup
    roster = O.synthetic_roster(135)
    upload(eng, roster)
    grp = EngineGroup(devices)
    grp.upload_roster(roster.ideology, roster.region, roster.papabile)
    p1 = to_params(O.ModelParams(**CFG2), EP2)
    p2 = to_params(O.ModelParams(**{**CFG2, "initial_beta_weight": 0.7}), O.EngineParams(max_rounds=20, supermajority_threshold=0.6))
    try:
        for wiring, engine, prec in ((REF_CLI, DENSE, FP32), (REF_CLI, TABLE, FP32), (REF_CLI, TABLE, FP64), (REF_CLI, PREFIX, FP32),
                                     (MODEL, PREFIX, FP32), (MODEL, FULL, FP32), (MODEL, DENSE, FP64), (REF_CLI, 0, FP32), (MODEL, 0, FP32)):
            nn = n if prec == FP32 else 257
            a = eng.run([p1, p2], nn, seed=11, sim_id_begin=5, wiring=wiring, engine=engine, precision=prec)
            b = grp.run([p1, p2], nn, seed=11, sim_id_begin=5, wiring=wiring, engine=engine, precision=prec)
            same(a, b)
            assert b.devices == tuple(devices) and len(b.kernel_ms_per_device) == len(devices) and b.engine == a.engine
        # per-trial tallies / final votes and a uniform tape through the strided row copies
        tapes = np.random.default_rng(3).random((2 * 37, 20 * 135))
        a = eng.run([p1, p2], 37, wiring=REF_CLI, engine=DENSE, precision=FP64, tapes=tapes, want_round_tallies=True, want_final_votes=True)
        b = grp.run([p1, p2], 37, wiring=REF_CLI, engine=DENSE, precision=FP64, tapes=tapes, want_round_tallies=True, want_final_votes=True)
        same(a, b, tall=True)
        # fewer trials than devices: the idle device still joins the all-reduce
        a = eng.run([p1], 1, seed=3, wiring=REF_CLI, engine=PREFIX, precision=FP32)
        b = grp.run([p1], 1, seed=3, wiring=REF_CLI, engine=PREFIX, precision=FP32)
        same(a, b)
        # empty batch
        z = grp.run([p1], 0)
        assert z.winner.shape == (0,) and z.totals.sum() == 0
        # AUTO must not depend on the split: 60 001 ref_cli trials are a TABLE job on one device, and each device of the group — whose
        # own shard alone would be a PREFIX job — must run TABLE too (csim_run_args.auto_batch_sims), or the rows would differ
        a = eng.run([p1], 60001, seed=21, wiring=REF_CLI)
        b = grp.run([p1], 60001, seed=21, wiring=REF_CLI)
        assert a.engine == "table" and b.engine == "table"
        same(a, b)
    finally:
        grp.close()


def test_group_of_one_device_equals_engine(eng):
    group_vs_single(eng, [0])


@need2
def test_group_of_two_devices_equals_one(eng):
    group_vs_single(eng, [0, 1])


@need2
def test_group_of_all_devices_equals_one(eng):
    group_vs_single(eng, list(range(n_gpus())), n=10007)


def _frame(E=135):
    ro = O.synthetic_roster(E)
    return pd.DataFrame({"ideology_score": ro.ideology, "region": ro.region, "is_papabile": ro.papabile},
                        index=pd.Index([str(i) for i in range(E)], name="elector_id"))


def test_run_simulations_devices_argument(eng):
    from conclave_sim_b200 import simulate as S
    df = _frame()
    kw = dict(max_rounds=20, supermajority_threshold=0.56, runoff_threshold_rounds=5, runoff_candidate_count=2, seed=99)
    a = S.run_simulations(df, CFG2, 2000, **kw)
    b = S.run_simulations(df, CFG2, 2000, devices=1, **kw)
    c = S.run_simulations(df, CFG2, 2000, devices=[0], **kw)
    d = S.run_simulations(df, CFG2, 2000, workers=64, **kw)               # -w: clamped to the devices the box has
    for o in (b, c, d):
        same(a, o)
    assert a.engine in ("table", "prefix") and a.precision == "fp32" and not a.fell_through
    if n_gpus() >= 2:
        e = S.run_simulations(df, CFG2, 2000, devices=2, **kw)
        same(a, e)
        assert e.devices == (0, 1)
        assert d.devices == tuple(range(min(64, n_gpus())))
    with pytest.raises(ValueError):
        S.run_simulations(df, CFG2, 10, devices=n_gpus() + 1, **kw)


def _cli(args, env_extra=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, "-m", "src.simulate"] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def _roster_csv(tmp_path, E=135):
    df = _frame(E).reset_index()
    path = tmp_path / "roster.csv"
    df.to_csv(path, index=False)
    return str(path)


CLI_CFG2 = ["--max-rounds", "20", "--initial-beta-weight", "1.5", "--enable-dynamic-beta", "--beta-increment-amount", "0.3",
            "--beta-increment-interval-rounds", "1", "--bandwagon-strength", "0.7", "--papabile-weight-factor", "2.0",
            "--stickiness-factor", "0.8", "-t", "0.56", "--seed", "20250507"]


def _run_cli(tmp_path, tag, extra, n=20000):
    csv, js = tmp_path / f"{tag}.csv", tmp_path / f"{tag}.json"
    _cli(["-n", str(n), "-f", _roster_csv(tmp_path), "-o", str(csv), "-s", str(js)] + CLI_CFG2 + extra)
    return csv.read_bytes(), js.read_bytes()


def test_cli_devices_1_and_workers_write_the_same_files(tmp_path):
    base = _run_cli(tmp_path, "base", [])
    assert _run_cli(tmp_path, "d1", ["--devices", "1"]) == base
    assert _run_cli(tmp_path, "w1", ["-w", "1"]) == base
    assert _run_cli(tmp_path, "wmany", ["-w", "64"]) == base            # clamped to the visible devices; same bytes whatever their number
    assert base[0].count(b"\n") == 20001


@need2
def test_cli_devices_2_writes_the_same_files_as_1(tmp_path):
    """Done-criterion of VERDICT r1 item 1: `python -m src.simulate -n N --devices K` writes the same CSV / JSON as --devices 1."""
    base = _run_cli(tmp_path, "base", ["--devices", "1"], n=200000)
    for k in sorted({2, n_gpus()}):
        assert _run_cli(tmp_path, f"d{k}", ["--devices", str(k)], n=200000) == base
    assert _run_cli(tmp_path, "dall", ["--devices", "all"], n=200000) == base
    assert _run_cli(tmp_path, "dlist", ["--devices", "1,0"], n=200000) == base
    m1 = _run_cli(tmp_path, "m1", ["--devices", "1", "--wiring", "model"], n=50000)
    assert _run_cli(tmp_path, "m2", ["--devices", "2", "--wiring", "model"], n=50000) == m1


def test_cli_tape_replays_reference_draws(tmp_path):
    """--tape: the reference's own MT19937 uniforms -> the golden winners / rounds of the unmodified reference (cfg2 case)."""
    import golden_util as G
    case = next(c for c in G.engine_cases("ref_cli") if c["name"] == "cfg2")
    tapes = G.case_tapes(case)
    np.save(tmp_path / "tapes.npy", tapes)
    ro = case["roster"]
    df = pd.DataFrame({"elector_id": ro.ids, "ideology_score": ro.ideology, "region": ro.region, "is_papabile": ro.papabile})
    df.to_csv(tmp_path / "real.csv", index=False)
    csv, js = tmp_path / "t.csv", tmp_path / "t.json"
    _cli(["-n", str(case["n"]), "-f", str(tmp_path / "real.csv"), "-o", str(csv), "-s", str(js), "--tape", str(tmp_path / "tapes.npy")] + CLI_CFG2)
    got = pd.read_csv(csv, dtype={"winner": str})
    want_w = [None if w < 0 else str(ro.ids[w]) for w in case["winners"]]
    assert [None if pd.isna(w) else w for w in got["winner"]] == want_w
    assert got["rounds"].tolist() == case["rounds"].tolist()


def test_allreduce_hists_through_the_c_abi_single_rank(eng):
    """csim_nccl_unique_id / comm_create / csim_allreduce_hists on a communicator of ONE rank: the NCCL binding (dlopen) and the
    call path are exercised on a one-GPU box; the sum over one rank is the identity."""
    import torch
    from conclave_sim_b200 import FP32
    from conclave_sim_b200.dist import DeviceBatch
    from conclave_sim_b200.engine import nccl_unique_id, nccl_comm_destroy
    from conclave_sim_b200 import _cabi
    assert _cabi.nccl_load() >= 21800
    upload(eng, O.synthetic_roster(135))
    comm = eng.nccl_comm_create(nccl_unique_id(), 1, 0)
    try:
        p = to_params(O.ModelParams(**CFG2), EP2)
        ref = eng.run([p], 4096, seed=5, wiring=REF_CLI, engine=PREFIX, precision=FP32)
        b = DeviceBatch(eng, 1, 4096, 20, want_reason=True, comm=comm)
        b.run([p], sim_id_begin=0, seed=5, wiring=REF_CLI, engine=PREFIX, precision=FP32)
        torch.cuda.synchronize()
        wh, rh, tot = b.hists()
        assert np.array_equal(wh, ref.winner_hist) and np.array_equal(rh, ref.rounds_hist) and np.array_equal(tot, ref.totals)
        assert np.array_equal(b.winner.cpu().numpy(), ref.winner)
        # non-contiguous histograms take the grouped form
        w2 = torch.zeros(136, dtype=torch.int64, device="cuda"); r2 = torch.zeros(21, dtype=torch.int64, device="cuda")
        t2 = torch.zeros(2, dtype=torch.int64, device="cuda")
        eng.run_device([p], 4096, winner=b.winner, rounds=b.rounds, winner_hist=w2, rounds_hist=r2, totals=t2, seed=5, engine=PREFIX)
        eng.allreduce_hists(comm, w2, r2, t2, 1, 20)
        torch.cuda.synchronize()
        assert np.array_equal(w2.cpu().numpy(), ref.winner_hist[0]) and np.array_equal(t2.cpu().numpy(), ref.totals[0])
    finally:
        nccl_comm_destroy(comm)


def test_empty_batch_and_null_pointers(eng):
    """csim_run returns CSIM_OK for an empty batch before looking at any pointer (a rank with an empty shard)."""
    from conclave_sim_b200._cabi import CsimRunArgs, CsimParams, load_library
    lib = load_library()
    upload(eng, O.synthetic_roster(20))
    a = CsimRunArgs()
    p = to_params(O.ModelParams(**CFG2), EP2)
    sets = (CsimParams * 1)(p)
    a.params = sets; a.n_param_sets = 1; a.n_sims = 0; a.buffers_on_device = 1
    assert lib.csim_run(eng._ctx, C.byref(a)) == 0
    a.n_param_sets = 0; a.n_sims = 5; a.params = None
    assert lib.csim_run(eng._ctx, C.byref(a)) == 0


def test_auto_reports_engine_and_warns_on_precision_change(eng):
    from conclave_sim_b200 import FP32, FP64
    upload(eng, O.synthetic_roster(135))
    p = to_params(O.ModelParams(**CFG2), EP2)
    assert eng.run([p], 100000, wiring=REF_CLI).engine == "table"
    assert eng.run([p], 100, wiring=REF_CLI).engine == "prefix"
    assert eng.run([p], 100, wiring=MODEL).engine == "prefix"
    # a shard of a larger job: AUTO chooses for the whole job (auto_batch_sims), here TABLE although 1000 trials alone mean PREFIX
    import torch
    from conclave_sim_b200.dist import DeviceBatch
    b = DeviceBatch(eng, 1, 1000, 20, want_reason=True)
    b.run([p], sim_id_begin=0, seed=1, wiring=REF_CLI, engine=0, precision=FP32, reduce=False, job_sims=100000)
    torch.cuda.synchronize()
    assert eng.last_run_info()[0] == "table"
    b.run([p], sim_id_begin=0, seed=1, wiring=REF_CLI, engine=0, precision=FP32, reduce=False)
    torch.cuda.synchronize()
    assert eng.last_run_info()[0] == "prefix"
    r = eng.run([p], 64, wiring=MODEL, precision=FP64)
    assert (r.engine, r.precision) == ("dense", "fp64")
    px = to_params(O.ModelParams(**{**CFG2, "enable_candidate_fatigue": True}), O.EngineParams(max_rounds=20, supermajority_threshold=0.56, enable_candidate_fatigue=True))
    r = eng.run([px], 100, wiring=MODEL)
    assert (r.engine, r.precision, r.fell_through) == ("full", "fp32", False)
    # > 64 regions and a roster whose FULL tables do not fit: AUTO lands on the fp64 exact engine and SAYS so
    rng = np.random.default_rng(1)
    E = 1024
    eng.upload_roster(rng.beta(2, 4, E), rng.integers(0, 80, E).astype(np.int32), rng.random(E) < 0.15)
    pn = to_params(O.ModelParams(**{**CFG2, "regional_affinity_bonus": -0.05}), O.EngineParams(max_rounds=3, supermajority_threshold=0.56))
    with pytest.warns(RuntimeWarning, match="fp64 exact engine"):
        r = eng.run([pn], 4, wiring=MODEL)
    assert (r.engine, r.precision, r.fell_through) == ("dense", "fp64", True)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        eng.run([pn], 4, wiring=MODEL, precision=FP64)                  # asked for fp64: nothing to warn about


def test_sweep_beyond_65535_set_rounds(eng):
    """n_param_sets * max_rounds > 65535 (gridDim.y/z of the table-build kernels) — launches are tiled now."""
    from conclave_sim_b200 import FP32, FP64
    ro = O.synthetic_roster(12)
    upload(eng, ro)
    n_sets, R = 700, 100
    sets, osets = [], []
    for i in range(n_sets):
        mp = O.ModelParams(**{**CFG2, "initial_beta_weight": 0.5 + 0.002 * i})
        ep = O.EngineParams(max_rounds=R, supermajority_threshold=0.9)
        sets.append(to_params(mp, ep)); osets.append((mp, ep))
    ref = CO.run(ro, osets, O.WIRING_REF_CLI, 3, seed=4)
    for engine, prec in ((0, FP32), (TABLE, FP64), (DENSE, FP64), (PREFIX, FP32), (DENSE, FP32), (TABLE, FP32)):
        got = eng.run(sets, 3, seed=4, wiring=REF_CLI, engine=engine, precision=prec)
        agree = np.mean((got.winner == ref["winner"]) & (got.rounds == ref["rounds"]))
        assert agree == 1.0 if prec == FP64 else agree > 0.995, (engine, prec, agree)


def test_full_engine_negative_threat_share_round_one(eng):
    """with a negative stop_candidate_threat_min_vote_share every candidate is a threat already in round 1
    (0.0 > threshold, simulate.py:73 / model.py:397): FULL must agree with the fp64 engine and the oracle."""
    from conclave_sim_b200 import FP32, FP64
    ro = O.synthetic_roster(60)
    upload(eng, ro)
    mp = O.ModelParams(**{**CFG2, "enable_stop_candidate": True, "stop_candidate_threat_min_vote_share": -0.1,
                          "stop_candidate_threshold_unacceptable_distance": 0.25, "stop_candidate_boost_factor": 2.5})
    ep = O.EngineParams(max_rounds=6, supermajority_threshold=0.5, enable_stop_candidate=True)
    p = to_params(mp, ep)
    ref = CO.run(ro, [(mp, ep)], O.WIRING_MODEL, 3000, seed=8, want_tallies=True)
    ex = eng.run([p], 3000, seed=8, wiring=MODEL, engine=DENSE, precision=FP64, want_round_tallies=True)
    assert np.array_equal(ex.winner, ref["winner"]) and np.array_equal(ex.round_tallies, ref["round_tallies"])
    got = eng.run([p], 3000, seed=8, wiring=MODEL, engine=FULL, precision=FP32, want_round_tallies=True)
    # round 1 is where the bug was: compare its tallies trial by trial
    same_r1 = np.mean([(got.round_tallies[i, 0] == ref["round_tallies"][i, 0]).all() for i in range(3000)])
    assert same_r1 > 0.97, same_r1
    assert np.mean((got.winner == ref["winner"]) & (got.rounds == ref["rounds"])) > 0.99
    # and the probe: round-1 probabilities with the boost live
    from test_gpu_probs import check_against
    P = eng.prob_matrix(p, eng.beta_schedule(p, 1, 1)[0], shares=np.zeros(60), precision=FP64)
    check_against(eng, p, MODEL, FULL, 1, P, what="negative threat share, round 1")


def _main_cases():
    with open(os.path.join(ROOT, "tests", "golden", "main", "main_cases.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _main_cases(), ids=lambda c: c["name"])
def test_cli_end_to_end_equals_the_reference_main(tmp_path, case):
    """`python -m src.simulate` against the UNMODIFIED reference's main() (tests/golden/main, oracle/gen_golden_main.py): same
    command line + the reference's own MT19937 draws (--tape) -> the same CSV and stats JSON, byte for byte — engine,
    aggregation (src/simulate.py:356-387) and writers (:395-426) in one comparison."""
    import golden_util as G
    gold = os.path.join(ROOT, "tests", "golden", "main")
    ro = G.real_roster()
    pd.DataFrame({"elector_id": ro.ids, "ideology_score": ro.ideology, "region": ro.region, "is_papabile": ro.papabile}).to_csv(
        tmp_path / "real.csv", index=False)
    args = ["-n", str(case["n"]), "-f", str(tmp_path / "real.csv"), "-o", str(tmp_path / "o.csv"), "-s", str(tmp_path / "o.json")] + case["args"]
    if case["n"]:
        R = int(case["args"][case["args"].index("--max-rounds") + 1])
        np.save(tmp_path / "tapes.npy", np.stack([G.mt_tape(case["seed0"] + i, R * ro.E) for i in range(case["n"])]))
        args += ["--tape", str(tmp_path / "tapes.npy")]
    _cli(args)
    assert open(tmp_path / "o.csv", "rb").read() == open(os.path.join(gold, case["name"] + ".csv"), "rb").read()
    assert open(tmp_path / "o.json", "rb").read() == open(os.path.join(gold, case["name"] + ".json"), "rb").read()


# ---------------------------------------------------------------------------------------------------------------------
# a C host over the group entry points (no Python, no torch in the process that drives the GPUs)
# ---------------------------------------------------------------------------------------------------------------------
def _c_client(tmp_path):
    from conclave_sim_b200._cabi import LIB_PATH
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "c_client"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", "c_client.c"), LIB_PATH, "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-o", str(exe)],
                   check=True)
    return str(exe)


def _c_client_group_equals_single(tmp_path, n_devices, n=3001):
    """tests/c/c_client.c with a device count drives csim_group_create / upload_roster / run / destroy from plain C; its whole
    output (per-trial rows, totals, both histograms, the status of an unsupported request) is the single-context run's."""
    from conclave_sim_b200._cabi import nccl_library_path
    exe = _c_client(tmp_path)
    env = {k: v for k, v in os.environ.items() if k != "NCCL_DEBUG"}       # (NCCL's version banner goes to stdout otherwise)
    lib = nccl_library_path()
    if lib:
        env["CSIM_NCCL_LIB"] = lib                    # the NCCL this image's torch ships; the system copy works as well
    for engine, wiring, precision in ((0, 0, 0), (1, 1, 0), (2, 0, 0), (2, 0, 1), (3, 1, 0), (4, 1, 0)):
        base = [exe, "135", str(n), str(engine), str(wiring), str(precision), "20250507"]
        one = subprocess.run(base, capture_output=True, text=True, env=env, timeout=600)
        grp = subprocess.run(base + [str(n_devices)], capture_output=True, text=True, env=env, timeout=600)
        assert one.returncode == 0, one.stderr
        assert grp.returncode == 0, grp.stderr
        a = one.stdout.strip().split("\n")
        b = [ln for ln in grp.stdout.strip().split("\n") if not ln.startswith("NCCL version")]
        assert len(a) == n + 4 and a == b, (engine, wiring, precision, [i for i, (u, v) in enumerate(zip(a, b)) if u != v][:5])


def test_c_host_group_of_one_device(tmp_path):
    _c_client_group_equals_single(tmp_path, 1)


@need2
def test_c_host_group_of_two_devices(tmp_path):
    _c_client_group_equals_single(tmp_path, 2)


@need2
def test_c_host_group_of_all_devices(tmp_path):
    _c_client_group_equals_single(tmp_path, n_gpus(), n=10007)


def test_c_host_group_asks_for_too_many_devices(tmp_path):
    exe = _c_client(tmp_path)
    r = subprocess.run([exe, "135", "100", "0", "0", "0", "1", str(n_gpus() + 1)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 3 and "visible" in r.stderr

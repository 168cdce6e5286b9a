"""ctypes wrapper + build recipe for oracle/conclave_oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__ (build, smoke) and bench.py's cpu_baseline / --impl reference
legs may import this.  The struct mirror below is deliberately independent of the product's
own binding (conclave_sim_b200/_cabi.py); tests/test_host_cpu.py::test_library_exports_every_declared_symbol checks both against the header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libconclave_oracle.so")
SRC = os.path.join(HERE, "conclave_oracle.c")
INC = os.path.join(os.path.dirname(HERE), "include")


class CParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "initial_beta_weight", "beta_increment_amount", "beta_growth_rate", "stickiness_factor",
        "bandwagon_strength", "regional_affinity_bonus", "papabile_weight_factor",
        "fatigue_vote_share_threshold", "fatigue_penalty_factor",
        "stop_candidate_threshold_unacceptable_distance", "stop_candidate_threat_min_vote_share",
        "stop_candidate_boost_factor", "supermajority_threshold")] + [(n, C.c_int32) for n in (
        "enable_dynamic_beta", "beta_mode", "beta_increment_interval_rounds", "enable_candidate_fatigue",
        "fatigue_threshold_rounds", "fatigue_top_n_immune", "enable_stop_candidate", "max_rounds",
        "runoff_threshold_rounds", "runoff_candidate_count", "reserved0", "reserved1")]


def build(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(INC, "conclave_b200.h"))):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-I", INC, SRC, "-o", LIB, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.co_need_votes.restype = C.c_int
        _lib.co_need_votes.argtypes = [C.c_double, C.c_int]
        _lib.co_pairwise_sum.restype = C.c_double
        _lib.co_pairwise_sum.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.co_run.restype = C.c_int
    return _lib


def set_faithful(on: bool) -> None:
    """True: co_run re-evaluates the whole E x C matrix per trial per round like the reference does
    (no trial-invariant tables).  Same results, reference-like cost; used for the CPU baseline."""
    lib().co_set_faithful(C.c_int(1 if on else 0))


def to_cparams(mp, ep) -> CParams:
    """mp/ep: oracle.conclave_oracle.ModelParams / EngineParams."""
    p = CParams()
    p.initial_beta_weight = float(mp.initial_beta_weight) if mp.initial_beta_weight is not None else 0.0
    p.beta_increment_amount = 0.0 if mp.beta_increment_amount is None else float(mp.beta_increment_amount)
    p.beta_mode = 1 if mp.beta_increment_amount is None else 0
    p.beta_growth_rate = float(mp.beta_growth_rate)
    p.beta_increment_interval_rounds = int(mp.beta_increment_interval_rounds)
    p.enable_dynamic_beta = int(bool(mp.enable_dynamic_beta))
    p.stickiness_factor = float(mp.stickiness_factor)
    p.bandwagon_strength = float(mp.bandwagon_strength)
    p.regional_affinity_bonus = float(mp.regional_affinity_bonus)
    p.papabile_weight_factor = float(mp.papabile_weight_factor)
    p.fatigue_penalty_factor = float(mp.fatigue_penalty_factor)
    p.enable_candidate_fatigue = int(bool(mp.enable_candidate_fatigue))
    p.enable_stop_candidate = int(bool(mp.enable_stop_candidate))
    p.stop_candidate_threshold_unacceptable_distance = float(mp.stop_candidate_threshold_unacceptable_distance)
    p.stop_candidate_threat_min_vote_share = float(mp.stop_candidate_threat_min_vote_share)
    p.stop_candidate_boost_factor = float(mp.stop_candidate_boost_factor)
    if ep is not None:
        p.fatigue_threshold_rounds = int(ep.fatigue_threshold_rounds)
        p.fatigue_vote_share_threshold = float(ep.fatigue_vote_share_threshold)
        p.fatigue_top_n_immune = int(ep.fatigue_top_n_immune)
        p.supermajority_threshold = float(ep.supermajority_threshold)
        p.max_rounds = int(ep.max_rounds)
        p.runoff_threshold_rounds = int(ep.runoff_threshold_rounds)
        p.runoff_candidate_count = int(ep.runoff_candidate_count)
    else:
        p.fatigue_threshold_rounds = int(mp.fatigue_threshold_rounds)
        p.fatigue_vote_share_threshold = float(mp.fatigue_vote_share_threshold)
        p.fatigue_top_n_immune = int(mp.fatigue_top_n_immune)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def beta_schedule(mp, first_round: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.float64)
    p = to_cparams(mp, None)
    lib().co_beta_schedule(C.byref(p), C.c_int(first_round), C.c_int(n), _ptr(out))
    return out


def need_votes(thr: float, E: int) -> int:
    return int(lib().co_need_votes(thr, E))


def pairwise_sum(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float64)
    return float(lib().co_pairwise_sum(_ptr(a), len(a), 1))


def prob_matrix(roster, mp, beta, prev_vote=None, prev_count=None, fatigued=None, shares=None) -> np.ndarray:
    E = roster.E
    x = np.ascontiguousarray(roster.ideology, np.float64)
    g = np.ascontiguousarray(roster.region, np.int32)
    pap = np.ascontiguousarray(roster.papabile, np.uint8)
    p = to_cparams(mp, None)
    P = np.zeros((E, E), np.float64)
    pv = None if prev_vote is None else np.ascontiguousarray(prev_vote, np.int32)
    pc = None if prev_count is None else np.ascontiguousarray(prev_count, np.int32)
    if pc is None and pv is not None:
        pc = np.bincount(pv[pv >= 0], minlength=E).astype(np.int32)
    ft = None if fatigued is None else np.ascontiguousarray(fatigued, np.uint8)
    sh = None if shares is None else np.ascontiguousarray(shares, np.float64)
    lib().co_prob_matrix(C.c_int(E), _ptr(x), _ptr(g), _ptr(pap), C.byref(p), C.c_double(beta),
                         _ptr(pv), _ptr(pc), _ptr(ft), _ptr(sh), _ptr(P))
    return P


def philox_uniforms(seed: int, sim: int, round_num: int, E: int) -> np.ndarray:
    u = np.zeros(E + 1, np.float64)
    lib().co_philox_uniforms(C.c_uint64(seed), C.c_uint64(sim), C.c_int(round_num), C.c_int(E), _ptr(u))
    return u[:E]


def run(roster, param_sets, wiring: int, n_sims: int, seed: int = 0, sim_id_begin: int = 0,
        tapes: Optional[np.ndarray] = None, want_tallies: bool = False, want_final: bool = True,
        threads: int = 0) -> dict:
    """param_sets: list of (ModelParams, EngineParams).  Returns dict of numpy arrays."""
    E = roster.E
    n_sets = len(param_sets)
    arr = (CParams * n_sets)(*[to_cparams(mp, ep) for mp, ep in param_sets])
    R = int(arr[0].max_rounds)
    n = n_sets * n_sims
    x = np.ascontiguousarray(roster.ideology, np.float64)
    g = np.ascontiguousarray(roster.region, np.int32)
    pap = np.ascontiguousarray(roster.papabile, np.uint8)
    winner = np.zeros(n, np.int32); rounds = np.zeros(n, np.int32); reason = np.zeros(n, np.uint8)
    final = np.zeros((n, E), np.int32) if want_final else None
    tall = np.zeros((n, R, E), np.int32) if want_tallies else None
    wh = np.zeros((n_sets, E + 1), np.int64); rh = np.zeros((n_sets, R + 1), np.int64); tot = np.zeros((n_sets, 2), np.int64)
    tp = None if tapes is None else np.ascontiguousarray(tapes, np.float64)
    if tp is not None:
        assert tp.shape == (n, R * E), tp.shape
    rc = lib().co_run(C.c_int(E), _ptr(x), _ptr(g), _ptr(pap), arr, C.c_int(n_sets), C.c_int(wiring),
                      C.c_uint64(sim_id_begin), C.c_uint64(n_sims), C.c_uint64(seed), _ptr(tp),
                      _ptr(winner), _ptr(rounds), _ptr(reason), _ptr(final), _ptr(tall),
                      _ptr(wh), _ptr(rh), _ptr(tot), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"co_run failed rc={rc}")
    return dict(winner=winner, rounds=rounds, reason=reason, final_votes=final, round_tallies=tall,
                winner_hist=wh, rounds_hist=rh, totals=tot)

"""CPU oracle: a NumPy restatement of the conclave voting engine.

TEST INFRASTRUCTURE ONLY.  Nothing under ``conclave_sim_b200/`` may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

It restates, in fp64 NumPy, the algorithm of the reference (file:line into
/root/reference):

  * ``beta_schedule``        <- src/model.py:562-580  (update_beta_weight)
  * ``prob_matrix``          <- src/model.py:303-486  (calculate_transition_probabilities)
  * ``choice_index``         <- numpy.random.RandomState.choice as called at
                                src/simulate.py:153 (third-party: numpy>=1.23,
                                requirements.txt:2; contract: cdf=cumsum(p);
                                cdf/=cdf[-1]; searchsorted(cdf,u,'right'))
  * ``need_votes``           <- src/simulate.py:175    (max >= thr*E)
  * ``top_k``                <- pandas.Series.nlargest as called at
                                src/simulate.py:185 (third-party: pandas>=1.5)
  * ``run_trial`` (ref_cli)  <- src/simulate.py:66-206 as wired by the CLI
  * ``run_trial`` (model)    <- same loop, with the model fed what its own
                                signature documents
  * ``fatigue_set``          <- src/simulate.py:84-118
  * ``aggregate``            <- src/simulate.py:356-387

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
against golden vectors produced by running the unmodified reference in the
build container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``), and
against the closed-form known answers of the reference's tests/test_model.py.

Wirings
-------
``ref_cli``: what ``python -m src.simulate`` actually computes.  The worker
passes the previous round's vote-COUNT series where the model expects an
elector->candidate map (simulate.py:122 vs model.py:326-367); with string ids
(ingest.py:653) both bandwagon and stickiness silently do nothing, and the
fatigue / stop-candidate flags are never forwarded (simulate.py:263-277), so
the probability matrix depends on (roster, params, round) only.
``model``: bandwagon, stickiness, fatigue and stop-candidate live, fed from
the trial's previous votes / tallies.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

WIRING_REF_CLI = 0
WIRING_MODEL = 1

REASON_MAX_ROUNDS = 0
REASON_WINNER = 1
REASON_NO_PROBS = 2


@dataclass
class Roster:
    ideology: np.ndarray  # f64 [E]
    region: np.ndarray    # int32 codes [E]
    papabile: np.ndarray  # bool [E]
    ids: Optional[List[str]] = None

    @property
    def E(self) -> int:
        return int(len(self.ideology))

    @staticmethod
    def from_arrays(ideology, region, papabile, ids=None) -> "Roster":
        region = np.asarray(region)
        if region.dtype.kind not in "iu":
            _, region = np.unique(region.astype(str), return_inverse=True)
        return Roster(np.asarray(ideology, np.float64), region.astype(np.int32),
                      np.asarray(papabile, bool), None if ids is None else [str(i) for i in ids])


@dataclass
class ModelParams:
    """Field names = TransitionModel.__init__ kwargs (src/model.py:172-191)."""
    initial_beta_weight: float = 1.0
    enable_dynamic_beta: bool = False
    beta_growth_rate: float = 1.05
    beta_increment_amount: Optional[float] = None
    beta_increment_interval_rounds: int = 1
    stickiness_factor: float = 0.0
    bandwagon_strength: float = 0.0
    regional_affinity_bonus: float = 0.0
    papabile_weight_factor: float = 1.0
    enable_candidate_fatigue: bool = False
    fatigue_threshold_rounds: int = 3
    fatigue_vote_share_threshold: float = 0.05
    fatigue_penalty_factor: float = 0.5
    fatigue_top_n_immune: int = 0
    enable_stop_candidate: bool = False
    stop_candidate_threshold_unacceptable_distance: float = 1.0
    stop_candidate_threat_min_vote_share: float = 0.15
    stop_candidate_boost_factor: float = 1.5


@dataclass
class EngineParams:
    """Arguments of _run_single_simulation_worker (src/simulate.py:30-41)."""
    max_rounds: int = 100
    supermajority_threshold: float = 2 / 3
    runoff_threshold_rounds: int = 5
    runoff_candidate_count: int = 2
    enable_candidate_fatigue: bool = False
    fatigue_threshold_rounds: int = 3
    fatigue_vote_share_threshold: float = 0.05
    fatigue_top_n_immune: int = 3
    enable_stop_candidate: bool = False


# --------------------------------------------------------------------------
# beta schedule  (src/model.py:562-580)
# --------------------------------------------------------------------------
def beta_schedule(mp: ModelParams, max_rounds: int, first_round: int = 1) -> np.ndarray:
    """beta[r-1] = effective beta used in round r, for r = first_round..max_rounds,
    when the model is called once per round in order (state carried, fp64
    repeated addition / multiplication exactly as the reference does)."""
    beta = float(mp.initial_beta_weight) if mp.initial_beta_weight is not None else 0.0
    out = np.empty(max(0, max_rounds - first_round + 1), dtype=np.float64)
    for k, r in enumerate(range(first_round, max_rounds + 1)):
        beta = beta_step(mp, beta, r)
        out[k] = beta
    return out


def beta_step(mp: ModelParams, beta: float, r: int) -> float:
    if not mp.enable_dynamic_beta or r <= 0:
        return beta
    if mp.beta_increment_amount is not None:
        iv = mp.beta_increment_interval_rounds
        if iv > 0 and r % iv == 0:
            beta = beta + mp.beta_increment_amount
    elif mp.beta_growth_rate > 1.0:
        beta = beta * mp.beta_growth_rate
    return beta


# --------------------------------------------------------------------------
# numpy's pairwise summation, restated (numpy/core/src/umath/loops_utils.h.src,
# DOUBLE_pairwise_sum; third-party, called via np.sum at model.py:448,451 and
# simulate.py:148).  Used to make the C / CUDA fp64 row sums bit-identical.
# --------------------------------------------------------------------------
def pairwise_sum(a: Sequence[float]) -> float:
    a = np.asarray(a, dtype=np.float64)
    n = len(a)
    if n < 8:
        res = np.float64(0.0)
        for v in a:
            res = res + v
        return float(res)
    if n <= 128:
        r = [np.float64(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return float(res)
    n2 = n // 2
    n2 -= n2 % 8
    return float(np.float64(pairwise_sum(a[:n2])) + np.float64(pairwise_sum(a[n2:])))


# --------------------------------------------------------------------------
# probability matrix  (src/model.py:303-486)
# --------------------------------------------------------------------------
def prob_matrix(
    roster: Roster,
    mp: ModelParams,
    beta: float,
    prev_vote: Optional[np.ndarray] = None,   # int [E], candidate index of elector's previous vote, -1 = none
    prev_count: Optional[np.ndarray] = None,  # int [C], previous-round counts for bandwagon (default: from prev_vote)
    fatigued: Optional[np.ndarray] = None,    # bool [C]
    shares: Optional[np.ndarray] = None,      # f64 [C], previous-round vote shares (stop-candidate)
    cand_index: Optional[np.ndarray] = None,  # int [n_cand]: the effective candidates (model.py:280-296), column order
) -> np.ndarray:
    """P[E, C] fp64 for one round with effective beta ``beta``.
    ``prev_vote`` / ``prev_count`` None == the reference's "no previous votes"
    (and == the ref_cli wiring in every round)."""
    x = roster.ideology
    E = roster.E
    if E == 0:
        return np.zeros((0, 0))
    if cand_index is not None:
        # votes for candidates outside the effective set are ignored (model.py:337,359)
        inside = np.zeros(E, bool)
        inside[np.asarray(cand_index)] = True
        if prev_vote is not None:
            pv0 = np.asarray(prev_vote).copy()
            if prev_count is None:
                prev_count = np.bincount(pv0[pv0 >= 0], minlength=E)
            pv0[(pv0 >= 0) & ~inside[np.clip(pv0, 0, E - 1)]] = -1
            prev_vote = pv0
        if prev_count is not None:
            prev_count = np.where(inside, np.asarray(prev_count), 0)
    d = np.abs(x[:, None] - x[None, :])                      # model.py:312
    s = np.exp(-beta * d)                                    # model.py:313
    pap = np.where(roster.papabile)[0]
    if pap.size:
        s[:, pap] *= mp.papabile_weight_factor               # model.py:317-319
    same = roster.region[:, None] == roster.region[None, :]
    s[same] += mp.regional_affinity_bonus                    # model.py:322-323
    if prev_count is None and prev_vote is not None:
        pv = np.asarray(prev_vote)
        prev_count = np.bincount(pv[pv >= 0], minlength=E)
    if mp.bandwagon_strength > 0 and prev_count is not None:  # model.py:326-354
        cnt = np.asarray(prev_count, dtype=np.int64)
        mx = int(cnt.max()) if cnt.size else 0
        if mx > 0:
            bonus = (cnt / mx) * mp.bandwagon_strength
            bonus[cnt == 0] = 0.0
            s = s + bonus[None, :]
    if prev_vote is not None and mp.stickiness_factor > 0:    # model.py:357-367
        pv = np.asarray(prev_vote)
        rows = np.where(pv >= 0)[0]
        s[rows, pv[rows]] *= (1 + mp.stickiness_factor)
    s = np.maximum(s, 0)                                      # model.py:369
    if mp.enable_candidate_fatigue and fatigued is not None and np.any(fatigued):
        s[:, np.where(fatigued)[0]] *= mp.fatigue_penalty_factor   # model.py:372-379
    if mp.enable_stop_candidate and shares is not None:       # model.py:384-415
        D = mp.stop_candidate_threshold_unacceptable_distance
        T = mp.stop_candidate_threat_min_vote_share
        unacc = d > D
        threat = np.asarray(shares) > T
        trig = (unacc & threat[None, :]).any(axis=1)
        s = s * np.where(trig[:, None] & ~unacc, mp.stop_candidate_boost_factor, 1.0)
    s[np.arange(E), np.arange(E)] = 0.0                       # model.py:418-424
    s = np.maximum(s, 0)                                      # model.py:427
    if cand_index is not None:
        s = np.ascontiguousarray(s[:, np.asarray(cand_index)])
    rs = s.sum(axis=1)
    zero = rs == 0
    if zero.any():
        s[zero, :] = 1e-9                                     # model.py:447-449
        rs = s.sum(axis=1)
    return s / rs[:, None]                                    # model.py:451,481


# --------------------------------------------------------------------------
# vote draw, threshold, run-off list
# --------------------------------------------------------------------------
def choice_index(p: np.ndarray, u: float) -> int:
    cdf = np.cumsum(p)
    cdf = cdf / cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def need_votes(threshold: float, E: int) -> int:
    """Smallest integer v with float64(v) >= float64(threshold) * E  (simulate.py:175)."""
    t = float(threshold) * E
    v = int(np.floor(t))
    while float(v) < t:
        v += 1
    while v > 0 and float(v - 1) >= t:
        v -= 1
    return max(v, 0)


def top_k(tally: np.ndarray, k: int) -> np.ndarray:
    """First k of a stable descending sort (ties -> lowest index) == Series.nlargest(k)."""
    order = np.argsort(-np.asarray(tally, dtype=np.int64), kind="stable")
    return order[:k].astype(np.int32)


def runoff_rows(P: np.ndarray, k: int) -> np.ndarray:
    """simulate.py:144-151: row = P[e, :k]; renormalise unless isclose(sum, 1)."""
    rows = P[:, :k].copy()
    out = np.empty_like(rows)
    for e in range(rows.shape[0]):
        row = rows[e]
        sm = np.sum(row)
        if not np.isclose(sm, 1.0):
            row = row / sm if sm > 0 else np.ones_like(row) / k
        out[e] = row
    return out


def cdf_rows(rows: np.ndarray) -> np.ndarray:
    cdf = np.cumsum(rows, axis=1)
    return cdf / cdf[:, -1:]


def fatigue_set(shares_hist: List[np.ndarray], r: int, ep: EngineParams, C: int) -> Tuple[np.ndarray, bool]:
    """simulate.py:84-118.  Returns (bool mask [C], boundary_tie flag).  On a tie
    across the immunity boundary the reference's pick depends on numpy's unstable
    argsort; this restatement takes the HIGHER indices among equals (what a
    stable ascending sort's tail gives) and reports the tie."""
    mask = np.zeros(C, dtype=bool)
    tie = False
    F = ep.fatigue_threshold_rounds
    if not ep.enable_candidate_fatigue or r <= F:
        return mask, tie
    immune = np.zeros(C, dtype=bool)
    n = ep.fatigue_top_n_immune
    if n > 0 and shares_hist:
        nh = min(len(shares_hist), F)
        if nh > 0:
            avg = np.mean(shares_hist[-nh:], axis=0)
            order = np.argsort(avg, kind="stable")
            immune[order[-n:]] = True
            srt = avg[order]
            if n < C and srt[-n] == srt[-n - 1]:
                tie = True
    if len(shares_hist) >= F and F > 0:
        rel = np.stack(shares_hist[-F:])
        low = (rel < ep.fatigue_vote_share_threshold).all(axis=0)
        mask = low & ~immune
    elif len(shares_hist) >= F and F == 0:
        mask = ~immune
    return mask, tie


# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11), the native RNG shared with the CUDA path
# key = (seed_lo, seed_hi);
#   call A: counter = (sim_lo, sim_hi, round, e >> 2)                -> word (e & 3) = w_a(e)
#   call B: counter = (sim_lo, sim_hi, round, (e >> 2) | 0x80000000) -> word (e & 3) = w_b(e)
#   u = ((w_a >> 5) * 2^26 + (w_b >> 6)) / 2^53   (same construction as MT19937's
#   random_sample, so tape and native modes share one sampling path); the fp32 fast
#   path of the CUDA engine uses (w_a >> 8) * 2^-24, the leading 24 bits of u.
# --------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = 0x9E3779B9
_PH_W1 = 0xBB67AE85
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64); c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64); c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PH_M0 * c0
        p1 = _PH_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _M32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _PH_W0) & 0xFFFFFFFF
        k1 = (k1 + _PH_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_words(seed: int, sim_id: int, round_num: int, E: int, second: bool = False) -> np.ndarray:
    nq = (E + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    if second:
        q = q | np.uint64(0x80000000)
    w = philox4x32_10(np.full(nq, sim_id & 0xFFFFFFFF, np.uint64),
                      np.full(nq, (sim_id >> 32) & 0xFFFFFFFF, np.uint64),
                      np.full(nq, round_num, np.uint64), q,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=1).reshape(-1)[:E]


def philox_uniforms(seed: int, sim_id: int, round_num: int, E: int) -> np.ndarray:
    """The E fp64 uniforms of (sim, round) in elector order."""
    a = philox_words(seed, sim_id, round_num, E, False)
    b = philox_words(seed, sim_id, round_num, E, True)
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def philox_uniforms_f32(seed: int, sim_id: int, round_num: int, E: int) -> np.ndarray:
    """The 24-bit uniforms the fp32 fast path consumes (leading bits of the fp64 ones)."""
    a = philox_words(seed, sim_id, round_num, E, False)
    return ((a >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def philox_tape(seed: int, sim_id: int, max_rounds: int, E: int) -> np.ndarray:
    return np.concatenate([philox_uniforms(seed, sim_id, r, E) for r in range(1, max_rounds + 1)])


# --------------------------------------------------------------------------
# one trial  (src/simulate.py:66-206)
# --------------------------------------------------------------------------
@dataclass
class TrialResult:
    winner: int               # candidate index, -1 = none
    rounds: int
    reason: int
    final_votes: np.ndarray   # int32 [C]
    tallies: np.ndarray       # int32 [rounds, C]
    votes: np.ndarray         # int32 [rounds, E]
    fatigue_tie: bool = False


class RoundTables:
    """Trial-invariant per-round tables of the ref_cli wiring, built lazily:
    P_r, full-field CDF rows, and run-off CDF rows for k leading columns."""

    def __init__(self, roster: Roster, mp: ModelParams, ep: EngineParams):
        self.roster, self.mp, self.ep = roster, mp, ep
        self.betas = beta_schedule(mp, ep.max_rounds)
        self._P: Dict[int, np.ndarray] = {}
        self._cdf: Dict[int, np.ndarray] = {}
        self._rcdf: Dict[int, np.ndarray] = {}

    def P(self, r: int) -> np.ndarray:
        if r not in self._P:
            self._P[r] = prob_matrix(self.roster, self.mp, self.betas[r - 1])
        return self._P[r]

    def cdf(self, r: int) -> np.ndarray:
        if r not in self._cdf:
            self._cdf[r] = cdf_rows(self.P(r))
        return self._cdf[r]

    def runoff_cdf(self, r: int) -> np.ndarray:
        if r not in self._rcdf:
            self._rcdf[r] = cdf_rows(runoff_rows(self.P(r), self.ep.runoff_candidate_count))
        return self._rcdf[r]


def run_trial(
    roster: Roster,
    mp: ModelParams,
    ep: EngineParams,
    tape: np.ndarray,
    wiring: int = WIRING_REF_CLI,
    tables: Optional[RoundTables] = None,
) -> TrialResult:
    """One trial consuming uniforms from ``tape`` (E per round, elector order)."""
    E = C = roster.E
    if E == 0:
        return TrialResult(-1, 1 if ep.max_rounds >= 1 else 0, REASON_NO_PROBS if ep.max_rounds >= 1 else REASON_MAX_ROUNDS,
                           np.zeros(0, np.int32), np.zeros((0, 0), np.int32), np.zeros((0, 0), np.int32))
    need = need_votes(ep.supermajority_threshold, E)
    k = ep.runoff_candidate_count
    eff: Optional[np.ndarray] = None
    tallies: List[np.ndarray] = []
    votes_all: List[np.ndarray] = []
    winner, rounds, reason = -1, 0, REASON_MAX_ROUNDS
    tally = np.zeros(C, np.int32)
    prev_vote = None
    shares_hist: List[np.ndarray] = []
    fatigue_tie = False
    beta = float(mp.initial_beta_weight) if mp.initial_beta_weight is not None else 0.0
    if wiring == WIRING_REF_CLI and tables is None:
        tables = RoundTables(roster, mp, ep)
    pos = 0
    for r in range(1, ep.max_rounds + 1):
        rounds = r
        if wiring == WIRING_REF_CLI:
            cdf = tables.cdf(r) if eff is None else tables.runoff_cdf(r)
        else:
            beta = beta_step(mp, beta, r)
            fat, tie = fatigue_set(shares_hist, r, ep, C)
            fatigue_tie |= tie
            shares_prev = shares_hist[-1] if shares_hist else np.zeros(C)
            P = prob_matrix(roster, mp, beta, prev_vote=prev_vote,
                            fatigued=fat, shares=shares_prev)
            cdf = cdf_rows(P) if eff is None else cdf_rows(runoff_rows(P, len(eff)))
        u = np.asarray(tape[pos:pos + E], dtype=np.float64)
        pos += E
        j = (cdf <= u[:, None]).sum(axis=1)
        votes = j.astype(np.int32) if eff is None else eff[j].astype(np.int32)
        tally = np.bincount(votes, minlength=C).astype(np.int32)
        tallies.append(tally)
        votes_all.append(votes)
        shares_hist.append(tally / float(E))
        prev_vote = votes
        if int(tally.max()) >= need:
            winner = int(np.argmax(tally))
            reason = REASON_WINNER
            break
        if r >= ep.runoff_threshold_rounds and k > 0:
            eff = top_k(tally, k)
    return TrialResult(winner, rounds, reason, tally.copy(),
                       np.stack(tallies) if tallies else np.zeros((0, C), np.int32),
                       np.stack(votes_all) if votes_all else np.zeros((0, E), np.int32),
                       fatigue_tie)


def run_batch(roster, mp, ep, n, seed=0, sim_id_begin=0, wiring=WIRING_REF_CLI, tapes=None):
    """n trials.  ``tapes`` [n, R*E] overrides the native Philox stream."""
    tables = RoundTables(roster, mp, ep) if wiring == WIRING_REF_CLI else None
    out = []
    for i in range(n):
        tape = tapes[i] if tapes is not None else philox_tape(seed, sim_id_begin + i, ep.max_rounds, roster.E)
        out.append(run_trial(roster, mp, ep, tape, wiring, tables))
    return out


# --------------------------------------------------------------------------
# aggregation  (src/simulate.py:356-387)
# --------------------------------------------------------------------------
def aggregate(winners: Sequence[Optional[str]], rounds: Sequence[int], num_simulations: int) -> dict:
    from collections import Counter

    ok = [i for i, w in enumerate(winners) if w is not None]
    rs = [int(rounds[i]) for i in ok]
    wc = Counter(winners[i] for i in ok)
    top = wc.most_common(1)
    return {
        "total_simulations": num_simulations,
        "successful_simulations": len(ok),
        "success_rate": len(ok) / num_simulations if num_simulations > 0 else 0,
        "average_rounds_successful": float(np.mean(rs)) if rs else None,
        "min_rounds_successful": int(np.min(rs)) if rs else None,
        "max_rounds_successful": int(np.max(rs)) if rs else None,
        "median_rounds_successful": float(np.median(rs)) if rs else None,
        "winner_counts": dict(wc),
        "most_frequent_winner": top[0][0] if top else None,
        "most_frequent_winner_count": top[0][1] if top else 0,
        "most_frequent_winner_percentage": ((top[0][1] if top else 0) / num_simulations) if num_simulations > 0 else 0,   # :371: 0 / n = 0.0 when nobody wins
        "average_rounds_all": float(np.mean([int(r) for r in rounds])) if len(rounds) else 0,
    }


# --------------------------------------------------------------------------
# synthetic rosters (SURVEY §8d recipe)
# --------------------------------------------------------------------------
def synthetic_roster(E: int, seed: int = 20250507) -> Roster:
    rng = np.random.default_rng(seed)
    ideology = rng.beta(2, 4, E)
    p = np.array([53, 23, 18, 18, 10, 5, 4, 3], dtype=np.float64)
    region = rng.choice(8, size=E, p=p / p.sum()).astype(np.int32)
    pap = rng.random(E) < 0.15
    return Roster(ideology, region, pap, [str(i) for i in range(E)])

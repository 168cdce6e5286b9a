"""Harness that imports the UNMODIFIED reference from /root/reference and runs it.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the CPU
restatements in ``oracle/conclave_oracle.py`` / ``oracle/conclave_oracle.c``
and (b) generate the golden vectors committed under ``tests/golden/``
(see ``oracle/gen_golden.py``).  It can only run in the build container,
where ``/root/reference`` is mounted; nothing that runs on the GPU box
imports it.  No reference code is copied: the reference modules are imported
as they lie, and only *patched from the outside* (``np.random.choice`` is
replaced by a tape reader, ``calculate_transition_probabilities`` is wrapped
to record its inputs/outputs).

Reference entry points exercised (file:line into /root/reference):
  * src/simulate.py:30-218   _run_single_simulation_worker  (one trial)
  * src/model.py:172-254     TransitionModel.__init__
  * src/model.py:269-560     TransitionModel.calculate_transition_probabilities
  * src/model.py:562-580     TransitionModel.update_beta_weight
"""
from __future__ import annotations

import logging
import os
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("CONCLAVE_REFERENCE_ROOT", "/root/reference")

_loaded = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "simulate.py"))


def load_reference():
    """Import ``src.model`` / ``src.simulate`` of the reference, unmodified.

    ``src/ingest.py:12-14`` imports bs4 / unidecode / google.generativeai at
    module top; none is installed here and none is used on the simulation
    path, so empty stub modules are registered.  pandas 3 would make the
    ``region`` column an Arrow string array, on which ``src/model.py:322``
    (2-D boolean indexing of ``.values``) raises; ``future.infer_string=False``
    restores the object dtype the reference was written against.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import pandas as pd

    try:
        pd.set_option("future.infer_string", False)
    except Exception:  # older pandas: option absent, behaviour already as needed
        pass
    for name, attrs in (
        ("bs4", {"BeautifulSoup": object}),
        ("unidecode", {"unidecode": lambda s: s}),
        ("google", {}),
        ("google.generativeai", {}),
    ):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    sys.modules["google"].generativeai = sys.modules["google.generativeai"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    prev_disable = logging.root.manager.disable
    import src.model as ref_model  # noqa: E402
    import src.simulate as ref_sim  # noqa: E402

    logging.disable(prev_disable)
    _loaded = (ref_model, ref_sim)
    return _loaded


def quiet():
    """The reference logs one WARNING per elector per round (model.py:526-530)."""
    logging.disable(logging.CRITICAL)


def load_roster_df(path: Optional[str] = None):
    """Roster exactly as src/ingest.py:625-663 prepares it (ids -> str, index)."""
    _, ref_sim = load_reference()
    path = path or os.path.join(REFERENCE_ROOT, "merged_electors.csv")
    return ref_sim.ElectorDataIngester(path).load_and_prepare_data()


def make_df(ideology, region, papabile, ids=None):
    import pandas as pd

    n = len(ideology)
    ids = [str(i) for i in range(n)] if ids is None else list(ids)
    df = pd.DataFrame(
        {
            "ideology_score": np.asarray(ideology, dtype=np.float64),
            "region": np.asarray(region, dtype=object),
            "is_papabile": np.asarray(papabile, dtype=bool),
        },
        index=pd.Index(ids, name="elector_id", dtype=object),
    )
    return df


class _Tape:
    """Replacement for ``np.random.choice(n, p=row)`` (simulate.py:153) that
    consumes uniforms from a supplied tape.  Mirrors numpy's legacy
    RandomState.choice: cdf = cumsum(p); cdf /= cdf[-1];
    idx = searchsorted(cdf, u, side='right')."""

    def __init__(self, tape: np.ndarray):
        self.tape = np.asarray(tape, dtype=np.float64)
        self.pos = 0
        self.rows: List[np.ndarray] = []

    def choice(self, n, p=None):
        u = self.tape[self.pos]
        self.pos += 1
        p = np.asarray(p, dtype=np.float64)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        return int(cdf.searchsorted(u, side="right"))


def mt_tape(seed: int, n: int) -> np.ndarray:
    """Uniforms the reference itself would draw: legacy global MT19937."""
    rs = np.random.RandomState(seed)
    return rs.random_sample(n)


def run_worker_with_tape(
    df,
    model_params: Dict[str, Any],
    tape: np.ndarray,
    max_rounds: int,
    supermajority_threshold: float,
    runoff_threshold_rounds: int = 5,
    runoff_candidate_count: int = 2,
    enable_candidate_fatigue: bool = False,
    fatigue_threshold_rounds: int = 3,
    fatigue_vote_share_threshold: float = 0.05,
    fatigue_top_n_immune: int = 3,
    enable_stop_candidate: bool = False,
    sim_id: int = 0,
    record_probs: bool = False,
):
    """Run the reference worker (simulate.py:30-218) for one trial, feeding its
    ``np.random.choice`` from ``tape``.  Returns (result_dict, per_round_tallies
    [rounds, C] int32, probs list or None, draws_used)."""
    ref_model, ref_sim = load_reference()
    t = _Tape(tape)
    tallies: List[np.ndarray] = []
    probs: List[np.ndarray] = []
    orig_calc = ref_model.TransitionModel.calculate_transition_probabilities

    def wrapped(self, previous_round_votes=None, *a, **kw):
        # previous_round_votes is the COUNT series of the previous round
        # (simulate.py:122); record it = tally of round r-1.
        if previous_round_votes is not None and len(previous_round_votes) > 0:
            tallies.append(np.asarray(previous_round_votes.values, dtype=np.int32).copy())
        out = orig_calc(self, previous_round_votes, *a, **kw)
        if record_probs:
            probs.append(np.array(out[0], dtype=np.float64, copy=True))
        return out

    orig_choice = np.random.choice
    ref_model.TransitionModel.calculate_transition_probabilities = wrapped
    np.random.choice = t.choice
    try:
        res = ref_sim._run_single_simulation_worker(
            sim_id, df, model_params, max_rounds, supermajority_threshold,
            runoff_threshold_rounds, runoff_candidate_count, False,
            enable_candidate_fatigue, fatigue_threshold_rounds,
            fatigue_vote_share_threshold, fatigue_top_n_immune, enable_stop_candidate,
        )
    finally:
        np.random.choice = orig_choice
        ref_model.TransitionModel.calculate_transition_probabilities = orig_calc
    cand = list(df.index)
    fv = res["final_votes"]
    if fv:
        tallies.append(np.array([fv[c] for c in cand], dtype=np.int32))
    tl = np.stack(tallies) if tallies else np.zeros((0, len(cand)), np.int32)
    return res, tl, (probs if record_probs else None), t.pos


def run_model_wiring_trial(
    df,
    model_kwargs: Dict[str, Any],
    tape: np.ndarray,
    max_rounds: int,
    supermajority_threshold: float,
    runoff_threshold_rounds: int = 5,
    runoff_candidate_count: int = 2,
    enable_candidate_fatigue: bool = False,
    fatigue_threshold_rounds: int = 3,
    fatigue_vote_share_threshold: float = 0.05,
    fatigue_top_n_immune: int = 3,
    enable_stop_candidate: bool = False,
    record_probs: bool = False,
):
    """`model` wiring: the round loop of
    simulate.py:66-206 restated here, but handing the reference's
    TransitionModel what its signature documents and its tests pin
    (tests/test_model.py:708-1000): ``previous_round_votes`` = elector-id ->
    candidate-id of the previous round, and the fatigue / stop-candidate flags
    forwarded to the model.  All probability arithmetic is the reference's own
    ``calculate_transition_probabilities``; only the loop is restated.

    Returns (winner_idx or -1, rounds, tallies [rounds, C], votes [rounds, E],
    probs list|None, fatigue_tie flag)."""
    import pandas as pd

    ref_model, _ = load_reference()
    kwargs = dict(model_kwargs)
    kwargs["enable_candidate_fatigue"] = enable_candidate_fatigue
    kwargs["enable_stop_candidate"] = enable_stop_candidate
    model = ref_model.TransitionModel(elector_data=df.copy(), **kwargs)
    cand = model.get_candidate_ids()
    C = len(cand)
    E = len(df)
    ids = list(df.index)
    pos = 0
    prev_votes = None
    shares_hist: List[np.ndarray] = []
    eff: Optional[List[int]] = None
    tallies, votes_all, probs = [], [], []
    winner, rounds = -1, 0
    fatigue_tie = False
    for r in range(1, max_rounds + 1):
        rounds = r
        fatigued = set()
        shares_prev = shares_hist[-1] if shares_hist else np.zeros(C)
        if enable_candidate_fatigue and r > fatigue_threshold_rounds:
            immune = set()
            if fatigue_top_n_immune > 0 and shares_hist:
                nh = min(len(shares_hist), fatigue_threshold_rounds)
                avg = np.mean(shares_hist[-nh:], axis=0)
                order = np.argsort(avg)
                immune.update(int(i) for i in order[-fatigue_top_n_immune:])
                # a tie across the immunity boundary makes the reference's own
                # choice depend on numpy's (unstable) argsort: flag it.
                srt = np.sort(avg)
                if fatigue_top_n_immune < C and srt[-fatigue_top_n_immune] == srt[-fatigue_top_n_immune - 1]:
                    fatigue_tie = True
            if len(shares_hist) >= fatigue_threshold_rounds:
                rel = shares_hist[-fatigue_threshold_rounds:]
                for c in range(C):
                    if c in immune:
                        continue
                    if all(rel[k][c] < fatigue_vote_share_threshold for k in range(fatigue_threshold_rounds)):
                        fatigued.add(c)
        P, _ = model.calculate_transition_probabilities(
            previous_round_votes=prev_votes,
            current_round_num=r,
            fatigued_candidate_indices=fatigued,
            candidate_vote_shares_current_round=shares_prev,
        )
        if record_probs:
            probs.append(np.array(P, copy=True))
        tally = np.zeros(C, dtype=np.int32)
        votes = np.zeros(E, dtype=np.int32)
        for e in range(E):
            if eff is None:
                row = P[e, :C]
            else:
                row = P[e, : len(eff)]
            if not np.isclose(np.sum(row), 1.0):
                s = np.sum(row)
                row = row / s if s > 0 else np.ones_like(row) / len(row)
            cdf = row.cumsum()
            cdf /= cdf[-1]
            j = int(cdf.searchsorted(tape[pos], side="right"))
            pos += 1
            c = j if eff is None else eff[j]
            votes[e] = c
            tally[c] += 1
        tallies.append(tally)
        votes_all.append(votes)
        shares_hist.append(tally / float(tally.sum()))
        prev_votes = pd.Series({ids[e]: cand[votes[e]] for e in range(E)}, dtype=object)
        if tally.max() >= supermajority_threshold * E:
            winner = int(np.argmax(tally))
            break
        if r >= runoff_threshold_rounds and runoff_candidate_count > 0:
            s = pd.Series(tally, index=cand)
            top = s.nlargest(runoff_candidate_count).index.tolist()
            eff = [cand.index(t) for t in top]
    return winner, rounds, np.stack(tallies), np.stack(votes_all), (probs if record_probs else None), fatigue_tie

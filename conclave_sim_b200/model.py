"""TransitionModel — mirror of src/model.py:133-586 with the arithmetic on the GPU.

Same constructor keywords, attributes, method names, return shapes and log messages as the
reference; `calculate_transition_probabilities` returns the fp64 matrix computed by the CUDA
library (csim_prob_matrix), which reproduces the reference's values to the last few ulps
(exp aside, every operation is the same IEEE operation in the same order).
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from . import _cabi
from .engine import Engine, get_engine, make_params
from .ingest import roster_arrays

log = logging.getLogger(__name__)


class TransitionModel:
    log = logging.getLogger(__name__ + ".TransitionModel")

    def _validate_elector_data(self, elector_data: pd.DataFrame) -> None:
        # src/model.py:148-170
        if elector_data.empty:
            return
        if elector_data.index.name != "elector_id":
            if "elector_id" not in elector_data.columns:
                self.log.warning(
                    "Elector data index is not named 'elector_id' and 'elector_id' column not found. "
                    "Model might not function as expected if index is not unique elector identifiers.")
            else:
                self.log.info(
                    "Elector data index is not named 'elector_id', but an 'elector_id' column was found. "
                    "Ensure the 'elector_id' column contains unique elector identifiers if the index is not used for this purpose.")
        elif "elector_id" in elector_data.columns:
            self.log.info(
                "Elector data index is named 'elector_id', and an 'elector_id' column also exists. "
                "Ensure the index correctly represents unique elector identifiers.")

    def __init__(self, elector_data: pd.DataFrame,
                 initial_beta_weight: float = 1.0,
                 enable_dynamic_beta: bool = False,
                 beta_growth_rate: float = 1.05,
                 beta_increment_amount: Optional[float] = None,
                 beta_increment_interval_rounds: int = 1,
                 stickiness_factor: float = 0.0,
                 bandwagon_strength: float = 0.0,
                 regional_affinity_bonus: float = 0.0,
                 papabile_weight_factor: float = 1.0,
                 enable_candidate_fatigue: bool = False,
                 fatigue_threshold_rounds: int = 3,
                 fatigue_vote_share_threshold: float = 0.05,
                 fatigue_penalty_factor: float = 0.5,
                 fatigue_top_n_immune: int = 0,
                 enable_stop_candidate: bool = False,
                 stop_candidate_threshold_unacceptable_distance: float = 1.0,
                 stop_candidate_threat_min_vote_share: float = 0.15,
                 stop_candidate_boost_factor: float = 1.5,
                 *, device: int = 0, precision: int = _cabi.FP64):
        self._validate_elector_data(elector_data)
        if elector_data.empty:
            log.warning("TransitionModel initialized with empty elector_data. Most operations will result in empty outputs.")
        self.elector_data = elector_data
        self.num_electors = len(elector_data)
        have_cols = (not elector_data.empty) and "ideology_score" in elector_data.columns and "region" in elector_data.columns
        self.elector_ideologies = elector_data["ideology_score"].values if "ideology_score" in elector_data else np.array([])
        self.elector_regions = elector_data["region"].values if "region" in elector_data else np.array([])
        if have_cols:
            self.candidate_ids = list(elector_data.index)
            self.master_candidate_ideologies = self.elector_ideologies
            self.master_candidate_regions = self.elector_regions
            if "is_papabile" in elector_data:
                self.master_candidate_is_papabile = elector_data["is_papabile"].astype(bool).values
            else:
                log.info("'is_papabile' column not found in elector_data. Assuming all candidates are papabile.")
                self.master_candidate_is_papabile = np.full(len(self.candidate_ids), True, dtype=bool)
        else:
            self.candidate_ids = []
            self.master_candidate_ideologies = np.array([])
            self.master_candidate_regions = np.array([])
            self.master_candidate_is_papabile = np.array([], dtype=bool)
        self.candidate_ids_set = set(self.candidate_ids)
        self.num_total_candidates = len(self.candidate_ids)

        self.initial_beta_weight = initial_beta_weight
        self.effective_beta_weight = float(initial_beta_weight) if initial_beta_weight is not None else 0.0
        self.enable_dynamic_beta = enable_dynamic_beta
        self.beta_growth_rate = beta_growth_rate
        self.beta_increment_amount = beta_increment_amount
        self.beta_increment_interval_rounds = beta_increment_interval_rounds
        self.stickiness_factor = stickiness_factor
        self.bandwagon_strength = bandwagon_strength
        self.regional_affinity_bonus = regional_affinity_bonus
        self.papabile_weight_factor = papabile_weight_factor
        self.enable_candidate_fatigue = enable_candidate_fatigue
        self.fatigue_threshold_rounds = fatigue_threshold_rounds
        self.fatigue_vote_share_threshold = fatigue_vote_share_threshold
        self.fatigue_penalty_factor = fatigue_penalty_factor
        self.fatigue_top_n_immune = fatigue_top_n_immune
        self.enable_stop_candidate = enable_stop_candidate
        self.stop_candidate_threshold_unacceptable_distance = stop_candidate_threshold_unacceptable_distance
        self.stop_candidate_threat_min_vote_share = stop_candidate_threat_min_vote_share
        self.stop_candidate_boost_factor = stop_candidate_boost_factor

        if self.enable_dynamic_beta:
            if self.beta_increment_amount is not None and self.beta_growth_rate not in (1.0, 1.05):
                log.warning("Both beta_increment_amount and a custom beta_growth_rate are set. Stepped increment (beta_increment_amount) will be prioritized.")
            elif self.beta_increment_amount is None and self.beta_growth_rate <= 1.0:
                log.warning(f"Dynamic beta enabled, beta_increment_amount is None, and beta_growth_rate ({self.beta_growth_rate}) is not > 1.0. Beta will not change multiplicatively.")
        elif self.initial_beta_weight is None:
            log.warning("Initial_beta_weight is None and dynamic_beta is not enabled. Effective_beta_weight defaults to 0.0. This might lead to uniform probabilities if not intended.")

        self._device = device
        self._precision = precision
        self._index_pos = {eid: i for i, eid in enumerate(elector_data.index)} if have_cols else {}
        self._cand_pos = {cid: i for i, cid in enumerate(self.candidate_ids)}

    # ------------------------------------------------------------------ helpers
    def model_kwargs(self) -> Dict[str, Any]:
        return dict(
            initial_beta_weight=self.initial_beta_weight, enable_dynamic_beta=self.enable_dynamic_beta,
            beta_growth_rate=self.beta_growth_rate, beta_increment_amount=self.beta_increment_amount,
            beta_increment_interval_rounds=self.beta_increment_interval_rounds, stickiness_factor=self.stickiness_factor,
            bandwagon_strength=self.bandwagon_strength, regional_affinity_bonus=self.regional_affinity_bonus,
            papabile_weight_factor=self.papabile_weight_factor, enable_candidate_fatigue=self.enable_candidate_fatigue,
            fatigue_threshold_rounds=self.fatigue_threshold_rounds,
            fatigue_vote_share_threshold=self.fatigue_vote_share_threshold,
            fatigue_penalty_factor=self.fatigue_penalty_factor, fatigue_top_n_immune=self.fatigue_top_n_immune,
            enable_stop_candidate=self.enable_stop_candidate,
            stop_candidate_threshold_unacceptable_distance=self.stop_candidate_threshold_unacceptable_distance,
            stop_candidate_threat_min_vote_share=self.stop_candidate_threat_min_vote_share,
            stop_candidate_boost_factor=self.stop_candidate_boost_factor)

    def _engine(self) -> Engine:
        eng = get_engine(self._device)
        x, g, pap = roster_arrays(self.elector_data)
        eng.upload_roster(x, g, pap)
        return eng

    # ------------------------------------------------------------------ API
    def calculate_transition_probabilities(self, previous_round_votes=None, current_round_num: int = 1,
                                           fatigued_candidate_indices: Optional[set] = None,
                                           candidate_vote_shares_current_round: Optional[np.ndarray] = None,
                                           effective_candidate_ids: Optional[List[Any]] = None):
        """(P[E, C] float64, []) — src/model.py:269-560."""
        cand_index = None        # roster indices of the candidates of this call, in column order (model.py:280-296)
        if effective_candidate_ids is not None and len(effective_candidate_ids) > 0:
            subset = [cid for cid in effective_candidate_ids if _hashable(cid) and cid in self.candidate_ids_set]
            if not subset:
                log.warning("Provided effective_candidate_ids resulted in an empty list after validation. Falling back to all candidates.")
            else:
                cand_index = np.array([self._cand_pos[cid] for cid in subset], dtype=np.int32)
        E = self.num_electors
        if E == 0 or self.num_total_candidates == 0:
            log.warning("No electors or actual candidates to calculate transition probabilities for.")
            return np.array([]).reshape(E, self.num_total_candidates), []

        self.update_beta_weight(current_round_num)
        prev_vote = prev_count = None
        if previous_round_votes is not None:
            items = None
            if isinstance(previous_round_votes, (dict, pd.Series)):
                items = list(previous_round_votes.items())
            else:
                logging.warning(f"Unsupported type for previous_round_votes in bandwagon: {type(previous_round_votes)}. Skipping effect.")
            if items is not None:
                # bandwagon counts every recorded vote for a known candidate (model.py:337-340);
                # stickiness needs the voter to be a known elector as well (model.py:358-363)
                prev_vote = np.full(E, -1, np.int32)
                prev_count = np.zeros(E, np.int32)
                in_subset = None if cand_index is None else set(int(i) for i in cand_index)
                for voter, chosen in items:
                    ci = self._cand_lookup(chosen)
                    if ci is None or (in_subset is not None and ci not in in_subset):
                        continue            # votes for candidates outside the effective set are ignored (model.py:337,359)
                    prev_count[ci] += 1
                    ei = self._index_pos.get(voter) if _hashable(voter) else None
                    if ei is not None:
                        prev_vote[ei] = ci
        fat = None
        if fatigued_candidate_indices:
            fat = np.zeros(E, bool)
            for i in fatigued_candidate_indices:
                if 0 <= int(i) < E:
                    fat[int(i)] = True
        shares = None
        if candidate_vote_shares_current_round is not None:
            shares = np.asarray(candidate_vote_shares_current_round, dtype=np.float64)
        params = make_params(self.model_kwargs())
        P = self._engine().prob_matrix(params, self.effective_beta_weight, prev_vote, prev_count, fat, shares,
                                       precision=self._precision, cand_index=cand_index)
        return P, []

    def _cand_lookup(self, cid):
        if not _hashable(cid):
            return None
        return self._cand_pos.get(cid)

    def update_beta_weight(self, current_round_num: int):
        """src/model.py:562-580 — stateful, fp64, same operations."""
        if not self.enable_dynamic_beta or current_round_num <= 0:
            return
        if self.beta_increment_amount is not None:
            iv = self.beta_increment_interval_rounds
            if iv > 0 and current_round_num % iv == 0:
                self.effective_beta_weight += self.beta_increment_amount
        elif self.beta_growth_rate > 1.0:
            self.effective_beta_weight *= self.beta_growth_rate

    def get_candidate_ids(self):
        return self.candidate_ids

    def get_elector_data(self):
        return self.elector_data


def _hashable(v) -> bool:
    try:
        hash(v)
        return True
    except TypeError:
        return False

"""Device context of the voting engine: roster upload (once), batched trials, probability matrix.

Replaces the per-trial process pool of the reference (src/simulate.py:330-351): one call runs a
whole batch of trials on the GPU.  All numerics happen in the CUDA library behind the C ABI
(include/conclave_b200.h); this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from ._cabi import CsimParams, CsimRunArgs, FP32, FP64, WIRING_MODEL, WIRING_REF_CLI, ENGINE_AUTO, check

# TransitionModel.__init__ defaults, src/model.py:172-191
MODEL_DEFAULTS: Dict[str, Any] = dict(
    initial_beta_weight=1.0, enable_dynamic_beta=False, beta_growth_rate=1.05, beta_increment_amount=None,
    beta_increment_interval_rounds=1, stickiness_factor=0.0, bandwagon_strength=0.0, regional_affinity_bonus=0.0,
    papabile_weight_factor=1.0, enable_candidate_fatigue=False, fatigue_threshold_rounds=3,
    fatigue_vote_share_threshold=0.05, fatigue_penalty_factor=0.5, fatigue_top_n_immune=0,
    enable_stop_candidate=False, stop_candidate_threshold_unacceptable_distance=1.0,
    stop_candidate_threat_min_vote_share=0.15, stop_candidate_boost_factor=1.5)


def make_params(model_params: Optional[Dict[str, Any]] = None, *, max_rounds: int = 100,
                supermajority_threshold: float = 2 / 3, runoff_threshold_rounds: int = 5,
                runoff_candidate_count: int = 2, enable_candidate_fatigue: Optional[bool] = None,
                fatigue_threshold_rounds: Optional[int] = None, fatigue_vote_share_threshold: Optional[float] = None,
                fatigue_top_n_immune: Optional[int] = None, enable_stop_candidate: Optional[bool] = None) -> CsimParams:
    """One csim_params from the reference's two argument groups: the TransitionModel kwargs dict
    (what model_params_from_args returns, src/simulate.py:263-277) and the worker arguments
    (src/simulate.py:30-41).  Worker-level fatigue / stop arguments, when given, override the
    model-level ones (the `model` wiring forwards them; under `ref_cli` the engine ignores both)."""
    unknown = set(model_params or {}) - set(MODEL_DEFAULTS)
    if unknown:
        raise TypeError(f"unexpected TransitionModel parameter(s): {sorted(unknown)}")
    mp = dict(MODEL_DEFAULTS)
    mp.update(model_params or {})
    p = CsimParams()
    p.initial_beta_weight = float(mp["initial_beta_weight"]) if mp["initial_beta_weight"] is not None else 0.0
    p.enable_dynamic_beta = int(bool(mp["enable_dynamic_beta"]))
    p.beta_mode = _cabi.BETA_GROWTH if mp["beta_increment_amount"] is None else _cabi.BETA_STEPPED
    p.beta_increment_amount = 0.0 if mp["beta_increment_amount"] is None else float(mp["beta_increment_amount"])
    p.beta_increment_interval_rounds = int(mp["beta_increment_interval_rounds"])
    p.beta_growth_rate = float(mp["beta_growth_rate"])
    p.stickiness_factor = float(mp["stickiness_factor"])
    p.bandwagon_strength = float(mp["bandwagon_strength"])
    p.regional_affinity_bonus = float(mp["regional_affinity_bonus"])
    p.papabile_weight_factor = float(mp["papabile_weight_factor"])
    p.fatigue_penalty_factor = float(mp["fatigue_penalty_factor"])
    p.stop_candidate_threshold_unacceptable_distance = float(mp["stop_candidate_threshold_unacceptable_distance"])
    p.stop_candidate_threat_min_vote_share = float(mp["stop_candidate_threat_min_vote_share"])
    p.stop_candidate_boost_factor = float(mp["stop_candidate_boost_factor"])
    pick = lambda override, key: mp[key] if override is None else override
    p.enable_candidate_fatigue = int(bool(pick(enable_candidate_fatigue, "enable_candidate_fatigue")))
    p.fatigue_threshold_rounds = int(pick(fatigue_threshold_rounds, "fatigue_threshold_rounds"))
    p.fatigue_vote_share_threshold = float(pick(fatigue_vote_share_threshold, "fatigue_vote_share_threshold"))
    p.fatigue_top_n_immune = int(pick(fatigue_top_n_immune, "fatigue_top_n_immune"))
    p.enable_stop_candidate = int(bool(pick(enable_stop_candidate, "enable_stop_candidate")))
    p.max_rounds = int(max_rounds)
    p.supermajority_threshold = float(supermajority_threshold)
    p.runoff_threshold_rounds = int(runoff_threshold_rounds)
    p.runoff_candidate_count = int(runoff_candidate_count)
    return p


@dataclass
class RunResult:
    """Per-trial arrays are laid out [set][sim] (flattened); histograms are [set, ...]."""
    winner: np.ndarray                  # int32 [n]   candidate index, -1 = no winner
    rounds: np.ndarray                  # int32 [n]
    reason: Optional[np.ndarray]        # uint8 [n]
    final_votes: Optional[np.ndarray]   # int32 [n, E]
    round_tallies: Optional[np.ndarray] # int32 [n, R, E], -1 beyond `rounds`
    winner_hist: np.ndarray             # int64 [sets, E+1]  (slot E = no winner)
    rounds_hist: np.ndarray             # int64 [sets, R+1]  (winning trials only)
    totals: np.ndarray                  # int64 [sets, 2]    (sum of rounds played, trials)
    n_electors: int
    kernel_ms: float = -1.0
    engine: str = ""                    # the engine that ran ("dense" | "table" | "prefix" | "full"): what AUTO resolved to
    precision: str = ""                 # the arithmetic it ran in ("fp32" | "fp64"); differs from the request only when AUTO
                                        # fell through to the fp64 exact engine
    fell_through: bool = False          # AUTO's first choice refused the shape
    devices: Tuple[int, ...] = ()       # CUDA devices the trials were sharded over
    kernel_ms_per_device: Tuple[float, ...] = ()

    @property
    def elector_rounds(self) -> int:
        return int(self.totals[:, 0].sum()) * self.n_electors


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """One page-locked allocation; goes back to its pool when the last array viewing it is collected."""
    __slots__ = ("pool", "ptr", "size")

    def __init__(self, pool, ptr, size):
        self.pool, self.ptr, self.size = pool, ptr, size

    def __del__(self):
        try:
            self.pool._give_back(self.ptr, self.size)
        except Exception:
            pass


class _PinnedPool:
    """Page-locked result arrays for large batches (csim_host_alloc).  cudaHostAlloc costs about a millisecond per call, so
    blocks are recycled by size once the numpy arrays built on them are gone; nothing is shared between live results."""
    MIN_BYTES = 1 << 18           # smaller results: pageable memory is just as fast

    def __init__(self, lib):
        self._lib = lib
        self._free: Dict[int, List[int]] = {}
        self._closed = False

    def array(self, n: int, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        nbytes = int(n) * dt.itemsize
        if nbytes < self.MIN_BYTES or self._closed:
            return np.empty(n, dt)
        size = 1 << (nbytes - 1).bit_length()                 # power-of-two buckets
        free = self._free.get(size)
        if free:
            ptr = free.pop()
        else:
            p = C.c_void_p()
            if self._lib.csim_host_alloc(size, C.byref(p)) != 0 or not p.value:
                return np.empty(n, dt)                         # pinning is an optimisation, never a requirement
            ptr = p.value
        buf = (C.c_ubyte * nbytes).from_address(ptr)
        buf._csim_block = _PinnedBlock(self, ptr, size)        # numpy array -> ctypes buffer -> block -> pool
        return np.frombuffer(buf, dtype=dt, count=n)

    def _give_back(self, ptr: int, size: int):
        if self._closed:
            self._lib.csim_host_free(C.c_void_p(ptr))
        else:
            self._free.setdefault(size, []).append(ptr)

    def close(self):
        self._closed = True
        for blocks in self._free.values():
            for ptr in blocks:
                self._lib.csim_host_free(C.c_void_p(ptr))
        self._free.clear()


class Engine:
    """One CUDA device.  Not a CPU fallback: construction fails without a usable sm_100 GPU."""

    def __init__(self, device: int = 0):
        self._lib = _cabi.load_library()
        self._ctx = C.c_void_p()
        check(self._lib.csim_create(C.byref(self._ctx), int(device)))
        self.device = int(device)
        self.n_electors = 0
        self._roster_key = None
        self._pinned = _PinnedPool(self._lib)

    def close(self):
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.csim_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ roster
    def upload_roster(self, ideology_score, region_code, is_papabile) -> None:
        """TransitionModel.__init__ captures exactly these three columns (src/model.py:199-213)."""
        x = np.ascontiguousarray(ideology_score, dtype=np.float64)
        g = np.ascontiguousarray(region_code, dtype=np.int32)
        p = np.ascontiguousarray(np.asarray(is_papabile).astype(bool), dtype=np.uint8)
        if not (x.ndim == g.ndim == p.ndim == 1 and len(x) == len(g) == len(p)):
            raise ValueError("roster columns must be 1-D and of equal length")
        key = (x.tobytes(), g.tobytes(), p.tobytes())
        if key == self._roster_key:
            return
        check(self._lib.csim_upload_roster(self._ctx, len(x), _ptr(x), _ptr(g), _ptr(p)))
        self.n_electors = len(x)
        self._roster_key = key

    def device_info(self) -> dict:
        sm, ma, mi, khz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self._lib.csim_device_info(self._ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(khz)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), clock_khz=khz.value)

    # ------------------------------------------------------------------ model
    def beta_schedule(self, params: CsimParams, first_round: int, n_rounds: int) -> np.ndarray:
        out = np.zeros(n_rounds, np.float64)
        check(self._lib.csim_beta_schedule(C.byref(params), first_round, n_rounds, _ptr(out)))
        return out

    def need_votes(self, threshold: float, n_electors: Optional[int] = None) -> int:
        return int(self._lib.csim_need_votes(float(threshold), int(self.n_electors if n_electors is None else n_electors)))

    def prob_matrix(self, params: CsimParams, beta: float, prev_vote=None, prev_count=None, fatigued=None,
                    shares=None, precision: int = FP64, cand_index=None) -> np.ndarray:
        """P[E, n_cand]; `cand_index` = roster indices of the effective candidates (None = all)."""
        E = self.n_electors
        pv = None if prev_vote is None else np.ascontiguousarray(prev_vote, np.int32)
        pc = None if prev_count is None else np.ascontiguousarray(prev_count, np.int32)
        ft = None if fatigued is None else np.ascontiguousarray(np.asarray(fatigued).astype(bool), np.uint8)
        sh = None if shares is None else np.ascontiguousarray(shares, np.float64)
        for a in (pv, pc, ft, sh):
            if a is not None and a.shape != (E,):
                raise ValueError(f"per-candidate / per-elector inputs must have shape ({E},)")
        ci = None if cand_index is None else np.ascontiguousarray(cand_index, np.int32)
        nc = E if ci is None else len(ci)
        P = np.empty((E, nc), np.float64 if precision == FP64 else np.float32)
        check(self._lib.csim_prob_matrix(self._ctx, C.byref(params), float(beta), _ptr(pv), _ptr(pc), _ptr(ft), _ptr(sh),
                                         _ptr(ci), int(nc), int(precision), _ptr(P)))
        return P

    # ------------------------------------------------------------------ trials
    def run(self, param_sets: Sequence[CsimParams], n_sims: int, *, sim_id_begin: int = 0, seed: int = 0,
            wiring: int = WIRING_REF_CLI, engine: int = ENGINE_AUTO, precision: int = FP32,
            tapes: Optional[np.ndarray] = None, want_reason: bool = True, want_final_votes: bool = False,
            want_round_tallies: bool = False, out: Optional[Dict[str, np.ndarray]] = None) -> RunResult:
        """Host-buffer call: copies in, runs, copies out, synchronises.
        `out` may supply preallocated result arrays ('winner', 'rounds', 'reason': int32 / int32 / uint8 of n trials),
        e.g. views of page-locked memory, so that the device-to-host copies run at full PCIe speed."""
        E = self.n_electors
        sets = (CsimParams * len(param_sets))(*param_sets)
        R = int(sets[0].max_rounds)
        n = len(param_sets) * int(n_sims)
        def _take(name, dtype):
            a = None if out is None else out.get(name)
            if a is None:
                return self._pinned.array(n, dtype)      # every entry is written by the device-to-host copy; page-locked when large
            if a.dtype != dtype or a.shape != (n,) or not a.flags.c_contiguous:
                raise ValueError(f"out['{name}'] must be a contiguous {np.dtype(dtype).name} array of shape ({n},)")
            return a
        winner = _take("winner", np.int32)
        rounds = _take("rounds", np.int32)
        reason = _take("reason", np.uint8) if want_reason else None
        final = np.empty((n, E), np.int32) if want_final_votes else None
        tall = np.empty((n, R, E), np.int32) if want_round_tallies else None
        wh = np.zeros((len(param_sets), E + 1), np.int64)
        rh = np.zeros((len(param_sets), R + 1), np.int64)
        tot = np.zeros((len(param_sets), 2), np.int64)
        tp = None
        if tapes is not None:
            tp = np.ascontiguousarray(tapes, np.float64)
            if tp.shape != (n, R * E):
                raise ValueError(f"uniform tape must have shape ({n}, {R * E}), got {tp.shape}")
        a = CsimRunArgs()
        a.params = sets; a.n_param_sets = len(param_sets); a.wiring = wiring; a.engine = engine; a.precision = precision
        a.sim_id_begin = int(sim_id_begin); a.n_sims = int(n_sims); a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.uniform_tape = _ptr(tp)
        a.winner = _ptr(winner); a.rounds = _ptr(rounds); a.reason = _ptr(reason)
        a.final_votes = _ptr(final); a.round_tallies = _ptr(tall)
        a.winner_hist = _ptr(wh); a.rounds_hist = _ptr(rh); a.totals = _ptr(tot)
        a.buffers_on_device = 0; a.stream = None
        check(self._lib.csim_run(self._ctx, C.byref(a)))
        eng_name, prec_name, fell = self.last_run_info(requested_precision=precision) if n else ("", "", False)
        ms = self.stats()[1]
        return RunResult(winner, rounds, reason, final, tall, wh, rh, tot, E, ms, eng_name, prec_name, fell, (self.device,), (ms,))

    def run_device(self, param_sets: Sequence[CsimParams], n_sims: int, *, winner, rounds, winner_hist=None,
                   rounds_hist=None, totals=None, reason=None, final_votes=None, round_tallies=None,
                   uniform_tape=None, sim_id_begin: int = 0, seed: int = 0, wiring: int = WIRING_REF_CLI,
                   engine: int = ENGINE_AUTO, precision: int = FP32, stream: int = 0, auto_batch_sims: int = 0) -> None:
        """Device-buffer call: every array argument is a torch CUDA tensor (or anything with
        data_ptr()) on this engine's device; work is only ENQUEUED on `stream` (a raw cudaStream_t
        handle, e.g. torch.cuda.current_stream().cuda_stream).  Histograms are accumulated into."""
        sets = (CsimParams * len(param_sets))(*param_sets)
        dp = lambda t: None if t is None else C.c_void_p(int(t.data_ptr()))
        a = CsimRunArgs()
        a.params = sets; a.n_param_sets = len(param_sets); a.wiring = wiring; a.engine = engine; a.precision = precision
        a.sim_id_begin = int(sim_id_begin); a.n_sims = int(n_sims); a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.uniform_tape = dp(uniform_tape)
        a.winner = dp(winner); a.rounds = dp(rounds); a.reason = dp(reason)
        a.final_votes = dp(final_votes); a.round_tallies = dp(round_tallies)
        a.winner_hist = dp(winner_hist); a.rounds_hist = dp(rounds_hist); a.totals = dp(totals)
        a.buffers_on_device = 1; a.stream = C.c_void_p(int(stream)) if stream else None
        a.auto_batch_sims = int(auto_batch_sims)       # the whole job's trials per set when this batch is one rank's shard (AUTO's choice)
        check(self._lib.csim_run(self._ctx, C.byref(a)))

    def last_run_info(self, requested_precision: Optional[int] = None):
        """(engine name, precision name, fell_through) of the last csim_run: what CSIM_ENGINE_AUTO resolved to.  Warns when the
        fall-through changed the arithmetic (an fp32 request served by the fp64 exact engine, ~100x slower)."""
        return _run_info(self._lib, self._ctx, requested_precision)

    def engine_probs(self, params: CsimParams, *, wiring: int, engine: int, round: int = 1, prev_vote=None, prev_count=None,
                     fatigued=None):
        """Parity probe (csim_engine_probs): the scores and chunk prefixes the named fp32 TRIAL engine uses in its hot loop for one
        full-field round.  Returns (scores[E, E] float32, prefix[E, n_chunks] float32, chunk)."""
        E = self.n_electors
        pv = None if prev_vote is None else np.ascontiguousarray(prev_vote, np.int32)
        pc = None if prev_count is None else np.ascontiguousarray(prev_count, np.int32)
        ft = None if fatigued is None else np.ascontiguousarray(np.asarray(fatigued).astype(bool), np.uint8)
        scores = np.zeros((E, E), np.float32)
        cap = max(E, 32)
        prefix = np.zeros((E, cap), np.float32)
        nch, chunk = C.c_int32(), C.c_int32()
        check(self._lib.csim_engine_probs(self._ctx, C.byref(params), int(wiring), int(engine), int(round), _ptr(pv), _ptr(pc), _ptr(ft),
                                          _ptr(scores), _ptr(prefix), cap, C.byref(nch), C.byref(chunk)))
        return scores, prefix[:, :nch.value].copy(), chunk.value

    # ------------------------------------------------------------------ multi-process NCCL (one process per GPU)
    def nccl_comm_create(self, unique_id: bytes, n_ranks: int, rank: int) -> int:
        """ncclComm_t (as an int) of this rank on this engine's device; `unique_id` = the 128 bytes rank 0 got from
        nccl_unique_id() and broadcast to every rank."""
        _cabi.nccl_load()
        buf = (C.c_uint8 * _cabi.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        comm = C.c_void_p()
        check(self._lib.csim_nccl_comm_create(self._ctx, buf, int(n_ranks), int(rank), C.byref(comm)))
        return comm.value

    def allreduce_hists(self, comm: int, winner_hist, rounds_hist, totals, n_sets: int, max_rounds: int, stream: int = 0) -> None:
        """csim_allreduce_hists: in-place NCCL sum of this rank's DEVICE histograms (anything with data_ptr()), enqueued on `stream`."""
        dp = lambda t: None if t is None else C.c_void_p(int(t.data_ptr()))
        check(self._lib.csim_allreduce_hists(self._ctx, C.c_void_p(comm), dp(winner_hist), dp(rounds_hist), dp(totals), int(n_sets),
                                             int(max_rounds), C.c_void_p(int(stream)) if stream else None))

    # ------------------------------------------------------------------ diagnostics
    def stats(self):
        n, ms = C.c_int64(), C.c_float()
        check(self._lib.csim_stats(self._ctx, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def microbench(self, kind: int) -> float:
        g = C.c_double()
        check(self._lib.csim_microbench(self._ctx, int(kind), C.byref(g)))
        return g.value


def _run_info(lib, ctx, requested_precision: Optional[int]):
    e, p, f = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.csim_last_run_info(ctx, C.byref(e), C.byref(p), C.byref(f)))
    if requested_precision is not None and p.value != requested_precision:
        warnings.warn(f"conclave_sim_b200: no fp32 engine runs this batch (shape / parameters); CSIM_ENGINE_AUTO fell through to the "
                      f"fp64 exact engine '{_cabi.ENGINE_NAMES[e.value]}' — results are of fp64 quality but ~100x slower", RuntimeWarning,
                      stacklevel=3)
    return _cabi.ENGINE_NAMES[e.value], "fp64" if p.value == FP64 else "fp32", bool(f.value)


def nccl_unique_id() -> bytes:
    """128-byte NCCL id (rank 0 calls this and broadcasts it; see Engine.nccl_comm_create)."""
    _cabi.nccl_load()
    buf = (C.c_uint8 * _cabi.NCCL_ID_BYTES)()
    check(_cabi.load_library().csim_nccl_unique_id(buf))
    return bytes(buf)


def nccl_comm_destroy(comm: int) -> None:
    check(_cabi.load_library().csim_nccl_comm_destroy(C.c_void_p(comm)))


class EngineGroup:
    """Several CUDA devices driven from ONE process (csim_group_*): the trials of a call are sharded over the devices by disjoint
    sim-id ranges, the histograms are combined by one NCCL all-reduce and the per-trial rows land in one host array — the
    replacement of the reference's `-w/--workers` process pool (src/simulate.py:231,330-351).  Same `run` as Engine; results
    are bit-identical to a single-device run."""

    def __init__(self, devices: Sequence[int]):
        self._lib = _cabi.load_library()
        self.devices = tuple(int(d) for d in devices)
        if len(self.devices) > 1:
            _cabi.nccl_load()
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        self._grp = C.c_void_p()
        check(self._lib.csim_group_create(C.byref(self._grp), arr, len(self.devices)))
        self.n_electors = 0
        self._roster_key = None
        self._pinned = _PinnedPool(self._lib)

    def close(self):
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
        if getattr(self, "_grp", None) is not None and self._grp:
            self._lib.csim_group_destroy(self._grp)
            self._grp = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_roster(self, ideology_score, region_code, is_papabile) -> None:
        x = np.ascontiguousarray(ideology_score, dtype=np.float64)
        g = np.ascontiguousarray(region_code, dtype=np.int32)
        p = np.ascontiguousarray(np.asarray(is_papabile).astype(bool), dtype=np.uint8)
        if not (x.ndim == g.ndim == p.ndim == 1 and len(x) == len(g) == len(p)):
            raise ValueError("roster columns must be 1-D and of equal length")
        key = (x.tobytes(), g.tobytes(), p.tobytes())
        if key == self._roster_key:
            return
        check(self._lib.csim_group_upload_roster(self._grp, len(x), _ptr(x), _ptr(g), _ptr(p)))
        self.n_electors = len(x)
        self._roster_key = key

    def run(self, param_sets: Sequence[CsimParams], n_sims: int, *, sim_id_begin: int = 0, seed: int = 0,
            wiring: int = WIRING_REF_CLI, engine: int = ENGINE_AUTO, precision: int = FP32,
            tapes: Optional[np.ndarray] = None, want_reason: bool = True, want_final_votes: bool = False,
            want_round_tallies: bool = False) -> RunResult:
        E = self.n_electors
        sets = (CsimParams * len(param_sets))(*param_sets)
        R = int(sets[0].max_rounds)
        n = len(param_sets) * int(n_sims)
        winner = self._pinned.array(n, np.int32)
        rounds = self._pinned.array(n, np.int32)
        reason = self._pinned.array(n, np.uint8) if want_reason else None
        final = np.empty((n, E), np.int32) if want_final_votes else None
        tall = np.empty((n, R, E), np.int32) if want_round_tallies else None
        wh = np.zeros((len(param_sets), E + 1), np.int64)
        rh = np.zeros((len(param_sets), R + 1), np.int64)
        tot = np.zeros((len(param_sets), 2), np.int64)
        tp = None
        if tapes is not None:
            tp = np.ascontiguousarray(tapes, np.float64)
            if tp.shape != (n, R * E):
                raise ValueError(f"uniform tape must have shape ({n}, {R * E}), got {tp.shape}")
        a = CsimRunArgs()
        a.params = sets; a.n_param_sets = len(param_sets); a.wiring = wiring; a.engine = engine; a.precision = precision
        a.sim_id_begin = int(sim_id_begin); a.n_sims = int(n_sims); a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.uniform_tape = _ptr(tp)
        a.winner = _ptr(winner); a.rounds = _ptr(rounds); a.reason = _ptr(reason)
        a.final_votes = _ptr(final); a.round_tallies = _ptr(tall)
        a.winner_hist = _ptr(wh); a.rounds_hist = _ptr(rh); a.totals = _ptr(tot)
        a.buffers_on_device = 0; a.stream = None
        kms = (C.c_float * len(self.devices))()
        check(self._lib.csim_group_run(self._grp, C.byref(a), kms))
        eng_name = prec_name = ""
        fell = False
        if n:
            # every device resolves AUTO the same way unless its shard is much smaller; report the first device that ran
            for i in range(len(self.devices)):
                try:
                    eng_name, prec_name, fell = _run_info(self._lib, self._lib.csim_group_ctx(self._grp, i), precision)
                    break
                except _cabi.CsimError:
                    continue
        ms = tuple(float(v) for v in kms)
        return RunResult(winner, rounds, reason, final, tall, wh, rh, tot, E, max(ms) if ms else -1.0, eng_name, prec_name, fell,
                         self.devices, ms)


_engines: Dict[int, Engine] = {}
_groups: Dict[Tuple[int, ...], EngineGroup] = {}


def get_group(devices: Sequence[int]) -> EngineGroup:
    """Process-wide engine group per device tuple."""
    key = tuple(int(d) for d in devices)
    if key not in _groups:
        _groups[key] = EngineGroup(key)
    return _groups[key]


def device_count() -> int:
    """Visible CUDA devices (csim_device_count; 0 when the driver reports none)."""
    n = C.c_int32(0)
    if _cabi.load_library().csim_device_count(C.byref(n)) != 0:
        return 0
    return int(n.value)


def get_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the worker mirror is called once per trial)."""
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]

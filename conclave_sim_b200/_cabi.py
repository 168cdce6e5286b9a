"""ctypes binding of include/conclave_b200.h.  No torch types cross this boundary.

The product path has no CPU fallback: if the shared library is missing, or no CUDA device is
usable, calls raise (CsimError / RuntimeError) instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSIM_LIB") or os.path.join(HERE, "libconclave_b200.so")   # CSIM_LIB: tuning builds only

WIRING_REF_CLI, WIRING_MODEL = 0, 1
ENGINE_AUTO, ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX, ENGINE_FULL = 0, 1, 2, 3, 4
FP32, FP64 = 0, 1
REASON_MAX_ROUNDS, REASON_WINNER, REASON_NO_PROBS = 0, 1, 2
BETA_STEPPED, BETA_GROWTH = 0, 1
ABI_VERSION = 2
NCCL_ID_BYTES = 128
ENGINE_NAMES = {0: "auto", 1: "dense", 2: "table", 3: "prefix", 4: "full"}

EXPORTS = ["csim_abi_version", "csim_last_error", "csim_create", "csim_destroy", "csim_device_info",
           "csim_upload_roster", "csim_beta_schedule", "csim_need_votes", "csim_prob_matrix", "csim_run",
           "csim_stats", "csim_microbench", "csim_prefix_plan_query", "csim_prefix_slot", "csim_full_plan_query", "csim_host_alloc", "csim_host_free", "csim_vote_threshold",
           "csim_last_run_info", "csim_engine_probs", "csim_device_count",
           "csim_nccl_load", "csim_nccl_version", "csim_nccl_unique_id", "csim_nccl_comm_create", "csim_nccl_comm_destroy",
           "csim_allreduce_hists", "csim_shard_range",
           "csim_group_create", "csim_group_destroy", "csim_group_size", "csim_group_ctx", "csim_group_upload_roster", "csim_group_run"]


class CsimError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"conclave_b200 error {status}: {message}")
        self.status = status


class CsimParams(C.Structure):
    """struct csim_params"""
    _fields_ = [(n, C.c_double) for n in (
        "initial_beta_weight", "beta_increment_amount", "beta_growth_rate", "stickiness_factor",
        "bandwagon_strength", "regional_affinity_bonus", "papabile_weight_factor",
        "fatigue_vote_share_threshold", "fatigue_penalty_factor",
        "stop_candidate_threshold_unacceptable_distance", "stop_candidate_threat_min_vote_share",
        "stop_candidate_boost_factor", "supermajority_threshold")] + [(n, C.c_int32) for n in (
        "enable_dynamic_beta", "beta_mode", "beta_increment_interval_rounds", "enable_candidate_fatigue",
        "fatigue_threshold_rounds", "fatigue_top_n_immune", "enable_stop_candidate", "max_rounds",
        "runoff_threshold_rounds", "runoff_candidate_count", "reserved0", "reserved1")]


class CsimRunArgs(C.Structure):
    """struct csim_run_args"""
    _fields_ = [
        ("params", C.POINTER(CsimParams)), ("n_param_sets", C.c_int32), ("wiring", C.c_int32),
        ("engine", C.c_int32), ("precision", C.c_int32),
        ("sim_id_begin", C.c_uint64), ("n_sims", C.c_uint64), ("seed", C.c_uint64),
        ("uniform_tape", C.c_void_p),
        ("winner", C.c_void_p), ("rounds", C.c_void_p), ("reason", C.c_void_p),
        ("final_votes", C.c_void_p), ("round_tallies", C.c_void_p),
        ("winner_hist", C.c_void_p), ("rounds_hist", C.c_void_p), ("totals", C.c_void_p),
        ("buffers_on_device", C.c_int32), ("reserved", C.c_int32), ("stream", C.c_void_p),
        ("auto_batch_sims", C.c_uint64),
    ]


class CsimPrefixPlan(C.Structure):
    """struct csim_prefix_plan"""
    _fields_ = [(n, C.c_int32) for n in ("staged", "chunk", "n_chunks", "row_pitch", "rows", "table_rounds", "warps_per_cta",
                                         "e_quad_end")] + \
               [(n, C.c_uint64) for n in ("table_bytes_per_set", "smem_tables", "smem_cta", "smem_per_warp", "smem_total")] + \
               [("offsets", C.c_uint32 * 12)]


class CsimFullPlan(C.Structure):
    """struct csim_full_plan"""
    _fields_ = [(n, C.c_int32) for n in ("fast_path", "general_fits", "warps_per_cta", "cta_tabs", "n_chunks", "acc_pitch")] + \
               [(n, C.c_uint64) for n in ("smem_cta", "smem_per_warp", "smem_total")]


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  conclave_sim_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.csim_abi_version.restype = C.c_int
    lib.csim_last_error.restype = C.c_char_p
    lib.csim_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.csim_destroy.argtypes = [C.c_void_p]
    lib.csim_device_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    lib.csim_upload_roster.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.csim_beta_schedule.argtypes = [C.POINTER(CsimParams), C.c_int32, C.c_int32, C.c_void_p]
    lib.csim_need_votes.argtypes = [C.c_double, C.c_int32]
    lib.csim_need_votes.restype = C.c_int32
    lib.csim_vote_threshold.argtypes = [C.c_double, C.c_int32, C.c_int32]
    lib.csim_vote_threshold.restype = C.c_int32
    lib.csim_prob_matrix.argtypes = [C.c_void_p, C.POINTER(CsimParams), C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.csim_run.argtypes = [C.c_void_p, C.POINTER(CsimRunArgs)]
    lib.csim_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    lib.csim_microbench.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
    lib.csim_prefix_plan_query.argtypes = [C.c_int32] * 6 + [C.c_uint64, C.c_int32, C.POINTER(CsimPrefixPlan)]
    lib.csim_prefix_plan_query.restype = C.c_int
    lib.csim_full_plan_query.argtypes = [C.c_int32] * 7 + [C.POINTER(CsimFullPlan)]
    lib.csim_full_plan_query.restype = C.c_int
    lib.csim_host_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
    lib.csim_host_alloc.restype = C.c_int
    lib.csim_host_free.argtypes = [C.c_void_p]
    lib.csim_host_free.restype = C.c_int
    lib.csim_prefix_slot.argtypes = [C.c_int32, C.c_int32]
    lib.csim_prefix_slot.restype = C.c_int32
    lib.csim_device_count.argtypes = [C.POINTER(C.c_int32)]
    lib.csim_device_count.restype = C.c_int
    lib.csim_last_run_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
    lib.csim_engine_probs.argtypes = [C.c_void_p, C.POINTER(CsimParams), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.csim_nccl_load.argtypes = [C.c_char_p]
    lib.csim_nccl_version.restype = C.c_int32
    lib.csim_nccl_unique_id.argtypes = [C.c_void_p]
    lib.csim_nccl_comm_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.csim_nccl_comm_destroy.argtypes = [C.c_void_p]
    lib.csim_allreduce_hists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.csim_shard_range.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.csim_shard_range.restype = None
    lib.csim_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32]
    lib.csim_group_destroy.argtypes = [C.c_void_p]
    lib.csim_group_size.argtypes = [C.c_void_p]
    lib.csim_group_size.restype = C.c_int32
    lib.csim_group_ctx.argtypes = [C.c_void_p, C.c_int32]
    lib.csim_group_ctx.restype = C.c_void_p
    lib.csim_group_upload_roster.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.csim_group_run.argtypes = [C.c_void_p, C.POINTER(CsimRunArgs), C.c_void_p]
    for name in ("csim_create", "csim_destroy", "csim_device_info", "csim_upload_roster", "csim_beta_schedule",
                 "csim_prob_matrix", "csim_run", "csim_stats", "csim_microbench", "csim_last_run_info", "csim_engine_probs",
                 "csim_nccl_load", "csim_nccl_unique_id", "csim_nccl_comm_create", "csim_nccl_comm_destroy", "csim_allreduce_hists",
                 "csim_group_create", "csim_group_destroy", "csim_group_upload_roster", "csim_group_run"):
        getattr(lib, name).restype = C.c_int
    v = lib.csim_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {v}, binding expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load_library().csim_last_error()
        raise CsimError(status, msg.decode("utf-8", "replace") if msg else "")


def nccl_library_path():
    """The NCCL copy the process should bind: the wheel-bundled one next to torch (the copy torch itself loads, so one NCCL
    per process), found WITHOUT importing torch; None -> csim_nccl_load falls back to the loader's search path."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    return None


def nccl_load() -> int:
    """Bind NCCL (idempotent); returns its version code.  Raises CsimError when no NCCL can be loaded."""
    lib = load_library()
    path = os.environ.get("CSIM_NCCL_LIB") or nccl_library_path()
    check(lib.csim_nccl_load(path.encode() if path else None))
    return int(lib.csim_nccl_version())

"""conclave_sim_b200 — B200-native Monte Carlo voting engine, drop-in for the hot path of
potalora/conclave_sim (src/simulate.py per-round body + src/model.py transition model).

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/conclave_b200.h)
  _cabi.py         ctypes binding of the C ABI (fails loudly when the library is missing)
  engine.py        device context: roster upload, batched trials, probability matrix
  model.py         TransitionModel           (mirror of src/model.py:133-586)
  simulate.py      parse_args / model_params_from_args / _run_single_simulation_worker / main
                   (mirror of src/simulate.py)
  ingest.py        ElectorDataIngester        (mirror of src/ingest.py:611-663)
  dist.py          one process per GPU (torchrun): trial sharding + the histogram all-reduce through the C ABI
                   (engine.EngineGroup = the single-process form, csim_group_run)
"""
from ._cabi import (CsimError, WIRING_REF_CLI, WIRING_MODEL, ENGINE_AUTO, ENGINE_DENSE, ENGINE_TABLE, ENGINE_PREFIX, ENGINE_FULL,
                    FP32, FP64, REASON_MAX_ROUNDS, REASON_WINNER, REASON_NO_PROBS)
from .engine import Engine, EngineGroup, get_engine, get_group, device_count, make_params, RunResult

__all__ = ["Engine", "EngineGroup", "get_engine", "get_group", "device_count", "make_params", "RunResult", "CsimError",
           "WIRING_REF_CLI", "WIRING_MODEL", "ENGINE_AUTO", "ENGINE_DENSE", "ENGINE_TABLE", "ENGINE_PREFIX", "ENGINE_FULL", "FP32", "FP64",
           "REASON_MAX_ROUNDS", "REASON_WINNER", "REASON_NO_PROBS"]

"""Multi-GPU, one process per GPU (torchrun): trials shard by disjoint sim-id ranges; the only exchange is
ONE all-reduce (sum, int64) of the concatenated winner / rounds histograms and totals
(replaces the Counter / mean aggregation of src/simulate.py:356-387 across processes).

Because a trial's Philox stream is keyed by its global sim_id, results are independent of the
number of ranks.  On GPUs the all-reduce is issued through the C ABI (csim_allreduce_hists, NCCL bound by
the library at run time) on a communicator whose 128-byte id travels through torch.distributed — which is
plumbing only here (rendezvous, the id broadcast, gloo in the CPU tests).  The single-process form
(`--devices N`, csim_group_run) lives in engine.EngineGroup.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np


def shard_range(n_sims: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, begin+count) of rank's share of n_sims."""
    base, rem = divmod(int(n_sims), int(world))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def pack_hists(winner_hist: np.ndarray, rounds_hist: np.ndarray, totals: np.ndarray) -> np.ndarray:
    return np.concatenate([winner_hist.reshape(-1), rounds_hist.reshape(-1), totals.reshape(-1)]).astype(np.int64)


def unpack_hists(flat: np.ndarray, n_sets: int, E: int, R: int):
    a = n_sets * (E + 1)
    b = a + n_sets * (R + 1)
    return flat[:a].reshape(n_sets, E + 1), flat[a:b].reshape(n_sets, R + 1), flat[b:].reshape(n_sets, 2)


def allreduce_sum_(tensor) -> None:
    """In-place sum over ranks when a process group exists; no-op otherwise."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)


def run_sharded(run_shard: Callable[[int, int], Dict[str, np.ndarray]], n_sims: int, n_sets: int, E: int, R: int,
                rank: int, world: int, device: Optional[str] = None) -> Dict[str, np.ndarray]:
    """Host-side driver: `run_shard(begin, count)` returns this rank's histograms (numpy);
    they are summed over ranks with one all-reduce.  Used by the CPU (gloo) tests and by the CLI."""
    import torch
    begin, count = shard_range(n_sims, rank, world)
    local = run_shard(begin, count)
    flat = torch.from_numpy(pack_hists(local["winner_hist"], local["rounds_hist"], local["totals"]))
    if device is not None:
        flat = flat.to(device)
    allreduce_sum_(flat)
    wh, rh, tot = unpack_hists(flat.cpu().numpy(), n_sets, E, R)
    return dict(winner_hist=wh, rounds_hist=rh, totals=tot, begin=begin, count=count, local=local)


_comms: Dict[int, int] = {}


def csim_comm(engine) -> Optional[int]:
    """This rank's ncclComm_t for `engine`'s device (created once per process through the C ABI: rank 0 makes the NCCL id,
    torch.distributed broadcasts its 128 bytes).  None outside a multi-rank job."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    if engine.device in _comms:
        return _comms[engine.device]
    from .engine import nccl_unique_id
    box = [nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = engine.nccl_comm_create(box[0], dist.get_world_size(), dist.get_rank())
    _comms[engine.device] = comm
    return comm


class DeviceBatch:
    """Device-resident buffers of one rank's shard + the stream-ordered run / all-reduce pair
    (kernel, then ONE NCCL all-reduce of the packed histogram tensor on the same stream: csim_allreduce_hists on `comm`
    when one is given, else torch.distributed's)."""

    def __init__(self, engine, n_sets: int, n_sims_local: int, R: int, want_reason: bool = False, comm: Optional[int] = None):
        import torch
        self.torch = torch
        self.engine = engine
        self.comm = comm
        E = engine.n_electors
        dev = torch.device("cuda", engine.device)
        n = n_sets * n_sims_local
        self.n_sets, self.n_sims_local, self.R, self.E = n_sets, n_sims_local, R, E
        self.winner = torch.empty(n, dtype=torch.int32, device=dev)
        self.rounds = torch.empty(n, dtype=torch.int32, device=dev)
        self.reason = torch.empty(n, dtype=torch.uint8, device=dev) if want_reason else None
        self.hist = torch.zeros(n_sets * (E + 1) + n_sets * (R + 1) + n_sets * 2, dtype=torch.int64, device=dev)
        a = n_sets * (E + 1)
        b = a + n_sets * (R + 1)
        self.winner_hist, self.rounds_hist, self.totals = self.hist[:a], self.hist[a:b], self.hist[b:]

    def run(self, param_sets, *, sim_id_begin: int, seed: int, wiring: int, engine: int, precision: int, reduce: bool = True,
            job_sims: int = 0):
        """job_sims: trials per set of the whole job over all ranks (so that CSIM_ENGINE_AUTO chooses the same engine on every
        rank as a single-device run of the job would); 0 = this shard's own size."""
        torch = self.torch
        self.hist.zero_()
        stream = torch.cuda.current_stream(self.winner.device).cuda_stream
        if self.n_sims_local > 0:          # a rank whose shard is empty (fewer trials than ranks) still joins the all-reduce
            self.engine.run_device(param_sets, self.n_sims_local, winner=self.winner, rounds=self.rounds, reason=self.reason,
                                   winner_hist=self.winner_hist, rounds_hist=self.rounds_hist, totals=self.totals,
                                   sim_id_begin=sim_id_begin, seed=seed, wiring=wiring, engine=engine, precision=precision,
                                   stream=stream, auto_batch_sims=job_sims)
        if reduce:
            if self.comm is not None:
                self.engine.allreduce_hists(self.comm, self.winner_hist, self.rounds_hist, self.totals, self.n_sets, self.R, stream)
            else:
                allreduce_sum_(self.hist)

    def hists(self):
        flat = self.hist.cpu().numpy()
        return unpack_hists(flat, self.n_sets, self.E, self.R)

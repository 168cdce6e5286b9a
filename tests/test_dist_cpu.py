"""world_size-2 gloo test of the multi-GPU host logic (conclave_sim_b200/dist.py): disjoint
sim-id shards + ONE all-reduce of the packed histograms == the single-process result.  The
per-shard engine here is the CPU oracle (no GPU in this container); on the GPU box the same
driver code is exercised with the CUDA engine by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest

from oracle import conclave_oracle as O
import golden_util as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sims, q):
    import torch.distributed as dist
    from conclave_sim_b200 import dist as D
    from oracle import c_oracle as CO
    from oracle.gen_golden import CFG2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    roster = G.real_roster()
    sets = [(O.ModelParams(**CFG2), O.EngineParams(max_rounds=20, supermajority_threshold=0.56))]

    def run_shard(begin, count):
        out = CO.run(roster, sets, O.WIRING_REF_CLI, count, seed=77, sim_id_begin=begin, threads=1)
        return out

    res = D.run_sharded(run_shard, n_sims, 1, roster.E, 20, rank, world)
    q.put((rank, res["begin"], res["count"], res["winner_hist"], res["rounds_hist"], res["totals"],
           res["local"]["winner"], res["local"]["rounds"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    from conclave_sim_b200.dist import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gloo_allreduce_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import c_oracle as CO
    from oracle.gen_golden import CFG2
    CO.build()
    n_sims, world, port = 301, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sims, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    roster = G.real_roster()
    ref = CO.run(roster, [(O.ModelParams(**CFG2), O.EngineParams(max_rounds=20, supermajority_threshold=0.56))],
                 O.WIRING_REF_CLI, n_sims, seed=77)
    for rank, begin, count, wh, rh, tot, w_local, r_local in got:
        assert np.array_equal(wh, ref["winner_hist"]) and np.array_equal(rh, ref["rounds_hist"]) and np.array_equal(tot, ref["totals"])
        assert np.array_equal(w_local, ref["winner"][begin:begin + count]) and np.array_equal(r_local, ref["rounds"][begin:begin + count])
    assert got[0][2] + got[1][2] == n_sims


def test_pack_unpack_roundtrip():
    from conclave_sim_b200.dist import pack_hists, unpack_hists
    rng = np.random.default_rng(0)
    wh = rng.integers(0, 100, (3, 135)); rh = rng.integers(0, 100, (3, 21)); tot = rng.integers(0, 100, (3, 2))
    a, b, c = unpack_hists(pack_hists(wh, rh, tot), 3, 134, 20)
    assert np.array_equal(a, wh) and np.array_equal(b, rh) and np.array_equal(c, tot)

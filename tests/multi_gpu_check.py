#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs; not a pytest module: the CPU suite covers the sharding
logic with gloo, this script proves the same on real devices over NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py

For every engine: each rank simulates its contiguous share of the sim-id range (dist.shard_range) on its own GPU, the
winner / rounds histograms and totals are combined by the ONE NCCL all-reduce of the data path (dist.DeviceBatch), and
rank 0 recomputes the WHOLE range on its GPU alone: the reduced histograms must equal the single-GPU ones exactly, and the
per-trial rows gathered from the ranks must equal the single-GPU rows — a trial is a pure function of (seed, sim_id).
Round 2 additions: the all-reduce goes through the C ABI (csim_allreduce_hists on a communicator made by csim_nccl_comm_create;
torch.distributed only carries the 128-byte NCCL id), and the same split is checked through the PUBLIC surface:
`simulate.run_simulations(...)` inside the torchrun job (rank 0 receives every rank's rows + the reduced histograms) and
`simulate.main([...])` (rank 0 writes the CSV / JSON) must equal a single-device run of the same call.
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from bench import synthetic_roster, CFG2, ENGINE_KW
    from conclave_sim_b200 import (Engine, make_params, FP32, FP64, WIRING_REF_CLI, WIRING_MODEL, ENGINE_DENSE, ENGINE_TABLE,
                                   ENGINE_PREFIX, ENGINE_FULL)
    from conclave_sim_b200.dist import DeviceBatch, shard_range, csim_comm
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E, n, R = 135, 60_001, ENGINE_KW["max_rounds"]            # an odd total: ragged shards
    x, g, p = synthetic_roster(E)
    eng = Engine(local)
    eng.upload_roster(x, g, p)
    params = make_params(CFG2, **ENGINE_KW)
    params_x = make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), enable_candidate_fatigue=True,
                           enable_stop_candidate=True, **ENGINE_KW)
    begin, count = shard_range(n, rank, world)
    cases = [("dense", ENGINE_DENSE, WIRING_REF_CLI, FP32, params), ("table", ENGINE_TABLE, WIRING_REF_CLI, FP32, params),
             ("table64", ENGINE_TABLE, WIRING_REF_CLI, FP64, params), ("prefix", ENGINE_PREFIX, WIRING_REF_CLI, FP32, params),
             ("prefix/model", ENGINE_PREFIX, WIRING_MODEL, FP32, params), ("full/model+extras", ENGINE_FULL, WIRING_MODEL, FP32, params_x)]
    report, ok = {}, True
    comm = csim_comm(eng)                                  # this rank's ncclComm_t, created through the C ABI
    for name, engine, wiring, prec, prm in cases:
        batch = DeviceBatch(eng, 1, count, R, want_reason=False, comm=comm)
        batch.run([prm], sim_id_begin=begin, seed=20250507, wiring=wiring, engine=engine, precision=prec, reduce=True)
        torch.cuda.synchronize()
        hist = batch.hist.cpu().numpy().copy()
        rows = torch.stack([batch.winner, batch.rounds]).contiguous()
        gathered = [torch.empty((2, shard_range(n, r, world)[1]), dtype=rows.dtype, device=rows.device) for r in range(world)]
        dist.all_gather(gathered, rows)
        if rank == 0:
            whole = eng.run([prm], n, seed=20250507, wiring=wiring, engine=engine, precision=prec, want_reason=False)
            ref_hist = np.concatenate([whole.winner_hist.ravel(), whole.rounds_hist.ravel(), whole.totals.ravel()])
            w = torch.cat([t[0] for t in gathered]).cpu().numpy()
            rd = torch.cat([t[1] for t in gathered]).cpu().numpy()
            good = bool(np.array_equal(hist, ref_hist) and np.array_equal(w, whole.winner) and np.array_equal(rd, whole.rounds))
            report[name] = {"ok": good, "trials": n, "winners": int((whole.winner >= 0).sum()), "sum_rounds": int(whole.totals[0, 0])}
            ok = ok and good
    # ---- the public surface inside the torchrun job -----------------------------------------------------------
    import tempfile
    import pandas as pd
    from conclave_sim_b200 import simulate as S
    from conclave_sim_b200.engine import get_engine
    df = pd.DataFrame({"ideology_score": x, "region": g, "is_papabile": p}, index=pd.Index([str(i) for i in range(E)], name="elector_id"))
    api_kw = dict(max_rounds=R, supermajority_threshold=ENGINE_KW["supermajority_threshold"], runoff_threshold_rounds=ENGINE_KW["runoff_threshold_rounds"],
                  runoff_candidate_count=ENGINE_KW["runoff_candidate_count"], seed=77)
    S.RUNTIME.update(device=local)
    for wiring in ("ref_cli", "model"):
        got = S.run_simulations(df, CFG2, n, wiring=wiring, **api_kw)                       # sharded over the ranks
        if rank == 0:
            one = S.run_simulations(df, CFG2, n, wiring=wiring, distributed=False, device=local, **api_kw)   # this GPU alone
            good = bool(np.array_equal(got.winner, one.winner) and np.array_equal(got.rounds, one.rounds) and np.array_equal(got.reason, one.reason)
                        and np.array_equal(got.winner_hist, one.winner_hist) and np.array_equal(got.rounds_hist, one.rounds_hist)
                        and np.array_equal(got.totals, one.totals))
            report[f"run_simulations/{wiring}"] = {"ok": good, "engine": got.engine}
            ok = ok and good
    # main(): rank 0 writes the files; compare with a single-device main() of the same command line
    tmp = tempfile.mkdtemp() if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    if rank == 0:
        df.reset_index().to_csv(os.path.join(tmp, "roster.csv"), index=False)
    dist.barrier()
    argv = ["-n", str(n), "-f", os.path.join(tmp, "roster.csv"), "--max-rounds", "20", "--initial-beta-weight", "1.5", "--enable-dynamic-beta",
            "--beta-increment-amount", "0.3", "--beta-increment-interval-rounds", "1", "--bandwagon-strength", "0.7", "--papabile-weight-factor", "2.0",
            "--stickiness-factor", "0.8", "-t", "0.56", "--seed", "5"]
    S.main(argv + ["-o", os.path.join(tmp, "multi.csv"), "-s", os.path.join(tmp, "multi.json")])
    dist.barrier()
    if rank == 0:
        S.main(argv + ["-o", os.path.join(tmp, "one.csv"), "-s", os.path.join(tmp, "one.json"), "--devices", "1", "--device", str(local)])
        good = all(open(os.path.join(tmp, f"multi.{e}"), "rb").read() == open(os.path.join(tmp, f"one.{e}"), "rb").read() for e in ("csv", "json"))
        report["main()/csv+json bytes"] = {"ok": bool(good), "rows": n}
        ok = ok and good
    if rank == 0:
        print(json.dumps({"multi_gpu_check": "pass" if ok else "FAIL", "n_gpus": world,
                          "collective": "1 NCCL all-reduce (int64 sum) per engine run, issued through csim_allreduce_hists (C ABI)",
                          "engines": report}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

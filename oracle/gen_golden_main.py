"""Golden files of the reference's own main() — aggregation, CSV and stats JSON as the UNMODIFIED reference writes them
(src/simulate.py:297-434).  TEST INFRASTRUCTURE ONLY; runs in the build container where /root/reference is mounted.

The reference's main() is executed as it lies.  Patched from the outside only: `ProcessPoolExecutor` / `as_completed` in its
namespace are replaced by a synchronous executor that seeds numpy's global MT19937 with seed0 + sim_id before each (real,
unmodified) worker call, so that (a) every trial's uniforms are reproducible — np.random.seed(seed0 + i).random_sample(R * E)
is the tape `python -m src.simulate --tape` replays on the GPU — and (b) results arrive in sim-id order (the reference's
completion order is otherwise unspecified).  Outputs: tests/golden/main/<case>.csv, <case>.json, and main_cases.json.

    python -m oracle.gen_golden_main
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from oracle import ref_harness as H

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "main")

CFG2_ARGS = ["--max-rounds", "20", "--initial-beta-weight", "1.5", "--enable-dynamic-beta", "--beta-increment-amount", "0.3",
             "--beta-increment-interval-rounds", "1", "--bandwagon-strength", "0.7", "--papabile-weight-factor", "2.0",
             "--stickiness-factor", "0.8", "-t", "0.56"]
CASES = [
    dict(name="cfg2_n48", n=48, seed0=424200, args=CFG2_ARGS),                                   # winners and no-winner rows mixed
    dict(name="cfg1_n12", n=12, seed0=515100, args=["--max-rounds", "20"]),                      # CLI defaults: nobody wins (None stats)
    dict(name="cfg2_thr50_n24", n=24, seed0=616100, args=CFG2_ARGS[:-2] + ["-t", "0.5", "--runoff-rounds", "3"]),
    dict(name="n0", n=0, seed0=1, args=["--max-rounds", "20"]),                                  # empty run
]


class _Future:
    def __init__(self, value=None, exc=None):
        self._v, self._e = value, exc

    def result(self):
        if self._e is not None:
            raise self._e
        return self._v


class _SyncExecutor:
    seed0 = 0

    def __init__(self, max_workers=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def submit(self, fn, sim_id, *args):
        np.random.seed(_SyncExecutor.seed0 + int(sim_id))
        try:
            return _Future(fn(sim_id, *args))
        except Exception as exc:  # pragma: no cover
            return _Future(exc=exc)


def main():
    os.makedirs(GOLD, exist_ok=True)
    H.quiet()
    _, ref_sim = H.load_reference()
    ref_sim.ProcessPoolExecutor = _SyncExecutor
    ref_sim.as_completed = lambda futures: list(futures)
    roster = os.path.join(H.REFERENCE_ROOT, "merged_electors.csv")
    meta = []
    for c in CASES:
        csv, js = os.path.join(GOLD, c["name"] + ".csv"), os.path.join(GOLD, c["name"] + ".json")
        for p in (csv, js):
            if os.path.exists(p):
                os.remove(p)
        _SyncExecutor.seed0 = c["seed0"]
        argv = ["simulate", "-n", str(c["n"]), "-f", roster, "-o", csv, "-s", js] + c["args"]
        old = sys.argv
        sys.argv = argv
        try:
            ref_sim.main()
        finally:
            sys.argv = old
        meta.append(dict(name=c["name"], n=c["n"], seed0=c["seed0"], args=c["args"], csv_written=os.path.exists(csv),
                         json_written=os.path.exists(js)))
        print(c["name"], "csv", os.path.getsize(csv) if os.path.exists(csv) else None, "json", os.path.getsize(js) if os.path.exists(js) else None, flush=True)
    with open(os.path.join(GOLD, "main_cases.json"), "w") as f:
        json.dump(dict(roster="merged_electors.csv of the reference (tests/golden/roster_real.npz holds the three columns + ids)",
                       seeding="np.random.seed(seed0 + sim_id) before each unmodified worker call; results in sim-id order",
                       cases=meta), f, indent=1)


if __name__ == "__main__":
    main()

"""Tuning aid: times the FULL engine's EXT path (fatigue / stop-candidate live) at 250 k trials; CSIM_LIB / CSIM_EXT_WPC select the variant."""
import json, os, sys
sys.path.insert(0, '.')
from bench import synthetic_roster, CFG2, ENGINE_KW
from conclave_sim_b200 import Engine, make_params, FP32, WIRING_MODEL, ENGINE_FULL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
eng = Engine(0); x, g, p = synthetic_roster(135); eng.upload_roster(x, g, p)
for name, kw in (("fat+stop", dict(enable_candidate_fatigue=True, enable_stop_candidate=True)), ("fat", dict(enable_candidate_fatigue=True)),
                 ("stop", dict(enable_stop_candidate=True))):
    pr = make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), **kw, **ENGINE_KW)
    best = None
    for rep in range(3):
        r = eng.run([pr], n, seed=1, wiring=WIRING_MODEL, engine=ENGINE_FULL, precision=FP32, want_reason=False)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("CSIM_LIB", "default")), wpc=os.environ.get("CSIM_EXT_WPC", "16"), legacy=bool(os.environ.get("CSIM_FULL_LEGACY")),
                          what=name, trials=n, kernel_ms=best, er_per_s=r.elector_rounds / (best * 1e-3))), flush=True)

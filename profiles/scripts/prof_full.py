import sys, numpy as np
sys.path.insert(0,'.')
from bench import synthetic_roster, CFG2, ENGINE_KW
from conclave_sim_b200 import Engine, make_params, FP32, WIRING_MODEL, ENGINE_FULL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng=Engine(0); x,g,p=synthetic_roster(135); eng.upload_roster(x,g,p)
pr=make_params(dict(CFG2, stop_candidate_threshold_unacceptable_distance=0.3), enable_candidate_fatigue=True, enable_stop_candidate=True, **ENGINE_KW)
for rep in range(2):
    r=eng.run([pr], n, seed=1, wiring=WIRING_MODEL, engine=ENGINE_FULL, precision=FP32, want_reason=False)
print("kernel_ms", r.kernel_ms, "er/s", r.elector_rounds/(r.kernel_ms*1e-3), "mean rounds", r.rounds.mean())

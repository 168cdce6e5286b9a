import sys, numpy as np
sys.path.insert(0,'.')
from bench import synthetic_roster, CFG2, ENGINE_KW
from conclave_sim_b200 import Engine, make_params, FP32, WIRING_REF_CLI, WIRING_MODEL, ENGINE_PREFIX
wiring = WIRING_MODEL if (len(sys.argv) > 1 and sys.argv[1] == "model") else WIRING_REF_CLI
E = int(sys.argv[2]) if len(sys.argv) > 2 else 135
n = int(sys.argv[3]) if len(sys.argv) > 3 else 250000
R = 50 if E > 256 else 20
eng=Engine(0); x,g,p=synthetic_roster(E); eng.upload_roster(x,g,p)
kw=dict(ENGINE_KW); kw["max_rounds"]=R
pr=make_params(CFG2, **kw)
for rep in range(2):
    r=eng.run([pr], n, seed=1, wiring=wiring, engine=ENGINE_PREFIX, precision=FP32, want_reason=False)
print("kernel_ms", r.kernel_ms, "er/s", r.elector_rounds/(r.kernel_ms*1e-3))

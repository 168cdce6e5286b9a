"""CPU tests of the FULL engine's host-side launch plan (csim_full_plan_query): pure arithmetic behind the C ABI, no device needed.
Which of its two kernels runs a batch with candidate fatigue / stop-candidate live (the EXT instantiation of the dense kernel,
dense32.cuh, or the general k_trials_full32), with how many warps and how much shared memory — for any roster shape."""
import ctypes as C
import itertools

import pytest

from conclave_sim_b200 import _cabi

BUDGET = 220 * 1024      # what the fast path plans against
OPTIN = 232448           # opt-in shared memory per block on sm_100: what k_trials_full32 plans against


def plan(E, G=8, R=20, full_rounds=5, kmax=2, F=3, eligible=1):
    lib = _cabi.load_library()
    p = _cabi.CsimFullPlan()
    st = lib.csim_full_plan_query(E, min(G, max(E, 1)), R, full_rounds, kmax, F, eligible, C.byref(p))
    return st, p


def test_bench_shape_runs_the_fast_path_at_20_warps():
    """bench.py's FULL step: 135 electors, 8 regions, 20 rounds of which 5 can be full-field, k = 2, fatigue window 3."""
    st, p = plan(135)
    assert st == 0 and p.fast_path == 1 and p.general_fits == 1
    assert p.warps_per_cta == 20 and p.cta_tabs == 1 and p.n_chunks == 17 and p.acc_pitch == 136
    assert p.smem_total == p.smem_cta + 20 * p.smem_per_warp <= BUDGET
    assert p.smem_cta + 21 * p.smem_per_warp > BUDGET or p.warps_per_cta == 20          # 20 = the register-bounded CTA (640 threads)


def test_round_tables_go_per_warp_when_sharing_them_would_cost_a_warp():
    """110 rounds without a run-off: 110 whole tables (150 KB) do not fit beside 20 warps; every warp stages the round it plays."""
    st, p = plan(135, R=110, full_rounds=110, kmax=1)
    assert st == 0 and p.fast_path == 1 and p.cta_tabs == 0 and p.warps_per_cta == 20
    st, q = plan(135, R=110, full_rounds=5)             # a run-off from round 6 on: 5 whole + 105 one-chunk tables would cost one warp
    assert st == 0 and q.fast_path == 1 and q.cta_tabs == 0 and q.warps_per_cta == 20
    st, q = plan(135, R=30, full_rounds=5)              # 5 whole + 25 one-chunk tables fit beside 20 warps
    assert st == 0 and q.fast_path == 1 and q.cta_tabs == 1 and q.warps_per_cta == 20


def test_what_the_fast_path_does_not_take():
    for kw in (dict(E=300), dict(E=256), dict(E=1), dict(E=135, eligible=0)):
        st, p = plan(**kw)
        assert st == 0 and p.fast_path == 0 and p.general_fits == 1 and p.warps_per_cta >= 1, kw
    st, p = plan(135, G=80)                             # bonus rows of 80 regions: fewer than 10 warps fit, the general kernel's 28 run
    assert st == 0 and p.fast_path == 0 and p.general_fits == 1
    st, p = plan(255, G=255, F=3)                       # neither kernel holds 255 regions x 255 candidates of bonus rows
    assert st == -4 or p.fast_path in (0, 1)


@pytest.mark.parametrize("E", [2, 3, 7, 8, 9, 33, 64, 127, 128, 129, 134, 135, 136, 192, 200, 254, 255])
def test_plan_invariants(E):
    for G, R, kmax, F in itertools.product((1, 4, 8, 16, 40, 80), (1, 20, 100), (1, 2, 7), (1, 3, 6)):
        if G > E:
            continue
        for full_rounds in {1, min(5, R), R}:
            st, p = plan(E, G, R, full_rounds, kmax, F)
            if st != 0:
                assert st == -4                          # CSIM_ERR_UNSUPPORTED: neither kernel fits (many regions)
                assert G >= 40
                continue
            assert p.smem_total == p.smem_cta + p.smem_per_warp * p.warps_per_cta <= (BUDGET if p.fast_path else OPTIN)
            assert p.n_chunks >= 1
            if p.fast_path:
                assert p.n_chunks == (E + 7) // 8 and p.acc_pitch >= E and p.acc_pitch % 2 == 0
                assert 2 <= p.warps_per_cta <= 20
                assert p.warps_per_cta >= 10 or not p.general_fits
                assert p.smem_cta % 16 == 0 and p.smem_per_warp % 16 == 0
            else:
                assert p.general_fits and 1 <= p.warps_per_cta <= 28
            # the plan never depends on eligibility except through the path
            st2, q = plan(E, G, R, full_rounds, kmax, F, eligible=0)
            assert (st2 == 0) == bool(p.general_fits) and (st2 != 0 or q.fast_path == 0)


def test_argument_errors():
    lib = _cabi.load_library()
    p = _cabi.CsimFullPlan()
    assert lib.csim_full_plan_query(0, 1, 20, 5, 2, 3, 1, C.byref(p)) == -1
    assert lib.csim_full_plan_query(10, 11, 20, 5, 2, 3, 1, C.byref(p)) == -1
    assert lib.csim_full_plan_query(10, 2, 20, 5, 2, 3, 1, None) == -1

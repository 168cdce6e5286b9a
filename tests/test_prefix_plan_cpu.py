"""CPU tests of the PREFIX engine's host-side launch plan (csim_prefix_plan_query / csim_prefix_slot): pure arithmetic
behind the C ABI, no device needed.  The kernel trusts these numbers blindly (it reads the offsets from the constant
bank), so every invariant it relies on is pinned here for a wide range of roster shapes."""
import ctypes as C
import itertools

import pytest

from conclave_sim_b200 import _cabi

BUDGET = 232448          # opt-in shared memory per block on sm_100 (cudaDevAttrMaxSharedMemoryPerBlockOptin)


def plan(E, G=8, R=20, Rt=5, kmax=2, wiring=1, budget=BUDGET, allow_stage=1):
    lib = _cabi.load_library()
    p = _cabi.CsimPrefixPlan()
    st = lib.csim_prefix_plan_query(E, min(G, max(E, 1)), R, Rt, kmax, wiring, budget, allow_stage, C.byref(p))
    return st, p


@pytest.mark.parametrize("E", [1, 2, 3, 7, 8, 9, 33, 64, 127, 128, 129, 134, 135, 136, 192, 193, 200, 255, 256, 257, 300, 511, 512, 513,
                               1000, 1024, 1025, 2048, 5000, 32767])
def test_plan_invariants(E):
    for G, R, Rt, kmax, wiring, stage in itertools.product((1, 8, 64, 200), (0, 1, 20, 50), (1, 5, 50), (1, 2, 7), (0, 1), (0, 1)):
        if G > E:
            continue
        st, p = plan(E, G, R, Rt, kmax, wiring, allow_stage=stage)
        if st != 0:
            # the only legal failure: one warp of this roster does not fit next to the roster tables
            assert st == -4
            assert E > 2000 or G >= 64
            continue
        L, nch = p.chunk, p.n_chunks
        assert L >= 8 and L & (L - 1) == 0 and 1 <= nch <= 128 and nch * L >= E and (nch - 1) * L < E
        assert p.row_pitch % 2 == 1 and p.row_pitch >= nch
        assert p.rows >= E and p.e_quad_end % 128 == 0 and p.rows == max(E, p.e_quad_end)
        assert (E - p.e_quad_end) <= 64                                    # the tail dealt one elector per lane is at most two passes
        assert 1 <= p.table_rounds <= max(R, 1)
        assert 1 <= p.warps_per_cta <= 32
        assert p.smem_total == p.smem_cta + p.smem_per_warp * p.warps_per_cta <= BUDGET
        if not stage:
            assert not p.staged
        assert p.smem_tables == (p.table_bytes_per_set if p.staged else 0)
        assert p.table_bytes_per_set >= 4 * p.table_rounds * p.rows * p.row_pitch and p.table_bytes_per_set % 16 == 0
        if p.staged:
            assert L <= 16 and p.warps_per_cta >= 8
        o_xs, o_pw, o_rb, o_xg, o_nbs, w_bwp, w_tally, w_eff, w_votes, w_mark, cp4, step0_4 = list(p.offsets)
        assert cp4 == 4 * nch * (L + 4) and cp4 % 16 == 0
        # CTA arrays: 16-byte aligned (128-bit loads), in order, no overlap, inside the CTA part
        assert o_xs == p.smem_tables and o_pw == o_xs + cp4 and o_rb == o_pw + cp4 and o_xg == o_rb + G * cp4
        assert o_nbs >= o_xg + 8 * p.rows and o_nbs + 4 * R <= p.smem_cta
        assert all(o % 16 == 0 for o in (o_xs, o_pw, o_rb, o_xg, o_nbs, p.smem_cta, p.smem_per_warp))
        # warp arrays
        Ei = (E + 7) & ~7
        if wiring == 1:
            assert w_bwp == cp4 and w_tally == w_bwp + 4 * 128 and w_votes >= w_eff + 4 * kmax and w_mark >= w_votes + 2 * p.rows
        else:
            assert w_tally == 0 and w_mark == w_votes
        assert w_eff == w_tally + 4 * Ei and w_mark + E <= p.smem_per_warp
        assert w_tally % 16 == 0 and w_bwp % 16 == 0
        assert step0_4 % 4 == 0 and step0_4 // 4 <= nch < 2 * (step0_4 // 4)


@pytest.mark.parametrize("E", [1, 5, 64, 65, 127, 128, 134, 135, 192, 193, 255, 256, 300, 1024, 1100])
def test_slot_is_a_bijection_with_conflict_free_passes(E):
    lib = _cabi.load_library()
    st, p = plan(E)
    assert st == 0
    slots = [lib.csim_prefix_slot(e, p.e_quad_end) for e in range(E)]
    assert len(set(slots)) == E and min(slots) >= 0 and max(slots) < p.rows
    # one pass of a warp (fixed quad group qi and word i) touches 32 CONSECUTIVE rows: lane l <-> elector 4 (l + 32 qi) + i
    for e in range(min(E, p.e_quad_end)):
        q, i = e >> 2, e & 3
        assert slots[e] == (q >> 5) * 128 + i * 32 + (q & 31)
    for e in range(p.e_quad_end, E):
        assert slots[e] == e


def test_cfg_shapes():
    """The shapes DESIGN.md quotes."""
    st, p = plan(134, 8, 20, 5, 2, 1)
    assert st == 0 and p.staged and p.chunk == 8 and p.n_chunks == 17 and p.warps_per_cta == 32 and p.smem_tables == 45568
    st, p = plan(134, 8, 20, 20, 1, 1)          # no run-off: all 20 rounds can be full-field, 182 KB of tables still fit
    assert st == 0 and p.staged and p.smem_tables == 182240 and p.warps_per_cta >= 16
    st, p = plan(1024, 8, 50, 5, 2, 0)          # cfg-5: tables through L2
    assert st == 0 and not p.staged and p.chunk == 8 and p.n_chunks == 128 and p.table_bytes_per_set == 5 * 1024 * 129 * 4
    st, p = plan(134, 8, 20, 5, 2, 1, budget=40000)
    assert st == 0 and not p.staged             # a small budget falls back to L2 tables
    st, p = plan(134, 8, 20, 5, 2, 1, budget=4000)
    assert st == -4                             # CSIM_ERR_UNSUPPORTED, reported, never a silent fallback
    st, p = plan(0)
    assert st == -1

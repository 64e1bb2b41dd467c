"""CPU tests of the persistent GEMM kernel's static tile schedule (hyp_tc_engine.cuh: assign_units), through the
host-only debug entry point hyp_debug_schedule: every unit is scheduled exactly once, groups keep the window order,
and the predicted load spread stays small on the launch shapes of the C2 step."""
import ctypes

import numpy
import pytest

from hypelcnn_b200 import _native as N


@pytest.fixture(scope="module")
def lib():
    N.build_native()
    return N.lib()


def schedule(lib, costs, groups, windowed):
    costs = numpy.ascontiguousarray(costs, dtype=numpy.float64)
    g = numpy.empty(len(costs), dtype=numpy.int32)
    r = numpy.empty(len(costs), dtype=numpy.int32)
    N.check(lib.hyp_debug_schedule(costs.ctypes.data_as(ctypes.c_void_p), len(costs), groups, int(windowed),
                                   g.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p)))
    return g, r


def level_costs(batch_pairs=16, P=7, R=4, cin_blocks=8, fpad=32):
    """tile_cost of a level forward launch: per output position the taps inside the patch, a tap of ring r costing
    K blocks x (256 + N(r)); tiles emitted batch-pair major, heaviest position first (as tc_plan does)."""
    per_pos = []
    for y in range(P):
        for x in range(P):
            c = 1500.0
            for dy in range(-(R - 1), R):
                for dx in range(-(R - 1), R):
                    if 0 <= y + dy < P and 0 <= x + dx < P:
                        c += cin_blocks * (256.0 + (R - max(abs(dy), abs(dx))) * fpad)
            per_pos.append(c)
    per_pos.sort(reverse=True)
    return numpy.array(per_pos * batch_pairs)


def spread(costs, g, groups):
    load = numpy.bincount(g, weights=costs, minlength=groups)
    return load.max() / load.mean()


@pytest.mark.parametrize("windowed", [True, False])
def test_every_unit_is_scheduled_once_and_ranks_are_dense(lib, windowed):
    rng = numpy.random.default_rng(3)
    for units, groups in ((784, 74), (1, 74), (73, 74), (148, 74), (149, 74), (3108, 148), (5, 2), (0, 3)):
        costs = rng.uniform(1.0, 3.0, units)
        g, r = schedule(lib, costs, groups, windowed)
        assert g.min(initial=0) >= 0 and g.max(initial=0) < groups
        for grp in range(groups):
            ranks = numpy.sort(r[g == grp])
            assert numpy.array_equal(ranks, numpy.arange(len(ranks)))       # dense 0..n-1: no holes, no duplicates


def test_windowed_schedule_keeps_window_order_and_caps_units_per_window(lib):
    groups = 74
    costs = level_costs()
    g, r = schedule(lib, costs, groups, True)
    window = numpy.arange(len(costs)) // (2 * groups)
    for grp in range(groups):
        mine = numpy.flatnonzero(g == grp)
        order = mine[numpy.argsort(r[mine])]
        assert numpy.all(numpy.diff(window[order]) >= 0)                     # executed in the original window order
        assert numpy.bincount(window[mine]).max() <= 2                       # ceil(2G / G) units per window and group


def test_load_spread_on_the_c2_launch_shapes(lib):
    costs = level_costs()                                                    # connector_1 forward: 784 pair units
    g, _ = schedule(lib, costs, 74, True)
    assert spread(costs, g, 74) < 1.04                                       # was 1.20 with the per-window snake
    # round-robin over the same order, for comparison: what the kernel would do without a schedule
    rr = numpy.arange(len(costs)) % 74
    assert spread(costs, g, 74) < spread(costs, rr, 74)
    # LPT on a handful of big units: within the 4/3 bound of the optimum (= at least the mean and the largest unit)
    rng = numpy.random.default_rng(1)
    big = rng.uniform(50.0, 100.0, 196)
    g, _ = schedule(lib, big, 148, False)
    load = numpy.bincount(g, weights=big, minlength=148)
    assert load.max() <= 4.0 / 3.0 * max(load.mean(), big.max()) + 1e-9


def test_bad_arguments_are_rejected(lib):
    assert lib.hyp_debug_schedule(None, 3, 2, 1, None, None) == N.HYP_E_INVALID


def tap_group_rows(lib, P, R, nt, fpad, max_sets):
    n = ctypes.c_int(0)
    N.check(lib.hyp_debug_level_tap_groups(P, R, nt, fpad, max_sets, None, 0, ctypes.byref(n)))
    out = numpy.empty((n.value, 6), dtype=numpy.int32)
    N.check(lib.hyp_debug_level_tap_groups(P, R, nt, fpad, max_sets, out.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
    return out


@pytest.mark.parametrize("P,R,nt,fpad,max_sets", [(7, 4, 1, 32, 2), (7, 4, 1, 16, 2), (7, 4, 1, 16, 4), (7, 4, 1, 64, 2),
                                                  (11, 6, 1, 32, 2), (3, 3, 1, 16, 4), (7, 4, 1, 32, 1), (5, 2, 2, 128, 2)])
def test_level_wgrad_tap_groups_cover_every_tap_and_position_once(lib, P, R, nt, fpad, max_sets):
    """Every tap of the level belongs to exactly one (group, set); over the source positions a group visits, a tap reads
    exactly the output positions p with p + tap inside the patch (each once), everything else is the out-of-bounds
    marker; no visited position is all out of bounds; the group's accumulator fits the kernel's limits."""
    rows = tap_group_rows(lib, P, R, nt, fpad, max_sets)
    h = min(R - 1, P - 1)
    taps = {}
    for g, j, dy, dx, q, pos in rows.tolist():
        taps.setdefault((dy, dx), set()).add((g, j))
    assert set(taps) == {(dy, dx) for dy in range(-h, h + 1) for dx in range(-h, h + 1)}
    assert all(len(v) == 1 for v in taps.values())
    for (dy, dx), owner in taps.items():
        (g, j), = owner
        mine = rows[(rows[:, 0] == g) & (rows[:, 1] == j)]
        got = sorted((int(q), int(p)) for q, p in mine[:, 4:6] if p != 255)
        want = sorted(((y + dy) * P + (x + dx), y * P + x) for y in range(P) for x in range(P)
                      if 0 <= y + dy < P and 0 <= x + dx < P)
        assert got == want, (dy, dx)
        assert len(set(mine[:, 4].tolist())) == len(mine)                      # one row per visited source position
    for g in set(rows[:, 0].tolist()):
        grp = rows[rows[:, 0] == g]
        nset = len(set(grp[:, 1].tolist()))
        ring = max(abs(int(grp[0, 2])), abs(int(grp[0, 3])))
        assert all(max(abs(int(dy)), abs(int(dx))) == ring for dy, dx in grp[:, 2:4])   # one ring per group
        slots = min(R * nt, (R - ring) * nt)
        assert nset <= max_sets and (nset == 1 or (nset * slots * fpad <= 256 and nset * slots <= 8))
        for q in set(grp[:, 4].tolist()):
            assert (grp[grp[:, 4] == q][:, 5] != 255).any()                     # a visited position feeds some tap
    if max_sets > 1 and fpad * R * nt <= 128:
        assert max(len(set(rows[rows[:, 0] == g][:, 1].tolist())) for g in set(rows[:, 0].tolist())) > 1

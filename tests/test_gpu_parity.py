"""GPU parity tests (run on the B200 box with -m gpu).  Every test goes through the C ABI of
libb200coord.so (via the ctypes mirror of the action) and compares with the CPU oracle on the same seeded
inputs, with the committed golden outputs of the real reference, and -- at BASELINE.json sizes -- with
size-independent properties.

Tolerances (BASELINE.json north_star): neighbour-list pair sets bit-exact; value, derivatives and virial
within 1e-10 relative in FP64."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import plumed2_b200 as P
from helpers import oracle_from_line, rel_err, scatter_to_system, sort_pairs, water_box
from oracle import oracle as O
from plumed2_b200 import capi

pytestmark = pytest.mark.gpu

TOL = 1e-10
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NCPU = os.cpu_count() or 4


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_outputs.npz"))


def gpu_eval(line, pos, box, step=0, **kw):
    c = P.Coordination.from_input(line, **kw)
    c.prepare(step)
    c.calculate(pos, box)
    return c


def assert_parity(c, ref, tag="", tol=TOL):
    v = c.value
    assert abs(v - ref["value"]) <= tol * max(abs(ref["value"]), 1e-300), (tag, v, ref["value"])
    assert rel_err(c.derivatives, ref["deriv"]) <= tol, (tag, "derivatives", rel_err(c.derivatives, ref["deriv"]))
    assert rel_err(c.virial, ref["virial"]) <= tol, (tag, "virial", rel_err(c.virial, ref["virial"]))


SWITCHES = ["R_0=0.3", "R_0=0.3 NN=8 MM=16", "R_0=0.25 NN=6 MM=12 D_0=0.05", "R_0=0.3 D_MAX=0.7",
            "SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 NN=12 D_MAX=0.8}",
            "SWITCH={RATIONAL R_0=0.3 NN=2 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 NN=4 MM=10 D_MAX=0.8}",
            "SWITCH={RATIONAL R_0=0.3 NN=5 MM=11 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 NN=5 D_MAX=0.8}",
            "SWITCH={RATIONAL R_0=0.3 NN=14 MM=28 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 D_MAX=0.8 NOSTRETCH}",
            "SWITCH={EXP R_0=0.2 D_MAX=0.9}", "SWITCH={EXP R_0=0.2 D_0=0.1 D_MAX=0.8}",
            "SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.8}", "SWITCH={GAUSSIAN R_0=1.0 D_MAX=0.8}",
            "SWITCH={SMAP R_0=0.3 A=4 B=3 D_MAX=0.8}", "SWITCH={CUBIC D_0=0.1 D_MAX=0.8}",
            "SWITCH={TANH R_0=0.3 D_MAX=0.8}", "SWITCH={COSINUS R_0=0.5 D_0=0.2}",
            "SWITCH={Q R_0=1.0 D_0=0.1 BETA=30.0 LAMBDA=1.5 REF=0.3 D_MAX=0.8}"]


@pytest.mark.parametrize("sw", SWITCHES)
@pytest.mark.parametrize("tri", [False, True])
def test_every_switch_no_list(sw, tri):
    """all GPU-capable switching functions, orthorhombic and triclinic minimum image, all pairs (config 1 style)"""
    n = 400
    pos, box = water_box(n, 100.0, seed=7 + tri, triclinic=tri, jitter=2.0)
    line = "c: COORDINATION GROUPA=1-%d %s" % (n, sw)
    c = gpu_eval(line, pos, box)
    assert_parity(c, oracle_from_line(line, pos, box), line)
    c.close()


NL_LINES = [
    ("single", "GROUPA=1-900 SWITCH={RATIONAL R_0=0.3 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5"),
    ("two", "GROUPA=1-150 GROUPB=151-900 SWITCH={EXP R_0=0.2 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5"),
    ("two_overlap", "GROUPA=1-500 GROUPB=300-900 SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5"),
    ("beyond_cutoff", "GROUPA=1-900 SWITCH={RATIONAL R_0=0.3 D_MAX=1.2} %s NL_CUTOFF=0.6 NL_STRIDE=5"),
]


@pytest.mark.parametrize("name,tmpl", NL_LINES)
@pytest.mark.parametrize("mode", ["NLIST", "NLISTCELLS"])
@pytest.mark.parametrize("boxkind", ["ortho", "tri", "nobox", "nopbc"])
def test_neighbour_list_modes(name, tmpl, mode, boxkind):
    """classic and cell lists, single/two lists, the three Pbc types; pair sets must be identical"""
    n = 900
    pos, box = water_box(n, 100.0, seed=11, triclinic=(boxkind == "tri"), jitter=1.5)
    line = "c: COORDINATION " + (tmpl % mode)
    if boxkind == "nopbc":
        line += " NOPBC"
    if boxkind == "nobox":
        box = None
    c = gpu_eval(line, pos, box)
    ref = oracle_from_line(line, pos, box)
    assert_parity(c, ref, line)
    np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]), err_msg=line)
    assert c.stats()["nl_size"] == ref["pairs"].shape[0]
    c.close()


def test_pair_style():
    n = 800
    pos, box = water_box(n, 100.0, seed=5, triclinic=True)
    for extra in ("", " NLIST NL_CUTOFF=0.9 NL_STRIDE=3", " NOPBC"):
        line = "c: COORDINATION GROUPA=1-400 GROUPB=401-800 SWITCH={RATIONAL R_0=0.6 D_MAX=1.5} PAIR" + extra
        c = gpu_eval(line, pos, box)
        ref = oracle_from_line(line, pos, box)
        assert_parity(c, ref, line)
        if "NLIST" in extra:
            np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
        c.close()


def test_golden_reference_outputs(gold):
    """the CUDA path against numbers produced by the real reference (tests/golden/ref_outputs.npz)"""
    for case in json.loads(str(gold["cases_json"])):
        tag, line = case["tag"], case["line"]
        key = "rt42" if tag.startswith("rt42") else ("ortho" if tag.startswith("ortho") else "tri")
        pos = gold[key + "_pos"]
        box = None if tag.startswith("nobox") else gold[key + "_box"]
        c = gpu_eval(line, pos, box)
        want_v = float(gold[tag + "_value"])
        assert abs(c.value - want_v) <= TOL * abs(want_v), (tag, c.value, want_v)
        got = scatter_to_system(pos.shape[0], c.atoms, c.derivatives)
        want_d = gold[tag + "_deriv"]
        # in the perfect rt42 crystal the derivatives cancel to ~1e-16: measure against the size of the terms
        scale = max(np.abs(want_d).max(), abs(want_v) / pos.shape[0])
        assert np.abs(got - want_d).max() <= TOL * scale, (tag, np.abs(got - want_d).max(), scale)
        assert rel_err(c.virial, gold[tag + "_virial"]) <= TOL, tag
        c.close()


def test_golden_pair_sets(gold):
    """neighbour-list pair sets of the reference's NeighborList class, bit-exact"""
    for case in json.loads(str(gold["nl_cases_json"])):
        pos = gold[case["pos"]]
        box = None if case["box"] == "zero" else gold[case["box"]]
        n0, n1 = case["n0"], case["n1"]
        groups = "GROUPA=1-%d" % n0 + (" GROUPB=%d-%d" % (n0 + 1, n0 + n1) if n1 else "")
        line = "c: COORDINATION %s R_0=0.3 %s NL_CUTOFF=%r NL_STRIDE=2%s%s" % (
            groups, "NLISTCELLS" if case["cells"] else "NLIST", case["cutoff"], " PAIR" if case["style"] == 0 else "",
            "" if case["do_pbc"] else " NOPBC")
        c = P.Coordination.from_input(line)
        c.update_list(pos, box)
        np.testing.assert_array_equal(c.neighbor_pairs(), gold[case["tag"] + "_pairs"], err_msg=case["tag"])
        c.close()


def test_regtest_rt42_values(gold):
    """regtest/basic/rt42, rt42c: 171.1815 (two identical groups) and half of it (single list)"""
    pos, box = gold["rt42_pos"], gold["rt42_box"]
    c = gpu_eval("c: COORDINATION GROUPA=1-108 GROUPB=1-108 R_0=1", pos, box)
    assert abs(c.value - 171.1815) < 5.1e-5
    c2 = gpu_eval("c: COORDINATION GROUPA=1-108 R_0=1", pos, box)
    assert abs(c2.value - 171.1815 / 2) < 5.1e-5 and abs(2 * c2.value - c.value) < 1e-9
    # rt42-cells: NLIST and NLISTCELLS agree
    a = gpu_eval("c: COORDINATION GROUPA=1-108 SWITCH={RATIONAL R_0=1 D_MAX=1.5} NLIST NL_CUTOFF=2.0 NL_STRIDE=4", pos, box)
    b = gpu_eval("c: COORDINATION GROUPA=1-108 SWITCH={RATIONAL R_0=1 D_MAX=1.5} NLISTCELLS NL_CUTOFF=2.0 NL_STRIDE=4", pos, box)
    assert abs(a.value - b.value) < 1e-10 * abs(a.value)
    for x in (c, c2, a, b):
        x.close()


@pytest.mark.parametrize("mode", ["NLIST", "NLISTCELLS"])
@pytest.mark.parametrize("tri", [False, True])
def test_frozen_list_between_rebuilds(mode, tri):
    """NL_STRIDE=4: the pair set is frozen at the rebuild step while atoms move far enough that pairs cross
    the cutoff -- value/derivatives must follow the reference's frozen-list semantics at every step"""
    n = 700
    pos0, box = water_box(n, 100.0, seed=21, triclinic=tri)
    rng = np.random.default_rng(3)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.75} %s NL_CUTOFF=0.8 NL_STRIDE=4" % (n, mode)
    c = P.Coordination.from_input(line)
    onl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=0.8, stride=4, use_cells=(mode == "NLISTCELLS"))
    pos, list_pos = pos0.copy(), None
    for step in range(3, 14):
        pos = pos + 0.03 * rng.standard_normal(pos.shape)
        rebuild = c.prepare(step)
        assert rebuild == onl.prepare(step)
        if rebuild:
            list_pos = pos.copy()
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=list_pos)
        assert_parity(c, ref, "%s step %d" % (mode, step))
        if mode == "NLIST":
            np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
    assert c.stats()["rebuilds"] == 4  # steps 3 (first), 4, 8, 12
    c.close()


@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("sw", ["RATIONAL R_0=0.3 D_MAX=0.6", "EXP R_0=0.2 D_MAX=0.6", "GAUSSIAN R_0=0.25 D_0=0.1 D_MAX=0.6"])
def test_far_parts_of_rows_and_displacement_bound(sw, tri):
    """NLIST rows are stored as near part + far part (partners beyond D_MAX + skin at the rebuild); the far parts
    are skipped while 2 x (largest displacement since the rebuild) < skin and visited (trip by trip) otherwise.
    Small steps, steps beyond the skin, atoms wrapped by a box vector and a box change must all give the
    frozen-list reference numbers."""
    n = 3000
    pos0, box = water_box(n, 100.0, seed=33, triclinic=tri)
    rng = np.random.default_rng(5)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=1.0 NL_STRIDE=100" % (n, sw)  # skin = 0.1 nm
    c = P.Coordination.from_input(line)
    c.prepare(0)
    c.calculate(pos0, box)
    direction = rng.standard_normal(pos0.shape)
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    step = 0
    for amp in (0.0, 0.01, 0.04, 0.049, 0.051, 0.09, 0.3, 0.02):
        step += 1
        pos = pos0 + amp * direction
        if amp == 0.04:  # the MD engine may re-wrap atoms into the box: a jump by a box vector is no displacement
            pos[::7] += box[0]
            pos[::11] -= box[2]
        assert not c.prepare(step)
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=pos0)
        assert_parity(c, ref, "%s amp %g" % (sw, amp))
    # a changed box (NPT): the bound is not trusted, far parts are always visited
    box2 = box * 1.01
    step += 1
    assert not c.prepare(step)
    c.calculate(pos0 * 1.01, box2)
    ref = oracle_from_line(line, pos0 * 1.01, box2, list_positions=pos0, list_box=box)
    assert_parity(c, ref, "%s scaled box" % sw)
    np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
    c.close()


@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("groups", ["GROUPA=1-6000", "GROUPA=1-1500 GROUPB=1201-6000"])
def test_rebuilds_that_filter_the_super_list(groups, tri):
    """NLIST rebuilds normally FILTER a super-list (cutoff + 10 %) instead of scanning the cells, as long as nobody has
    moved by 5 % of the cutoff since it was built.  Slow drift, atoms re-wrapped by box vectors, a jump that invalidates
    the super-list, a box change and finally motion fast enough to switch the mechanism off: the pair SET of every
    rebuild must be the reference's, and value / derivatives / virial the frozen-list numbers."""
    n = 6000
    pos0, box = water_box(n, 100.0, seed=41, triclinic=tri)
    rng = np.random.default_rng(7)
    line = "c: COORDINATION %s SWITCH={RATIONAL R_0=0.3 D_MAX=0.55} NLIST NL_CUTOFF=0.7 NL_STRIDE=2" % groups
    c = P.Coordination.from_input(line)
    pos, cur_box, list_pos, list_box = pos0.copy(), box, None, None
    plan = [0.004] * 6 + ["wrap"] + [0.004] * 3 + [0.06] + [0.004] * 4 + ["box"] + [0.004] * 3 + [0.05] * 5 + [0.004] * 2
    for step, what in enumerate(plan):
        if what == "wrap":  # the MD engine re-wraps atoms: jumps by box vectors are not displacements
            pos = pos.copy()
            pos[::5] += cur_box[1]
            pos[::9] -= cur_box[0] + cur_box[2]
        elif what == "box":
            cur_box = cur_box * 1.003
            pos = pos * 1.003
        else:
            pos = pos + what * rng.standard_normal(pos.shape) / np.sqrt(3.0)
        if c.prepare(step):
            list_pos, list_box = pos.copy(), cur_box
        c.calculate(pos, cur_box)
        ref = oracle_from_line(line, pos, cur_box, list_positions=list_pos, list_box=list_box, nthreads=8, fast_list=True)
        assert_parity(c, ref, "step %d (%s)" % (step, what))
        if step % 2 == 0:
            np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]), err_msg="step %d" % step)
    st = c.stats()
    assert st["f32_search"] == 1
    assert st["filter_rebuilds"] >= 4 and st["super_builds"] >= 3, st
    assert st["filter_rebuilds"] + st["super_builds"] < st["rebuilds"], st  # the fast stretch switched it off
    c.close()


def test_switching_the_shortcuts_off_changes_nothing():
    """B200COORD_NO_SUPERLIST / B200COORD_NO_FAR_SPLIT / B200COORD_FILTER_FLAT=0 (A/B switches): same pair sets, same
    numbers (the per-row filter kernel and the one that pipelines across rows write identical rows)"""
    n = 8000
    pos0, box = water_box(n, 100.0, seed=52)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLIST NL_CUTOFF=0.8 NL_STRIDE=3" % n
    runs = []
    for env in ({}, {"B200COORD_NO_SUPERLIST": "1"}, {"B200COORD_NO_FAR_SPLIT": "1"}, {"B200COORD_FILTER_FLAT": "0"}):
        os.environ.update(env)
        try:
            c = P.Coordination.from_input(line)
        finally:
            for k in env:
                os.environ.pop(k)
        rng = np.random.default_rng(11)
        pos, out = pos0.copy(), []
        for step in range(8):
            pos = pos + 0.002 * rng.standard_normal(pos.shape)  # stays inside the super-list's 0.04 nm bound
            c.prepare(step)
            c.calculate(pos, box)
            out.append((c.value, c.derivatives.copy(), c.virial.copy(), c.neighbor_pairs() if step % 3 == 0 else None))
        runs.append((out, c.stats()))
        c.close()
    assert runs[0][1]["filter_rebuilds"] >= 2 and runs[1][1]["filter_rebuilds"] == 0 and runs[1][1]["super_builds"] == 0
    assert runs[3][1]["filter_rebuilds"] == runs[0][1]["filter_rebuilds"]
    for other, _ in runs[1:]:
        for (v0, d0, w0, p0), (v1, d1, w1, p1) in zip(runs[0][0], other):
            assert abs(v0 - v1) <= 1e-12 * abs(v0)
            assert rel_err(d1, d0) <= 1e-11 and rel_err(w1, w0) <= 1e-11
            if p0 is not None:
                np.testing.assert_array_equal(p0, p1)


def test_exchange_step_rules():
    c = P.Coordination.from_input("c: COORDINATION GROUPA=1-50 R_0=0.3 NLIST NL_CUTOFF=1.0 NL_STRIDE=5")
    assert c.prepare(0) is True
    with pytest.raises(capi.B200CoordError) as e:  # NeighborList.cpp:447-449
        c.prepare(1, exchange_step=True)
    assert "exchange" in str(e.value)
    assert c.prepare(2) is True  # firsttime was re-armed by the exchange step (:451-453)
    c.close()


EDGE = [
    ("two atoms", 2, "GROUPA=1-2 R_0=0.5"),
    ("two atoms nlist", 2, "GROUPA=1-2 R_0=0.5 NLIST NL_CUTOFF=3.0 NL_STRIDE=1"),
    ("one vs many", 64, "GROUPA=1 GROUPB=2-64 SWITCH={EXP R_0=0.3 D_MAX=1.0} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=1"),
    ("33 atoms", 33, "GROUPA=1-33 R_0=0.3 NLIST NL_CUTOFF=0.5 NL_STRIDE=1"),
    ("one cell", 50, "GROUPA=1-50 R_0=0.3 NLISTCELLS NL_CUTOFF=5.0 NL_STRIDE=1"),
    ("two cells per axis", 300, "GROUPA=1-300 SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLIST NL_CUTOFF=0.65 NL_STRIDE=1"),
    ("two cells per axis, cells", 300, "GROUPA=1-300 SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLISTCELLS NL_CUTOFF=0.65 NL_STRIDE=1"),
    ("duplicate atoms", 40, "GROUPA=1-40,5,6,7 R_0=0.3"),
    ("duplicate atoms nlist", 40, "GROUPA=1-40,5,6,7 R_0=0.3 NLIST NL_CUTOFF=0.6 NL_STRIDE=1"),
    ("identical groups", 60, "GROUPA=1-60 GROUPB=1-60 R_0=0.3 NLIST NL_CUTOFF=0.9 NL_STRIDE=1"),
    ("identical groups cells", 60, "GROUPA=1-60 GROUPB=1-60 R_0=0.3 NLISTCELLS NL_CUTOFF=0.9 NL_STRIDE=1"),
    ("empty rows", 30, "GROUPA=1-30 SWITCH={RATIONAL R_0=0.05 D_MAX=0.1} NLIST NL_CUTOFF=0.1 NL_STRIDE=1"),
    ("reversed ranges", 100, "GROUPA=100-1:-1 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=0.9 NL_STRIDE=1"),
]


@pytest.mark.parametrize("name,n,body", EDGE)
def test_edge_cases(name, n, body):
    pos, box = water_box(n, 100.0 if n > 10 else 5.0, seed=n, jitter=0.5)
    if "per axis" in name:
        box = np.diag([1.35, 1.4, 1.45])
    line = "c: COORDINATION " + body
    c = gpu_eval(line, pos, box)
    ref = oracle_from_line(line, pos, box)
    assert_parity(c, ref, name)
    if ref["pairs"] is not None:
        np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]), err_msg=name)
    c.close()


def test_far_away_and_unwrapped_coordinates():
    """positions many box lengths outside the cell (an MD engine that never wraps)"""
    n = 500
    pos, box = water_box(n, 100.0, seed=9, triclinic=True)
    shifts = np.random.default_rng(1).integers(-7, 8, size=(n, 3)).astype(np.float64) @ box
    for mode in ("NLIST", "NLISTCELLS"):
        line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=1" % (n, mode)
        c = gpu_eval(line, pos + shifts, box)
        ref = oracle_from_line(line, pos + shifts, box)
        assert_parity(c, ref, mode)
        np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
        c.close()


def test_run_to_run_determinism_and_device_entry_point():
    n = 3000
    pos, box = water_box(n, 100.0, seed=2)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=2" % n
    a = gpu_eval(line, pos, box)
    b = gpu_eval(line, pos, box)
    assert a.value == b.value and np.array_equal(a.derivatives, b.derivatives) and np.array_equal(a.virial, b.virial)
    # device-resident entry point gives the same bits as the host one
    L = capi.lib()
    dpos, dout = C.c_void_p(), C.c_void_p()
    capi.check(L.b200coord_device_alloc(pos.nbytes, C.byref(dpos)))
    capi.check(L.b200coord_device_alloc((3 * n + 10) * 8, C.byref(dout)))
    hp = np.ascontiguousarray(pos)
    capi.check(L.b200coord_memcpy_h2d(dpos, hp.ctypes.data_as(C.c_void_p), hp.nbytes))
    b.prepare(1)
    capi.check(L.b200coord_calculate_device(b._ctx, dpos, dout), b._ctx)
    out = np.zeros(3 * n + 10)
    capi.check(L.b200coord_memcpy_d2h(out.ctypes.data_as(C.c_void_p), dout, out.nbytes))
    assert out[3 * n + 9] == a.value and np.array_equal(out[:3 * n].reshape(n, 3), a.derivatives)
    assert np.array_equal(out[3 * n:3 * n + 9].reshape(3, 3), a.virial)
    L.b200coord_device_free(dpos)
    L.b200coord_device_free(dout)
    a.close()
    b.close()


def test_rank_sharding_partials_sum_to_full():
    """i-atom sharding (cfg.rank/nranks) without a communicator: partial outputs add up to the full result,
    which is what the plugin feeds to PLUMED's Comm::Sum under MPI"""
    n = 2500
    pos, box = water_box(n, 100.0, seed=4, triclinic=True)
    for body in ("GROUPA=1-2500 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=1",
                 "GROUPA=1-400 GROUPB=401-2500 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=1",
                 "GROUPA=1-1250 GROUPB=1251-2500 R_0=0.5 PAIR"):
        line = "c: COORDINATION " + body
        full = gpu_eval(line, pos, box)
        v, d, w = 0.0, 0.0, 0.0
        for r in range(3):
            part = gpu_eval(line, pos, box, rank=r, nranks=3)
            v, d, w = v + part.value, d + part.derivatives, w + part.virial
            part.close()
        assert abs(v - full.value) <= TOL * abs(full.value)
        assert rel_err(d, full.derivatives) <= TOL and rel_err(w, full.virial) <= TOL
        full.close()


# ------------------------------------------------------------------ BASELINE.json sizes
def test_config2_100k_atoms_full_parity():
    """configs[1]: 100k atoms, orthorhombic, NLIST 1.0/10, D_MAX=0.8 -- full comparison with the oracle"""
    n = 100000
    pos, box = water_box(n, 100.0)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n
    c = gpu_eval(line, pos, box)
    ref = oracle_from_line(line, pos, box, nthreads=NCPU, fast_list=True)
    assert_parity(c, ref, "config2")
    np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
    # second frame on the frozen list
    pos2 = pos + 0.005 * np.random.default_rng(0).standard_normal(pos.shape)
    assert c.prepare(1) is False
    c.calculate(pos2, box)
    assert_parity(c, oracle_from_line(line, pos2, box, list_positions=pos, nthreads=NCPU, fast_list=True), "config2 frame 2")
    c.close()


def test_config3_solute_solvent_triclinic_exp():
    """configs[2] scaled to what the O(N) oracle finishes quickly: 2k solute vs 200k solvent, triclinic, EXP,
    rebuild every step; NLISTCELLS and NLIST must agree with the oracle and with each other"""
    na, nb = 2000, 200000
    pos, box = water_box(na + nb, 100.0, seed=33, triclinic=True)
    base = "c: COORDINATION GROUPA=1-%d GROUPB=%d-%d SWITCH={EXP R_0=0.2 D_MAX=0.9} %%s NL_CUTOFF=1.0 NL_STRIDE=1" % (na, na + 1, na + nb)
    a = gpu_eval(base % "NLIST", pos, box)
    b = gpu_eval(base % "NLISTCELLS", pos, box)
    ref = oracle_from_line(base % "NLIST", pos, box, nthreads=NCPU, fast_list=True)
    assert_parity(a, ref, "config3 NLIST")
    assert_parity(b, ref, "config3 NLISTCELLS")
    np.testing.assert_array_equal(a.neighbor_pairs(), sort_pairs(ref["pairs"]))
    a.close()
    b.close()


def test_one_million_atoms_full_oracle_comparison():
    """headline size, the whole thing: value, 3 M derivatives and virial of 1 M atoms against the oracle (its O(N) list on all
    host cores), on the rebuild step and on a step of the drifting trajectory that keeps the list; pair count from the list"""
    n = 1000000
    pos, box = water_box(n, 100.0, seed=123)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n
    c = gpu_eval(line, pos, box)
    ref = oracle_from_line(line, pos, box, nthreads=NCPU, fast_list=True)
    assert_parity(c, ref, "1M rebuild step")
    assert c.stats()["nl_size"] == ref["pairs"].shape[0]
    del ref
    pos2 = pos + 0.004 * np.random.default_rng(2).standard_normal(pos.shape)
    pos2[::9] += box[1]  # the MD engine re-wrapped some molecules
    assert c.prepare(3) is False
    c.calculate(pos2, box)
    assert_parity(c, oracle_from_line(line, pos2, box, list_positions=pos, nthreads=NCPU, fast_list=True), "1M frozen list")
    c.close()


def test_one_million_atoms_properties():
    """headline size (1M atoms, NLIST): properties that need no O(N) oracle run"""
    n = 1000000
    pos, box = water_box(n, 100.0, seed=99)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n
    c = gpu_eval(line, pos, box)
    v, d, w = c.value, c.derivatives.copy(), c.virial.copy()
    scale = np.abs(d).max()
    assert np.abs(d.sum(axis=0)).max() <= 1e-9 * scale * np.sqrt(n)          # Newton's third law
    assert np.abs(w - w.T).max() <= 1e-12 * np.abs(w).max()                   # symmetric virial
    nl = c.stats()["nl_size"]
    expect = 0.5 * n * (4.0 / 3.0 * np.pi) * 100.0                            # pairs within 1.0 nm at 100/nm^3
    assert abs(nl - expect) < 0.01 * expect
    # rigid translation by an arbitrary vector + re-wrapping by whole box vectors changes nothing
    c2 = gpu_eval(line, pos + np.array([0.123, -4.56, 7.89]) + box[0] - 2 * box[2], box)
    assert abs(c2.value - v) <= TOL * abs(v) and rel_err(c2.derivatives, d) <= 1e-9
    assert c2.stats()["nl_size"] == nl
    # permutation of the atom order permutes the derivatives
    perm = np.random.default_rng(5).permutation(n)
    c3 = gpu_eval(line, pos[perm], box)
    assert abs(c3.value - v) <= TOL * abs(v) and rel_err(c3.derivatives, d[perm]) <= TOL
    assert rel_err(c3.virial, w) <= TOL
    # the cell-list flavour agrees
    c4 = gpu_eval(line.replace("NLIST", "NLISTCELLS"), pos, box)
    assert abs(c4.value - v) <= TOL * abs(v) and rel_err(c4.derivatives, d) <= TOL
    # 300 randomly chosen atoms against a brute-force numpy evaluation of their rows
    sw = O.make_switch("RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8")
    L = box[0, 0]
    for i in np.random.default_rng(6).choice(n, 300, replace=False):
        dd = pos - pos[i]
        dd -= L * np.round(dd / L)
        r2 = (dd * dd).sum(axis=1)
        idx = np.nonzero((r2 <= 0.64) & (r2 > 0))[0]
        g = np.zeros(3)
        for j in idx:
            s, df = O.switch_calculate_sqr(sw, r2[j])
            g -= df * dd[j]
        assert np.abs(g - d[i]).max() <= 1e-9 * scale
    for x in (c, c2, c3, c4):
        x.close()


def test_rt20_switch_regtests_on_gpu():
    """the reference's own rt20-switch-* fixtures (value + 15 derivatives per frame, 12 switch types x stretch on/off)"""
    with open(os.path.join(GOLD, "ref_regtest_kats.json")) as f:
        rt = json.load(f)["rt20_switch"]
    frames = rt["frames"]
    bad = []
    for name, test in rt["tests"].items():
        col = np.array(test["colvar"])
        der = np.array(test["deriv"]).reshape(len(frames), 15, 4)
        for ci, lab in enumerate(("c", "cs")):
            c = P.Coordination.from_input(test["lines"][lab])
            for fi, fr in enumerate(frames):
                c.prepare(fi)
                c.calculate(np.array(fr["pos"]), np.diag(fr["box"]))
                got = np.concatenate([c.derivatives.ravel(), c.virial.ravel()])
                ev, ed = abs(c.value - col[fi, 1 + ci]), np.abs(got - der[fi, :, 2 + ci]).max()
                if ev >= 6e-7 or ed >= 5.1e-5:
                    bad.append((name, lab, fi, fr["pos"], float(c.value), float(col[fi, 1 + ci]), float(ed)))
            c.close()
    assert not bad, bad


@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("body", ["GROUPA=1-5000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.6}",
                                  "GROUPA=1-5000 SWITCH={RATIONAL R_0=0.3 NN=8 MM=16 D_MAX=0.6}",
                                  "GROUPA=1-800 GROUPB=601-5000 SWITCH={EXP R_0=0.2 D_MAX=0.6}",
                                  "GROUPA=1-5000 SWITCH={GAUSSIAN R_0=0.25 D_0=0.1 D_MAX=0.6}",
                                  "GROUPA=1-5000 R_0=0.2"])
def test_image_sweep_against_oracle_and_general_sweep(body, tri):
    """The image-mode sweep (continuous coordinates, periodic image stored in the list entry, virial from positions;
    sweep_img.cuh) against the oracle and against the general sweep (minimum image per pair) on the same steps:
    rebuild steps, drifting atoms with far parts skipped and visited, atoms re-wrapped by the MD engine."""
    n = 5000
    pos0, box = water_box(n, 100.0, seed=61, triclinic=tri)
    line = "c: COORDINATION %s NLIST NL_CUTOFF=0.8 NL_STRIDE=4" % body
    runs = []
    for env in ({}, {"B200COORD_NO_IMG_SWEEP": "1"}):
        os.environ.update(env)
        try:
            c = P.Coordination.from_input(line)
        finally:
            for k in env:
                os.environ.pop(k)
        rng = np.random.default_rng(13)
        pos, out, list_pos = pos0.copy(), [], None
        for step in range(9):
            pos = pos + (0.02 if step in (5, 6) else 0.004) * rng.standard_normal(pos.shape)
            if step == 3:
                pos[::6] += box[1]
                pos[::10] -= box[0] + box[2]
            if c.prepare(step):
                list_pos = pos.copy()
            c.calculate(pos, box)
            if not env:
                ref = oracle_from_line(line, pos, box, list_positions=list_pos, nthreads=8, fast_list=True)
                assert_parity(c, ref, "step %d" % step)
            st = c.stats()
            out.append((c.value, c.derivatives.copy(), c.virial.copy(), st["pair_evals"], st["nl_size"]))
        runs.append(out)
        c.close()
    for (v0, d0, w0, e0, s0), (v1, d1, w1, e1, s1) in zip(*runs):
        assert abs(v0 - v1) <= 1e-12 * abs(v0)
        assert rel_err(d0, d1) <= 1e-11 and rel_err(w0, w1) <= 1e-11
        assert s0 == s1 and e0 == e1  # same list, same parts of it visited


def test_update_list_then_calculate_uses_the_new_list_and_its_displacement_origin():
    """calculate(P0) [rebuild], update_list(P1), calculate(P2 close to P0): the list, and the origin of the displacement
    bound that lets the sweep skip far parts, are those of P1"""
    n = 4000
    pos0, box = water_box(n, 100.0, seed=71)
    rng = np.random.default_rng(17)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLIST NL_CUTOFF=0.7 NL_STRIDE=50" % n
    c = P.Coordination.from_input(line)
    c.prepare(0)
    c.calculate(pos0, box)
    pos1 = pos0 + 0.05 * rng.standard_normal(pos0.shape)
    c.update_list(pos1, box)
    pos2 = pos0 + 0.001 * rng.standard_normal(pos0.shape)
    assert not c.prepare(1)
    c.calculate(pos2, box)
    ref = oracle_from_line(line, pos2, box, list_positions=pos1, nthreads=8, fast_list=True)
    assert_parity(c, ref, "after update_list")
    np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]))
    c.close()


@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("sw", ["SWITCH={EXP R_0=0.2 D_MAX=0.7}", "SWITCH={RATIONAL R_0=0.3 D_MAX=0.7}", "R_0=0.25"])
def test_few_group_a_atoms_scatter_to_their_partners(sw, tri):
    """A solute in its solvent (GROUPA 8x smaller than GROUPB or less): the GROUPB rows hold a handful of entries each, so
    the image sweep does not visit them -- the GROUPA rows add +dd to their partners' derivatives instead.  Against the
    oracle, and against the same steps with B200COORD_NO_SCATTER=1 (every GROUPB row swept)."""
    n, na = 6000, 300
    pos0, box = water_box(n, 100.0, seed=81, triclinic=tri)
    line = "c: COORDINATION GROUPA=1-%d GROUPB=%d-%d %s NLIST NL_CUTOFF=0.8 NL_STRIDE=3" % (na, na + 1, n, sw)
    runs = []
    for env in ({}, {"B200COORD_NO_SCATTER": "1"}):
        os.environ.update(env)
        try:
            c = P.Coordination.from_input(line)
        finally:
            for k in env:
                os.environ.pop(k)
        rng = np.random.default_rng(19)
        pos, out, list_pos = pos0.copy(), [], None
        for step in range(5):
            pos = pos + 0.005 * rng.standard_normal(pos.shape)
            if c.prepare(step):
                list_pos = pos.copy()
            c.calculate(pos, box)
            if not env:
                assert_parity(c, oracle_from_line(line, pos, box, list_positions=list_pos, nthreads=8, fast_list=True), "step %d" % step)
            out.append((c.value, c.derivatives.copy(), c.virial.copy()))
        runs.append(out)
        c.close()
    for (v0, d0, w0), (v1, d1, w1) in zip(*runs):
        assert abs(v0 - v1) <= 1e-12 * abs(v0) and rel_err(d0, d1) <= 1e-11 and rel_err(w0, w1) <= 1e-11


@pytest.mark.parametrize("groups", ["GROUPA=1-4000", "GROUPA=1-700 GROUPB=501-4000"])
def test_neighbour_list_as_a_tool_device_pairs(groups):
    """b200coord_nl_pairs_device: the list handed out on the DEVICE as (i0,i1) index pairs, what the reference's
    NeighborList is to ContactMap (src/colvar/ContactMap.cpp:190-260).  The pair set must be the oracle's (minus pairs of one
    and the same atom), and a consumer that sums a switching function over the pairs -- a CONTACTMAP SUM -- must reproduce
    COORDINATION's value."""
    import torch
    n = 4000
    pos, box = water_box(n, 100.0, seed=91)
    line = "c: COORDINATION %s SWITCH={RATIONAL R_0=0.3 D_MAX=0.7} NLIST NL_CUTOFF=0.8 NL_STRIDE=5" % groups
    c = gpu_eval(line, pos, box)
    L = capi.lib()
    npairs = C.c_ulonglong(0)
    capi.check(L.b200coord_nl_pairs_device(c._ctx, None, 0, C.byref(npairs)), c._ctx)
    buf = torch.empty((npairs.value, 2), dtype=torch.int32, device="cuda")
    capi.check(L.b200coord_nl_pairs_device(c._ctx, C.c_void_p(buf.data_ptr()), npairs.value, C.byref(npairs)), c._ctx)
    ref = oracle_from_line(line, pos, box, nthreads=8, fast_list=True)
    atoms = ref["atoms"]
    mine = buf.cpu().numpy().astype(np.uint32)
    rp = sort_pairs(ref["pairs"])
    rp = rp[atoms[rp[:, 0]] != atoms[rp[:, 1]]]  # the groups share atoms 501-700: the reference lists those self pairs
    np.testing.assert_array_equal(sort_pairs(mine), rp)
    np.testing.assert_array_equal(sort_pairs(mine), c.neighbor_pairs()[atoms[c.neighbor_pairs()[:, 0]] != atoms[c.neighbor_pairs()[:, 1]]])
    # the consumer: sum of s(r) over the pairs, on the device
    x = torch.tensor(pos[atoms], device="cuda")
    Ld = torch.tensor(np.diag(box), device="cuda")
    d = x[buf[:, 1].long()] - x[buf[:, 0].long()]
    d = d - torch.round(d / Ld) * Ld
    r2 = (d * d).sum(1)
    sw = O.make_switch("RATIONAL R_0=0.3 D_MAX=0.7")
    y = r2 * sw.invr0_2
    s = torch.where(r2 <= sw.dmax_2, (1.0 / (1.0 + y ** 3)) * sw.stretch + sw.shift, torch.zeros_like(r2))
    assert abs(float(s.sum().item()) - c.value) <= 1e-10 * abs(c.value)
    c.close()


@pytest.mark.parametrize("line", [
    "c: COORDINATION GROUPA=1-5000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
    "c: COORDINATION GROUPA=1-500 GROUPB=501-5000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=2",
    "c: COORDINATION GROUPA=1-2500 GROUPB=2501-5000 PAIR R_0=0.3",
])
def test_frames_submitted_ahead_give_the_synchronous_results(line):
    """b200coord_submit / _collect (two steps in flight, copies on their own streams): every frame's value, derivatives
    and virial are bit for bit what b200coord_calculate returns for the same sequence"""
    import ctypes as C
    n = 5000
    pos0, box = water_box(n, 100.0, seed=77, triclinic=True)
    rng = np.random.default_rng(5)
    frames = []
    p = pos0
    for _ in range(11):
        p = p + 0.004 * rng.standard_normal(p.shape)
        frames.append(np.ascontiguousarray(p))
    ref = []
    c = P.Coordination.from_input(line)
    for step, f in enumerate(frames):
        c.prepare(step)
        c.calculate(f, box)
        ref.append((c.value, c.derivatives.copy(), c.virial.copy()))
    c.close()
    c = P.Coordination.from_input(line)
    c._set_box(box)
    L = c._L
    vals = [C.c_double(0) for _ in frames]
    ders = [np.full((c.n, 3), np.nan) for _ in frames]
    virs = [np.full(9, np.nan) for _ in frames]
    gathered = [c.gather(f).copy() for f in frames]
    for step in range(len(frames)):
        c.prepare(step)
        capi.check(L.b200coord_submit(c._ctx, gathered[step].ctypes.data_as(C.c_void_p), C.byref(vals[step]),
                                      ders[step].ctypes.data_as(C.c_void_p), virs[step].ctypes.data_as(C.c_void_p)), c._ctx)
        if step >= 2:  # the step submitted two calls ago has been delivered
            assert abs(vals[step - 2].value - ref[step - 2][0]) <= 1e-13 * abs(ref[step - 2][0])
    capi.check(L.b200coord_collect(c._ctx), c._ctx)
    # few GROUPA atoms among many GROUPB atoms: the GROUPA rows add to their partners with atomics, whose
    # order is not fixed from run to run; everything else is bit-reproducible
    exact = "NLISTCELLS" not in line
    for step in range(len(frames)):
        if exact:
            assert vals[step].value == ref[step][0], step
            np.testing.assert_array_equal(ders[step], ref[step][1], err_msg="step %d" % step)
            np.testing.assert_array_equal(virs[step].reshape(3, 3), ref[step][2].reshape(3, 3), err_msg="step %d" % step)
        else:
            assert abs(vals[step].value - ref[step][0]) <= 1e-13 * abs(ref[step][0]), step
            assert rel_err(ders[step], ref[step][1]) <= 1e-13 and rel_err(virs[step].reshape(3, 3), ref[step][2].reshape(3, 3)) <= 1e-13
    c.close()

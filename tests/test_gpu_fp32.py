"""Opt-in FP32 sweep (B200COORD_FP32; BASELINE.json north_star: "1e-5 in an opt-in FP32 mode").

Same C ABI, same bit-exact neighbour list; only the pair arithmetic of the list / cell sweeps changes (FP64 minimum
image, FP32 r^2 / switching function / row sums, FP64 accumulation across rows).  Every case is compared with the CPU
oracle on the same inputs.  Tolerance: 1e-5 -- relative for the value, relative to the largest reference component
for derivatives and virial (helpers.rel_err), as in the FP64 tests."""
import numpy as np
import pytest

import plumed2_b200 as P
from helpers import oracle_from_line, rel_err, sort_pairs, water_box
from oracle import oracle as O
from plumed2_b200 import capi

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def assert_parity32(c, ref, tag=""):
    assert abs(c.value - ref["value"]) <= TOL32 * max(abs(ref["value"]), 1e-300), (tag, c.value, ref["value"])
    ed, ev = rel_err(c.derivatives, ref["deriv"]), rel_err(c.virial, ref["virial"])
    assert ed <= TOL32, (tag, "derivatives", ed)
    assert ev <= TOL32, (tag, "virial", ev)


SWITCHES = ["R_0=0.3", "R_0=0.3 NN=8 MM=16", "R_0=0.25 NN=6 MM=12 D_0=0.05",
            "SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 NN=2 D_MAX=0.8}",
            "SWITCH={RATIONAL R_0=0.3 NN=4 MM=10 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 NN=5 MM=11 D_MAX=0.8}",
            "SWITCH={RATIONAL R_0=0.3 NN=5 D_MAX=0.8}", "SWITCH={RATIONAL R_0=0.3 D_MAX=0.8 NOSTRETCH}",
            "SWITCH={EXP R_0=0.2 D_MAX=0.9}", "SWITCH={EXP R_0=0.2 D_0=0.1 D_MAX=0.8}",
            "SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.8}", "SWITCH={GAUSSIAN R_0=1.0 D_MAX=0.8}",
            "SWITCH={SMAP R_0=0.3 A=4 B=3 D_MAX=0.8}", "SWITCH={CUBIC D_0=0.1 D_MAX=0.8}",
            "SWITCH={TANH R_0=0.3 D_MAX=0.8}", "SWITCH={COSINUS R_0=0.5 D_0=0.2}",
            "SWITCH={Q R_0=1.0 D_0=0.1 BETA=30.0 LAMBDA=1.5 REF=0.3 D_MAX=0.8}"]


@pytest.mark.parametrize("sw", SWITCHES)
@pytest.mark.parametrize("tri", [False, True])
def test_fp32_every_switch_no_list(sw, tri):
    """all switching-function kinds through the FP32 cell sweep (no list = one cell), both minimum images"""
    n = 400
    pos, box = water_box(n, 100.0, seed=7 + tri, triclinic=tri, jitter=2.0)
    line = "c: COORDINATION GROUPA=1-%d %s" % (n, sw)
    c = P.Coordination.from_input(line, precision=capi.FP32)
    c.prepare(0)
    c.calculate(pos, box)
    assert_parity32(c, oracle_from_line(line, pos, box), line)
    c.close()


NL_LINES = [
    "GROUPA=1-900 SWITCH={RATIONAL R_0=0.3 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5",
    "GROUPA=1-150 GROUPB=151-900 SWITCH={EXP R_0=0.2 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5",
    "GROUPA=1-500 GROUPB=300-900 SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.7} %s NL_CUTOFF=0.8 NL_STRIDE=5",
    "GROUPA=1-900 SWITCH={RATIONAL R_0=0.3 D_MAX=1.2} %s NL_CUTOFF=0.6 NL_STRIDE=5",
]


@pytest.mark.parametrize("tmpl", NL_LINES)
@pytest.mark.parametrize("mode", ["NLIST", "NLISTCELLS"])
@pytest.mark.parametrize("boxkind", ["ortho", "tri", "nobox", "nopbc"])
def test_fp32_neighbour_list_modes(tmpl, mode, boxkind):
    """list and cell sweeps in FP32; the neighbour list itself stays bit-exact"""
    n = 900
    pos, box = water_box(n, 100.0, seed=11, triclinic=(boxkind == "tri"), jitter=1.5)
    line = "c: COORDINATION " + (tmpl % mode)
    if boxkind == "nopbc":
        line += " NOPBC"
    if boxkind == "nobox":
        box = None
    c = P.Coordination.from_input(line, precision=capi.FP32)
    c.prepare(0)
    c.calculate(pos, box)
    ref = oracle_from_line(line, pos, box)
    assert_parity32(c, ref, line)
    np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]), err_msg=line)
    c.close()


@pytest.mark.parametrize("tri", [False, True])
def test_fp32_frozen_list_near_and_far_parts(tri):
    """steps between rebuilds: small moves (far parts of the rows skipped) and then large ones (far parts visited)"""
    n = 6000
    pos0, box = water_box(n, 100.0, seed=31, triclinic=tri)
    rng = np.random.default_rng(5)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLIST NL_CUTOFF=1.0 NL_STRIDE=6" % n
    c = P.Coordination.from_input(line, precision=capi.FP32)
    pos, list_pos = pos0.copy(), None
    for step in range(12):
        pos = pos + (0.004 if step < 6 else 0.05) * rng.standard_normal(pos.shape)
        if c.prepare(step):
            list_pos = pos.copy()
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=list_pos, nthreads=8)
        assert_parity32(c, ref, "step %d" % step)
    c.close()


def test_fp32_against_fp64_at_100k_atoms():
    """BASELINE config 2 size: FP32 and FP64 contexts on the same frames; same list, results within 1e-5"""
    n = 100000
    pos0, box = water_box(n, 100.0, seed=3)
    rng = np.random.default_rng(9)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n
    c64 = P.Coordination.from_input(line)
    c32 = P.Coordination.from_input(line, precision=capi.FP32)
    pos = pos0
    for step in range(3):
        c64.prepare(step)
        c32.prepare(step)
        c64.calculate(pos, box)
        c32.calculate(pos, box)
        assert c32.stats()["nl_size"] == c64.stats()["nl_size"]
        assert abs(c32.value - c64.value) <= TOL32 * abs(c64.value), (step, c32.value, c64.value)
        assert rel_err(c32.derivatives, c64.derivatives) <= TOL32, step
        assert rel_err(c32.virial, c64.virial) <= TOL32, step
        assert np.any(c32.derivatives != c64.derivatives)  # it really is a different arithmetic
        pos = pos + 0.003 * rng.standard_normal(pos.shape)
    c64.close()
    c32.close()


def test_fp32_dhenergy_typed_sibling():
    n = 3000
    pos, box = water_box(n, 100.0, seed=8, triclinic=True)
    q = np.random.default_rng(2).standard_normal(n)
    line = "c: DHENERGY GROUPA=1-1200 GROUPB=1201-%d I=0.15 EPSILON=78.0 TEMP=298 NLIST NL_CUTOFF=1.2 NL_STRIDE=3" % n
    c = P.Coordination.from_input(line, precision=capi.FP32)
    c.set_charges(q)
    c.prepare(0)
    c.calculate(pos, box)
    ref = oracle_from_line(line, pos, box, charges=q, nthreads=8)
    # an energy of mixed-sign charges cancels: compare with the scale of the sum of |terms| through the derivatives
    assert abs(c.value - ref["value"]) <= TOL32 * max(abs(ref["value"]), np.abs(ref["deriv"]).max()), (c.value, ref["value"])
    assert rel_err(c.derivatives, ref["deriv"]) <= TOL32
    assert rel_err(c.virial, ref["virial"]) <= TOL32
    c.close()


@pytest.mark.parametrize("mode", ["", "NLIST NL_CUTOFF=0.8 NL_STRIDE=3", "NLISTCELLS NL_CUTOFF=0.8 NL_STRIDE=3"])
def test_fp32_ghbfix_typed_sibling(mode):
    n = 4000
    pos, box = water_box(n, 100.0, seed=9, triclinic=True)
    rng = np.random.default_rng(4)
    ntypes = 5
    types = rng.integers(0, ntypes, n).astype(np.uint32)
    etas = rng.standard_normal((ntypes, ntypes))  # asymmetric on purpose
    line = "c: GHBFIX GROUPA=1-%d D_0=0.2 D_MAX=0.6 C=0.7 TYPES=x PARAMS=y %s" % (n, mode)
    c = P.Coordination.from_input(line, precision=capi.FP32)
    c.set_types(types, ntypes, etas)
    c.prepare(0)
    c.calculate(pos, box)
    ref = oracle_from_line(line, pos, box, types=(types, ntypes, etas.ravel()), nthreads=8)
    # mixed-sign scaling parameters: the sum cancels, compare on the scale of its largest derivative
    assert abs(c.value - ref["value"]) <= TOL32 * max(abs(ref["value"]), np.abs(ref["deriv"]).max()), (c.value, ref["value"])
    assert rel_err(c.derivatives, ref["deriv"]) <= TOL32, rel_err(c.derivatives, ref["deriv"])
    assert rel_err(c.virial, ref["virial"]) <= TOL32
    c.close()


def test_precision_outside_the_enum_is_refused():
    with pytest.raises(capi.B200CoordError):
        P.Coordination.from_input("c: COORDINATION GROUPA=1-10 R_0=0.3", precision=7)

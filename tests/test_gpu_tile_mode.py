"""The opt-in shared-memory tile sweep (B200COORD_TILE=1, kernels_tile.cu: 16-bit tile-local neighbour list) must
give the same pair sets and the same numbers as the default list sweep and the oracle."""
import os

import numpy as np
import pytest

import plumed2_b200 as P
from helpers import oracle_from_line, rel_err, sort_pairs, water_box

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tile_env():
    os.environ["B200COORD_TILE"] = "1"
    yield
    os.environ.pop("B200COORD_TILE", None)


# boxes must hold >= pencil + 2*radius = 10 half-cutoff cells along x for the fine grid (else the engine keeps the list sweep)
CASES = [
    ("single ortho", 30000, False, "GROUPA=1-30000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3"),
    ("single triclinic", 30000, True, "GROUPA=1-30000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLIST NL_CUTOFF=1.0 NL_STRIDE=3"),
    ("two lists", 30000, True, "GROUPA=1-4000 GROUPB=4001-30000 SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.9} NLIST NL_CUTOFF=1.0 NL_STRIDE=3"),
    ("overlapping groups", 24000, False, "GROUPA=1-15000 GROUPB=9000-24000 R_0=0.3 NLIST NL_CUTOFF=0.9 NL_STRIDE=3"),
    ("coarse grid (radius 1)", 3000, False, "GROUPA=1-3000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.6} NLIST NL_CUTOFF=0.7 NL_STRIDE=3"),
    ("no box", 4000, False, "GROUPA=1-4000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=0.9 NL_STRIDE=3"),
]


@pytest.mark.parametrize("name,n,tri,body", CASES)
def test_tile_sweep_parity(tile_env, name, n, tri, body):
    pos0, box = water_box(n, 100.0, seed=len(name), triclinic=tri, jitter=1.0)
    if name == "no box":
        box = None
    line = "c: COORDINATION " + body
    c = P.Coordination.from_input(line)
    rng = np.random.default_rng(1)
    pos, list_pos = pos0.copy(), None
    for step in range(4):
        pos = pos + 0.01 * rng.standard_normal(pos.shape)
        if c.prepare(step):
            list_pos = pos.copy()
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=list_pos, nthreads=8, fast_list=True)
        assert abs(c.value - ref["value"]) <= 1e-10 * abs(ref["value"]), (name, step)
        assert rel_err(c.derivatives, ref["deriv"]) <= 1e-10 and rel_err(c.virial, ref["virial"]) <= 1e-10, (name, step)
        np.testing.assert_array_equal(c.neighbor_pairs(), sort_pairs(ref["pairs"]), err_msg=name)
    st = c.stats()
    assert st["f32_search"] == 1, (name, st)
    if "coarse" not in name:  # 3000 atoms: too few cells along x for a pencil, the engine keeps the list sweep
        assert st["tile_mode"] == 1, (name, st)
    c.close()


def test_tile_and_list_sweeps_agree_at_100k(tile_env):
    n = 100000
    pos, box = water_box(n, 100.0)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n
    t = P.Coordination.from_input(line)
    t.prepare(0)
    t.calculate(pos, box)
    assert t.stats()["tile_mode"] == 1
    os.environ.pop("B200COORD_TILE", None)
    l = P.Coordination.from_input(line)
    l.prepare(0)
    l.calculate(pos, box)
    assert l.stats()["tile_mode"] == 0
    assert abs(t.value - l.value) <= 1e-12 * abs(l.value)
    assert rel_err(t.derivatives, l.derivatives) <= 1e-11 and rel_err(t.virial, l.virial) <= 1e-11
    assert np.array_equal(t.neighbor_pairs(), l.neighbor_pairs())
    t.close()
    l.close()

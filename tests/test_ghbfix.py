"""GHBFIX (src/colvar/GHBFIX.cpp), the typed sibling of COORDINATION on CoordinationBase: same groups,
lists and loop; the pairing is a piecewise polynomial scaled by eta[type of the pair's first atom][type of its second].
Golden vectors come from the real reference (oracle/gen_golden_ghbfix.py -> tests/golden/ref_ghbfix.npz)."""
import json
import os

import numpy as np
import pytest

from helpers import oracle_from_line, rel_err, water_box
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold(tmp_path_factory):
    g = np.load(os.path.join(GOLD, "ref_ghbfix.npz"))
    d = tmp_path_factory.mktemp("ghbfix")
    tf, pf = d / "types.dat", d / "params.dat"
    tf.write_text(str(g["types_txt"]))
    pf.write_text(str(g["params_txt"]))
    return g, json.loads(str(g["cases_json"])), str(tf), str(pf)


def _frame(g, tag):
    return (g["pos_o"], g["box_o"]) if tag.startswith("o_") else (g["pos_t"], g["box_t"])


def _units(line):
    return "kcal/mol" if "ENERGY_UNITS=kcal/mol" in line else "plumed"


def test_table_reader_keeps_the_map_semantics(gold):
    g, _, tf, pf = gold
    types, n, etas = O.read_ghbfix_tables(tf, pf)
    assert n == 4 and types[0] == 0 and types.shape == (400,)
    e = etas.reshape(4, 4)
    names = {}
    for ln in str(g["types_txt"]).splitlines()[1:]:
        names.setdefault(ln, len(names))
    assert e[names["don"], names["acc"]] == -2.0 and e[names["acc"], names["don"]] == -1.5
    assert e[names["ion"], names["ion"]] == 0.0           # pair of types without a row
    assert e[0, names["ion"]] == 0.25                      # "ghost" is unknown to the types file -> index 0 (operator[])
    assert np.allclose(O.read_ghbfix_tables(tf, pf, "kcal/mol")[2], etas * 4.184)


def test_oracle_matches_reference_ghbfix(gold):
    g, cases, tf, pf = gold
    for case in cases:
        pos, box = _frame(g, case["tag"])
        tables = O.read_ghbfix_tables(tf, pf, _units(case["line"]))
        ref = oracle_from_line(case["line"], pos, box, types=tables)
        deriv = np.zeros((pos.shape[0], 3))
        np.add.at(deriv, ref["atoms"], ref["deriv"])
        want_v, want_d, want_vir = float(g[case["tag"] + "_value"]), g[case["tag"] + "_deriv"], g[case["tag"] + "_virial"]
        assert abs(ref["value"] - want_v) <= 1e-12 * max(1.0, abs(want_v)), case
        assert rel_err(deriv, want_d) <= 1e-12, case
        assert rel_err(ref["virial"], want_vir) <= 1e-12, case


@pytest.mark.gpu
def test_gpu_ghbfix_matches_reference_goldens(gold):
    import plumed2_b200 as P
    g, cases, tf, pf = gold
    for case in cases:
        pos, box = _frame(g, case["tag"])
        c = P.Coordination.from_input(case["line"] + " TYPES=%s PARAMS=%s" % (tf, pf))
        assert c.ghbfix_files[:2] == (tf, pf)
        c.set_types(*O.read_ghbfix_tables(tf, pf, _units(case["line"])))
        c.prepare(0)
        c.calculate(pos, box)
        deriv = np.zeros((pos.shape[0], 3))
        np.add.at(deriv, c.atoms, c.derivatives)
        want_v, want_d, want_vir = float(g[case["tag"] + "_value"]), g[case["tag"] + "_deriv"], g[case["tag"] + "_virial"]
        assert abs(c.value - want_v) <= 1e-10 * max(1.0, abs(want_v)), (case, c.value, want_v)
        assert rel_err(deriv, want_d) <= 1e-10, case
        assert rel_err(c.virial, want_vir) <= 1e-10, case
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["", "NLIST NL_CUTOFF=0.8 NL_STRIDE=3", "NLISTCELLS NL_CUTOFF=0.8 NL_STRIDE=3"])
def test_gpu_ghbfix_frozen_list_larger_system(gold, mode):
    import plumed2_b200 as P
    _, _, tf, pf = gold
    n = 6000
    pos0, box = water_box(n, 100.0, seed=9, triclinic=True)
    rng = np.random.default_rng(4)
    ntypes = 5
    types = rng.integers(0, ntypes, n).astype(np.uint32)
    etas = rng.standard_normal((ntypes, ntypes))  # asymmetric on purpose
    line = "c: GHBFIX GROUPA=1-%d D_0=0.2 D_MAX=0.6 C=0.7 TYPES=x PARAMS=y %s" % (n, mode)
    c = P.Coordination.from_input(line)
    with pytest.raises(P.capi.B200CoordError):  # types are compulsory
        c.prepare(0)
        c.calculate(pos0, box)
    with pytest.raises(P.capi.B200CoordError):
        c.set_types(types, ntypes - 1, etas[:-1, :-1])
    c.set_types(types, ntypes, etas)
    pos, list_pos = pos0.copy(), None
    for step in range(5):
        pos = pos + 0.012 * rng.standard_normal(pos.shape)
        if c.prepare(step) or list_pos is None:
            list_pos = pos.copy()
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=list_pos if mode else None, types=(types, ntypes, etas.ravel()),
                               nthreads=8)
        assert abs(c.value - ref["value"]) <= 1e-10 * abs(ref["value"]), (mode, step)
        assert rel_err(c.derivatives, ref["deriv"]) <= 1e-10 and rel_err(c.virial, ref["virial"]) <= 1e-10, (mode, step)
    c.close()

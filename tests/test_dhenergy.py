"""DHENERGY (src/colvar/DHEnergy.cpp), the sibling of COORDINATION on CoordinationBase: same groups, lists
and loop, Debye-Hueckel pairing with per-atom charges.  Golden vectors come from the real reference
(oracle/gen_golden_dhenergy.py -> tests/golden/ref_dhenergy.npz)."""
import json
import os

import numpy as np
import pytest

from helpers import oracle_from_line, rel_err, water_box

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    g = np.load(os.path.join(GOLD, "ref_dhenergy.npz"))
    return g, json.loads(str(g["cases_json"]))


def _frame(g, tag):
    if tag.startswith("o_"):
        return g["pos_o"], g["box_o"]
    return g["pos_t"], g["box_t"]


def test_oracle_matches_reference_dhenergy():
    """the C restatement of DHEnergy::pairing against the real reference, all list styles"""
    g, cases = _gold()
    for case in cases:
        pos, box = _frame(g, case["tag"])
        ref = oracle_from_line(case["line"], pos, box, charges=g["q"])
        n = pos.shape[0]
        deriv = np.zeros((n, 3))
        np.add.at(deriv, ref["atoms"], ref["deriv"])
        want_v, want_d, want_vir = float(g[case["tag"] + "_value"]), g[case["tag"] + "_deriv"], g[case["tag"] + "_virial"]
        assert abs(ref["value"] - want_v) <= 1e-12 * max(1.0, abs(want_v)), case
        assert rel_err(deriv, want_d) <= 1e-12, case
        assert rel_err(ref["virial"], want_vir) <= 1e-12, case


def test_dhenergy_constructor_constants():
    """k of the constructor (DHEnergy.cpp:120) in the oracle and in the library's host code"""
    from oracle import oracle as O
    from plumed2_b200 import capi
    k = np.sqrt(0.1 / (80.0 * 300.0)) * 502.903741125
    assert abs(O.make_dhenergy(0.1, 300.0, 80.0).beta - k) < 1e-12
    s = capi.pairing_dhenergy(0.1, 300.0, 80.0)
    assert s.type == 32 and abs(s.beta - k) < 1e-12 and abs(s.lambda_ - 138.935458111 / 80.0) < 1e-12
    with pytest.raises(capi.B200CoordError):
        capi.pairing_dhenergy(0.1, 300.0, 0.0)


@pytest.mark.gpu
def test_gpu_dhenergy_matches_reference_goldens():
    import plumed2_b200 as P
    g, cases = _gold()
    for case in cases:
        pos, box = _frame(g, case["tag"])
        c = P.Coordination.from_input(case["line"])
        c.set_charges(g["q"])
        c.prepare(0)
        c.calculate(pos, box)
        n = pos.shape[0]
        deriv = np.zeros((n, 3))
        np.add.at(deriv, c.atoms, c.derivatives)
        want_v, want_d, want_vir = float(g[case["tag"] + "_value"]), g[case["tag"] + "_deriv"], g[case["tag"] + "_virial"]
        assert abs(c.value - want_v) <= 1e-10 * max(1.0, abs(want_v)), (case, c.value, want_v)
        assert rel_err(deriv, want_d) <= 1e-10, case
        assert rel_err(c.virial, want_vir) <= 1e-10, case
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["", "NLIST NL_CUTOFF=1.2 NL_STRIDE=3", "NLISTCELLS NL_CUTOFF=1.2 NL_STRIDE=3"])
def test_gpu_dhenergy_frozen_list_and_charge_update(mode):
    import plumed2_b200 as P
    n = 3000
    pos0, box = water_box(n, 100.0, seed=8, triclinic=True)
    rng = np.random.default_rng(2)
    q = rng.standard_normal(n)
    line = "c: DHENERGY GROUPA=1-1200 GROUPB=1201-%d I=0.15 EPSILON=78.0 TEMP=298 %s" % (n, mode)
    c = P.Coordination.from_input(line)
    with pytest.raises(P.capi.B200CoordError):  # charges are compulsory
        c.prepare(0)
        c.calculate(pos0, box)
    c.set_charges(q)
    pos, list_pos = pos0.copy(), None
    for step in range(5):
        pos = pos + 0.01 * rng.standard_normal(pos.shape)
        if step == 3:
            q = q * 0.5 + 0.1
            c.set_charges(q)
        if c.prepare(step) or list_pos is None:
            list_pos = pos.copy()
        c.calculate(pos, box)
        ref = oracle_from_line(line, pos, box, list_positions=list_pos if mode else None, charges=q, nthreads=8)
        assert abs(c.value - ref["value"]) <= 1e-10 * abs(ref["value"]), (mode, step)
        assert rel_err(c.derivatives, ref["deriv"]) <= 1e-10 and rel_err(c.virial, ref["virial"]) <= 1e-10, (mode, step)
    c.close()

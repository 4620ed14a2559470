"""CPU tests: the C oracle (oracle/coord_oracle.c) against
  (a) the reference's own known-answer files (tests/golden/ref_regtest_kats.json, parsed from
      /root/reference/regtest by oracle/gen_golden.py), and
  (b) full-precision outputs of the real reference (tests/golden/ref_outputs.npz).
This is what pins the oracle; the GPU parity tests then compare the CUDA path with the oracle."""
import json
import os

import numpy as np
import pytest

from helpers import oracle_from_line, scatter_to_system, sort_pairs
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "ref_regtest_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_outputs.npz"))


# ------------------------------------------------------------------ switching functions
def test_switch_regtest_tables(kats):
    """regtest/basic/rt-make-switch/out_*.reference: value and derivative at 10 points, 6 decimals"""
    for name, entry in kats["switch_tables"].items():
        sw = O.make_switch(entry["definition"])
        for point, val, der in entry["rows"]:
            v, d = O.switch_calculate(sw, point)
            v2, d2 = O.switch_calculate_sqr(sw, point * point)
            assert abs(v - val) < 6e-7 and abs(d - der) < 6e-7, (name, point, v, val, d, der)
            assert abs(v2 - val) < 6e-7 and abs(d2 - der) < 6e-7, (name, "sqr", point)


def test_switch_tables_bit_exact(gold):
    defs = json.loads(str(gold["switch_defs_json"]))
    r, tab = gold["switch_r"], gold["switch_table"]
    for i, d in enumerate(defs):
        sw = O.make_switch(d)
        for j, x in enumerate(r):
            got = (*O.switch_calculate(sw, x), *O.switch_calculate_sqr(sw, x * x))
            np.testing.assert_array_equal(np.array(got), tab[i, j], err_msg="%s r=%r" % (d, x))


def test_switch_keyword_form_bit_exact(gold):
    """R_0/NN/MM/D_0 keywords: automatic D_MAX and stretch (SwitchingFunction.cpp:1176-1184)"""
    r, tab = gold["switch_r"], gold["switch_kw_table"]
    for i, (nn, mm, r0, d0) in enumerate(gold["switch_kw"]):
        sw = O.make_switch(nn=int(nn), mm=int(mm), r0=r0, d0=d0)
        for j, x in enumerate(r):
            got = (*O.switch_calculate(sw, x), *O.switch_calculate_sqr(sw, x * x))
            np.testing.assert_array_equal(np.array(got), tab[i, j])


def test_switch_errors():
    for bad in ["", "FOO R_0=1", "RATIONAL", "RATIONAL R_0=1 BAR=2", "SMAP R_0=1", "RATIONAL R_0=x"]:
        with pytest.raises(ValueError):
            O.make_switch(bad)


# ------------------------------------------------------------------ Pbc / LatticeReduction / Tools::pbc
def test_tools_pbc_bit_exact(gold):
    for x, y in zip(gold["tools_pbc_x"], gold["tools_pbc_y"]):
        assert O.tools_pbc(x) == y


def test_pbc_distance_and_reduction_bit_exact(gold):
    for b, vec, dist, red in zip(gold["pbc_boxes"], gold["pbc_vec"], gold["pbc_dist"], gold["pbc_reduced"]):
        p = O.make_pbc(b)
        np.testing.assert_array_equal(O.lattice_reduce(b), red)
        zero = np.zeros(3)
        for v, d in zip(vec, dist):
            np.testing.assert_array_equal(O.pbc_distance(p, zero, v), d)


def test_pbc_distance_is_minimum_image(gold):
    """regtest/basic/rt-make-1 logic: Pbc::distance agrees with the brute-force full search"""
    for b, vec in zip(gold["pbc_boxes"], gold["pbc_vec"]):
        p = O.make_pbc(b)
        for v in vec[:40]:
            d = O.pbc_distance(p, np.zeros(3), v)
            f = O.pbc_full_search(p, v)
            assert abs(np.dot(d, d) - np.dot(f, f)) < 1e-9


# ------------------------------------------------------------------ LinkCells
def test_linkcells_regtest_ncells(kats):
    """regtest/tools/rt-make-CellLists/outputIndexes.reference: cutoff 1.5 in three boxes"""
    boxes = [np.diag([10.0, 10, 10]), np.array([[10.0, 10, 0], [0, 10, 0], [0, 0, 10]]),
             np.array([[10.0, 5, 3], [5, 10, 2], [3, 2, 10]])]
    for b, want in zip(boxes, kats["linkcells_ncells"]):
        lc = O.linkcells(1.5, np.zeros((1, 3)), O.make_pbc(b))
        assert int(np.prod(list(lc.ncells))) == want


def test_linkcells_cells_and_stencils(gold):
    for i in range(6):
        b, pts, cut = gold["lc%d_box" % i], gold["lc%d_pts" % i], float(gold["lc%d_cut" % i])
        lc = O.linkcells(cut, pts, O.make_pbc(b))
        assert list(lc.ncells) == list(gold["lc%d_ncells" % i])
        got = np.array([O.linkcells_find_cell(lc, p) for p in pts], dtype=np.uint32)
        np.testing.assert_array_equal(got, gold["lc%d_cell" % i])
        for row in gold["lc%d_stencil" % i]:
            use_pbc, cx, cy, cz, m = (int(v) for v in row[:5])
            req = O.linkcells_required(lc, [cx, cy, cz], use_pbc)
            np.testing.assert_array_equal(req, row[5:5 + m].astype(np.uint32))


# ------------------------------------------------------------------ NeighborList
def _sc_lattice(n):
    """AtomDistribution 'sc' (src/tools/AtomDistribution.cpp:205-246): integer lattice, x fastest"""
    rmax = int(np.ceil(round(n ** (1 / 3), 9)))
    while rmax ** 3 < n:
        rmax += 1
    pts = [(i, j, k) for k in range(rmax) for j in range(rmax) for i in range(rmax)][:n]
    return np.array(pts, dtype=np.float64), np.diag([float(rmax)] * 3)


def _neighbours_by_atom(pairs, nslots, index_of_slot):
    out = {int(index_of_slot[s]): [] for s in range(nslots)}
    for a, b in pairs:
        out[int(index_of_slot[a])].append(int(index_of_slot[b]))
        out[int(index_of_slot[b])].append(int(index_of_slot[a]))
    return {k: sorted(v) for k, v in out.items()}


@pytest.mark.parametrize("do_pbc", [False, True])
def test_neighbourlist_regtest_golden_sets(kats, do_pbc):
    """regtest/tools/rt-Neigbourlist/unitTest.reference (single / two lists / pairs, cutoff 1.999 a)"""
    tag = "on" if do_pbc else "off"
    # single list, 125 atoms
    pos, box = _sc_lattice(125)
    cutoff = (box[0, 0] / 5) * 1.999
    pbc = O.make_pbc(box)
    nl = O.NeighborList(O.NL_SINGLELIST, 125, 0, do_pbc=do_pbc, cutoff=cutoff, stride=1)
    nl.update(pbc, pos)
    got = _neighbours_by_atom(nl.pairs(), 125, np.arange(125))
    want = kats["neighbour_sets"]["Single list|" + tag]
    assert {str(k): v for k, v in got.items()} == want
    # the cell-accelerated oracle variant finds the same set
    nl.update(pbc, pos, fast=True)
    assert {str(k): v for k, v in _neighbours_by_atom(nl.pairs(), 125, np.arange(125)).items()} == want
    # two lists / pairs: 124 atoms, A = even indices, B = odd indices
    pos, box = _sc_lattice(124)
    ia, ib = np.arange(0, 124, 2), np.arange(1, 124, 2)
    idx = np.concatenate([ia, ib])
    for style, label in ((O.NL_TWOLIST, "Two lists"), (O.NL_PAIR, "List of pairs")):
        nl = O.NeighborList(style, 62, 62, do_pbc=do_pbc, cutoff=cutoff, stride=1)
        nl.update(O.make_pbc(box), pos[idx])
        got = _neighbours_by_atom(nl.pairs(), 124, idx)
        assert {str(k): v for k, v in got.items()} == kats["neighbour_sets"][label + "|" + tag]


@pytest.mark.parametrize("do_pbc", [False, True])
def test_no_neighbourlist_regtest(kats, do_pbc):
    """regtest/tools/rt-Neigbourlist/testNoNL.reference: all pairs in getIndexPair order"""
    nl = O.NeighborList(O.NL_SINGLELIST, 27, 0, do_pbc=do_pbc)
    prs = []
    import ctypes as C
    for k in range(nl.size()):
        i0, i1 = C.c_uint(), C.c_uint()
        O.lib().orc_nl_index_pair(nl.h, k, C.byref(i0), C.byref(i1))
        prs.append((i0.value, i1.value))
    assert prs[:3] == [(0, 1), (0, 2), (0, 3)] and prs[-1] == (25, 26)
    got = _neighbours_by_atom(prs, 27, np.arange(27))
    assert {str(k): v for k, v in got.items()} == kats["neighbour_sets"]["NoNL|" + ("on" if do_pbc else "off")]


def test_neighbourlist_pair_sets_match_reference(gold):
    for case in json.loads(str(gold["nl_cases_json"])):
        pos = gold[case["pos"]]
        box = np.zeros((3, 3)) if case["box"] == "zero" else gold[case["box"]]
        nl = O.NeighborList(case["style"], case["n0"], case["n1"], do_pbc=bool(case["do_pbc"]),
                            use_cells=bool(case["cells"]), cutoff=case["cutoff"], stride=2)
        pbc = O.make_pbc(box)
        nl.update(pbc, pos)
        want = gold[case["tag"] + "_pairs"]
        np.testing.assert_array_equal(sort_pairs(nl.pairs()), want, err_msg=case["tag"])
        if not case["cells"]:
            nl.update(pbc, pos, fast=True)
            np.testing.assert_array_equal(sort_pairs(nl.pairs()), want, err_msg=case["tag"] + " (fast)")


def test_prepare_schedule():
    """NeighborList::prepare (NeighborList.cpp:433-456): rebuild on first step and every NL_STRIDE"""
    nl = O.NeighborList(O.NL_SINGLELIST, 10, 0, cutoff=1.0, stride=4)
    seq = [nl.prepare(s) for s in range(3, 13)]
    assert seq == [True, True, False, False, False, True, False, False, False, True]
    nl1 = O.NeighborList(O.NL_SINGLELIST, 10, 0, cutoff=1.0, stride=1)
    assert all(nl1.prepare(s) for s in range(5))
    nl2 = O.NeighborList(O.NL_SINGLELIST, 10, 0, cutoff=1.0, stride=5)
    assert nl2.prepare(7) is True and nl2.prepare(8, exchange_step=True) is False and nl2.prepare(9) is True


# ------------------------------------------------------------------ COORDINATION
def test_coordination_regtest_values(kats, gold):
    """regtest/basic/rt42 (171.1815 ...), rt42c (half of it), rt42-cells, printed with %8.4f"""
    pos, box = gold["rt42_pos"], gold["rt42_box"]
    v42 = kats["coordination_regtests"]["rt42"]["values"]
    for fn, line in (("check_c", "c: COORDINATION GROUPA=1-108 GROUPB=1-108 R_0=1"),
                     ("check_e", "c: COORDINATION GROUPA=1-108 GROUPB=1-108 R_0=1 NN=7 D_0=0.1"),
                     ("check_g", "c: COORDINATION GROUPA=1-108 GROUPB=1-108 SWITCH={EXP R_0=1}")):
        r = oracle_from_line(line, pos, box)
        assert abs(r["value"] - v42[fn][0][1]) < 5.1e-5, (fn, r["value"], v42[fn][0][1])
    half = oracle_from_line("c: COORDINATION GROUPA=1-108 R_0=1", pos, box)["value"]
    assert abs(half - 171.1815 / 2) < 1e-4


def test_coordination_matches_reference_outputs(gold):
    """value, 3N derivatives and virial of the real reference (through plumed_cmd) for 30 input lines"""
    cases = json.loads(str(gold["cases_json"]))
    assert len(cases) >= 30
    for case in cases:
        tag, line = case["tag"], case["line"]
        if tag.startswith("rt42"):
            pos, box = gold["rt42_pos"], gold["rt42_box"]
        elif tag.startswith("ortho"):
            pos, box = gold["ortho_pos"], gold["ortho_box"]
        elif tag.startswith("tri"):
            pos, box = gold["tri_pos"], gold["tri_box"]
        else:
            pos, box = gold["tri_pos"], None
        r = oracle_from_line(line, pos, box)
        want_v, want_d, want_vir = float(gold[tag + "_value"]), gold[tag + "_deriv"], gold[tag + "_virial"]
        assert abs(r["value"] - want_v) <= 1e-13 * max(1.0, abs(want_v)), (tag, r["value"], want_v)
        got_d = scatter_to_system(pos.shape[0], r["atoms"], r["deriv"])
        scale = max(np.abs(want_d).max(), 1e-300)
        assert np.abs(got_d - want_d).max() <= 1e-12 * scale, tag
        assert np.abs(r["virial"] - want_vir).max() <= 1e-12 * max(np.abs(want_vir).max(), 1e-300), tag


def test_rank_split_sums_to_full(gold):
    """the MPI stride split of CoordinationBase.cpp:152-170: partial results of all ranks add up"""
    pos, box = gold["ortho_pos"], gold["ortho_box"]
    n = pos.shape[0]
    sw = O.make_switch("RATIONAL R_0=0.3 D_MAX=0.8")
    pbc = O.make_pbc(box)
    nl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=0.9, stride=2)
    nl.update(pbc, pos)
    full = O.coordination(nl, pbc, True, sw, pos)
    acc_v, acc_d, acc_w = 0.0, np.zeros((n, 3)), np.zeros((3, 3))
    for r in range(3):
        v, d, w, _ = O.coordination(nl, pbc, True, sw, pos, rank=r, nranks=3)
        acc_v, acc_d, acc_w = acc_v + v, acc_d + d, acc_w + w
    assert abs(acc_v - full[0]) < 1e-10 * abs(full[0])
    np.testing.assert_allclose(acc_d, full[1], rtol=0, atol=1e-10 * np.abs(full[1]).max())
    np.testing.assert_allclose(acc_w, full[2], rtol=0, atol=1e-10 * np.abs(full[2]).max())
    # OpenMP path gives the same numbers to rounding
    v4, d4, w4, _ = O.coordination(nl, pbc, True, sw, pos, nthreads=4)
    assert abs(v4 - full[0]) < 1e-12 * abs(full[0])


# ------------------------------------------------------------------ more of the reference's own fixtures
def test_lattice_reduction_regtest(kats):
    """regtest/tools/rt-make-lattice-reduction/output.reference (testReduceFast)"""
    kat = kats["lattice_reduction"]
    np.testing.assert_array_equal(O.lattice_reduce(kat["input"]), np.array(kat["reduceFast"]))


def _rt20_expected(test, nframes):
    """per frame: (c, cs) values and their 15 derivatives (6 atomic + 9 box) as printed by the reference"""
    col = np.array(test["colvar"])
    der = np.array(test["deriv"]).reshape(nframes, 15, 4)
    return col[:, 1:3], der[:, :, 2:4]


def test_rt20_switch_regtests(kats):
    """regtest/basic/rt20-switch-*: 2-atom COORDINATION along switchtraj.xyz for every analytic switch,
    stretched and NOSTRETCH: value (%f) and DUMPDERIVATIVES (%8.4f) incl. the virial"""
    rt = kats["rt20_switch"]
    frames = rt["frames"]
    assert len(rt["tests"]) == 12
    for name, test in rt["tests"].items():
        vals, ders = _rt20_expected(test, len(frames))
        for fi, fr in enumerate(frames):
            pos = np.array(fr["pos"])
            box = np.diag(fr["box"])
            for ci, lab in enumerate(("c", "cs")):
                r = oracle_from_line(test["lines"][lab], pos, box)
                assert abs(r["value"] - vals[fi, ci]) < 6e-7, (name, lab, fi)
                got = np.concatenate([r["deriv"].ravel(), r["virial"].ravel()])
                assert np.abs(got - ders[fi, :, ci]).max() < 5.1e-5, (name, lab, fi)

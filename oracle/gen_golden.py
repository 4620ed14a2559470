"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/ from the reference.  Run where /root/reference and
oracle/_ref exist:   python oracle/gen_golden.py

Two kinds of fixtures are written:

  tests/golden/ref_regtest_kats.json
      the reference's OWN known-answer files for this path, parsed (numbers only, no source code):
        regtest/basic/rt-make-switch/out_*.reference          (12 switch definitions x stretch on/off x 10 points)
        regtest/tools/rt-Neigbourlist/unitTest.reference      (golden neighbour sets, 5x5x5 simple cubic)
        regtest/tools/rt-Neigbourlist/testNoNL.reference
        regtest/tools/rt-make-CellLists/outputIndexes.reference ("Ncells:" lines)
        regtest/basic/rt42/check_*.reference, rt42c, rt42-cells (COORDINATION values) + frame 0 of their trajectory
  tests/golden/ref_outputs.npz
      full-precision outputs of the REAL reference (oracle/_ref, driven through plumed_cmd and through the
      tool classes) on seeded synthetic inputs: value, 3N derivatives, virial, neighbour-list pair sets,
      switching-function tables, minimum-image tables, cell indices.  Inputs are stored too, so the tests
      need neither the reference nor the generator at run time.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:] = [q for q in sys.path if os.path.abspath(q or '.') != HERE]
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import refplumed as R  # noqa: E402

SWITCH_DEFS = {  # regtest/basic/rt-make-switch/main.cpp:96-148
    "cosinus": "COSINUS R_0=2.6",
    "exp": "EXP R_0=0.8 D_0=0.5 D_MAX=2.6",
    "fastgaussian": "GAUSSIAN R_0=1.0 D_0=0.0 D_MAX=2.6",
    "gaussian": "GAUSSIAN R_0=1.0 D_0=0.3 D_MAX=2.6",
    "fastrational": "RATIONAL R_0=1.3 NN=6 MM=10 D_MAX=2.6",
    "fastrational_NNeq2MM": "RATIONAL R_0=1.3 D_MAX=2.6",
    "rational": "RATIONAL R_0=1.3 NN=5 MM=11 D_MAX=2.6",
    "rational_NNeq2MM": "RATIONAL R_0=1.3 NN=5 D_MAX=2.6",
    "q": "Q R_0=1.0 D_0=0.3 BETA=5.0 LAMBDA=1.0 REF=1.3 D_MAX=2.6",
    "tanh": "TANH R_0=1.3 D_MAX=2.6",
    "smap": "SMAP R_0=1.3 A=3 B=2 D_MAX=2.6",
    "cubic": "CUBIC D_MAX=2.6 D_0=0.6",
}

EXTRA_SWITCHES = [
    "RATIONAL R_0=0.3 NN=6 MM=12", "RATIONAL R_0=0.3 D_MAX=0.8", "RATIONAL R_0=0.3 NN=8 MM=16 D_MAX=0.9",
    "RATIONAL R_0=0.3 NN=12 D_MAX=0.9", "RATIONAL R_0=0.3 NN=2 D_MAX=0.9", "RATIONAL R_0=0.3 NN=4 D_MAX=0.9",
    "RATIONAL R_0=0.3 NN=10 D_MAX=0.9", "RATIONAL R_0=0.3 NN=14 MM=28 D_MAX=0.9",
    "RATIONAL R_0=0.25 D_0=0.05 NN=6 MM=12 D_MAX=0.9", "RATIONAL R_0=0.3 NN=4 MM=10 D_MAX=0.9 NOSTRETCH",
    "RATIONAL R_0=0.3 NN=3 MM=7", "EXP R_0=0.2 D_MAX=0.9", "EXP R_0=0.2 D_0=0.1", "GAUSSIAN R_0=0.2 D_MAX=0.9",
    "GAUSSIAN R_0=1.0 D_MAX=0.9", "SMAP R_0=0.3 A=4 B=3 D_MAX=0.9", "CUBIC D_0=0.1 D_MAX=0.8",
    "TANH R_0=0.3 D_MAX=0.9", "COSINUS R_0=0.5 D_0=0.2", "Q R_0=1.0 D_0=0.1 BETA=30.0 LAMBDA=1.5 REF=0.3 D_MAX=0.9",
]


def read_xyz_frame(fn, frame=0):
    with open(fn) as f:
        for _ in range(frame + 1):
            n = int(f.readline())
            b = [float(x) for x in f.readline().split()]
            box = np.diag(b) if len(b) == 3 else np.array(b).reshape(3, 3)
            pos = np.array([[float(x) for x in f.readline().split()[1:4]] for _ in range(n)])
    return pos, box


def parse_kats():
    kats = {}
    # --- switching function tables
    sw = {}
    d = os.path.join(REF, "regtest/basic/rt-make-switch")
    for name, definition in SWITCH_DEFS.items():
        for suffix, extra in (("", ""), ("_nostretch", " NOSTRETCH")):
            rows = []
            with open(os.path.join(d, "out_%s%s.reference" % (name, suffix))) as f:
                next(f)
                for line in f:
                    left, right = line.split(":")
                    vals = right.split()
                    rows.append([float(left), float(vals[0]), float(vals[1])])
            sw[name + suffix] = {"definition": definition + extra, "rows": rows}
    kats["switch_tables"] = sw
    # --- golden neighbour sets
    nl = {}
    for fn in ("unitTest.reference", "testNoNL.reference"):
        with open(os.path.join(REF, "regtest/tools/rt-Neigbourlist", fn)) as f:
            for line in f:
                m = re.match(r"\[(.*), pbc (on|off)\] atom (\d+):(.*)", line)
                key = "%s|%s" % (m.group(1), m.group(2))
                nl.setdefault(key, {})[m.group(3)] = [int(x) for x in m.group(4).split()]
    kats["neighbour_sets"] = nl
    # --- cell counts
    with open(os.path.join(REF, "regtest/tools/rt-make-CellLists/outputIndexes.reference")) as f:
        kats["linkcells_ncells"] = [int(l.split()[1]) for l in f if l.startswith("Ncells:")]
    # --- COORDINATION regtests (values printed with %8.4f)
    coord = {}
    for test, files in (("rt42", ["check_c", "check_e", "check_g"]), ("rt42c", None), ("rt42-cells", None)):
        d = os.path.join(REF, "regtest/basic", test)
        entry = {"plumed_dat": open(os.path.join(d, "plumed.dat")).read()}
        vals = {}
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".reference") and not fn.startswith("ff") and not fn.startswith("forces"):
                with open(os.path.join(d, fn)) as f:
                    lines = [l for l in f if not l.startswith("#")]
                if lines:
                    try:
                        vals[fn[:-10]] = [[float(x) for x in l.split()] for l in lines[:4]]
                    except ValueError:
                        pass
        entry["values"] = vals
        coord[test] = entry
    kats["coordination_regtests"] = coord
    # --- regtest/basic/rt20-switch-*: 2-atom COORDINATION along switchtraj.xyz, every analytic switch type,
    #     with and without NOSTRETCH: COLVAR (6 decimals) and DUMPDERIVATIVES (%8.4f: 6 atom + 9 box derivatives)
    frames = []
    with open(os.path.join(REF, "regtest/trajectories/switchtraj.xyz")) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines) and lines[i].strip():
        n = int(lines[i]); b = [float(x) for x in lines[i + 1].split()]
        at = [[float(x) for x in lines[i + 2 + a].split()[1:4]] for a in range(n)]
        frames.append({"box": b, "pos": at}); i += 2 + n
    rt20 = {"frames": frames, "tests": {}}
    base = os.path.join(REF, "regtest/basic")
    for d in sorted(os.listdir(base)):
        if not d.startswith("rt20-switch-") or d.endswith("lepton"):
            continue
        dat = open(os.path.join(base, d, "plumed.dat")).read()
        lines_c = {}
        for ln in dat.split("\n"):
            m = re.match(r"\s*(c|cs):\s*(COORDINATION.*)", ln)
            if m:
                lines_c[m.group(1)] = m.group(1) + ": " + m.group(2).strip()
        colvar = [[float(x) for x in l.split()] for l in open(os.path.join(base, d, "COLVAR.reference")) if not l.startswith("#")]
        deriv = [[float(x) for x in l.split()] for l in open(os.path.join(base, d, "deriv.reference")) if not l.startswith("#")]
        rt20["tests"][d] = {"lines": lines_c, "colvar": colvar, "deriv": deriv}
    kats["rt20_switch"] = rt20
    # --- regtest/tools/rt-make-lattice-reduction/output.reference
    with open(os.path.join(REF, "regtest/tools/rt-make-lattice-reduction/output.reference")) as f:
        L = f.read().split("\n")
    k = L.index("testReduceFast")
    kats["lattice_reduction"] = {"input": [1.0, 2.0, 3.0, 5.0, 4.0, 3.0, 10.0, 8.0, 2.0],
                                 "reduceFast": [[float(x) for x in L[k + 1 + r].split()] for r in range(3)]}
    return kats


def gen_outputs():
    out = {}
    rng = np.random.default_rng(20261017)
    # --- rt42 frame 0 (108 Ar atoms, triclinic box) : input of the reference's own regtest
    pos42, box42 = read_xyz_frame(os.path.join(REF, "regtest/basic/rt42/trajectory.xyz"))
    out["rt42_pos"], out["rt42_box"] = pos42, box42
    cases = []

    def run_case(tag, natoms, line, pos, box):
        p = R.Plumed(natoms, [line, "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"], watch=["c"])
        r = p.calc(0, pos, box)
        out[tag + "_value"] = np.array(p.value("c"))
        out[tag + "_deriv"] = -r["forces"]
        out[tag + "_virial"] = -r["virial"]
        p.close()
        cases.append({"tag": tag, "line": line})

    # every GPU-capable switch on the rt42 frame, TwoList with identical groups (self pairs!) as in rt42
    sw_lines = ["R_0=1", "R_0=1 NN=7 D_0=0.1", "SWITCH={EXP R_0=1}", "SWITCH={RATIONAL R_0=1.0 D_MAX=2.5 NN=4 MM=10}",
                "SWITCH={GAUSSIAN R_0=1 D_MAX=2.5}", "SWITCH={GAUSSIAN R_0=0.8 D_0=0.2 D_MAX=2.5}",
                "SWITCH={SMAP R_0=1.0 A=3 B=2 D_MAX=2.5}", "SWITCH={CUBIC D_0=0.5 D_MAX=2.5}",
                "SWITCH={TANH R_0=1.0 D_MAX=2.5}", "SWITCH={COSINUS R_0=1.5 D_0=0.5}",
                "SWITCH={Q R_0=1.0 D_0=0.3 BETA=5.0 LAMBDA=1.0 REF=1.3 D_MAX=2.6}",
                "SWITCH={RATIONAL R_0=1.0 NN=5 MM=11 D_MAX=2.5}", "SWITCH={RATIONAL R_0=1.0 NN=8 D_MAX=2.5}",
                "SWITCH={RATIONAL R_0=1.0 NN=14 MM=28 D_MAX=2.5}", "SWITCH={RATIONAL R_0=1.0 D_MAX=2.5 NOSTRETCH}"]
    for i, s in enumerate(sw_lines):
        run_case("rt42_two_%02d" % i, 108, "c: COORDINATION GROUPA=1-108 GROUPB=1-108 " + s, pos42, box42)
    run_case("rt42_single", 108, "c: COORDINATION GROUPA=1-108 R_0=1", pos42, box42)
    run_case("rt42_nlist", 108, "c: COORDINATION GROUPA=1-108 SWITCH={RATIONAL R_0=1 D_MAX=1.5} NLIST NL_CUTOFF=2.0 NL_STRIDE=4", pos42, box42)
    run_case("rt42_cells", 108, "c: COORDINATION GROUPA=1-108 SWITCH={RATIONAL R_0=1 D_MAX=1.5} NLISTCELLS NL_CUTOFF=2.0 NL_STRIDE=4", pos42, box42)
    run_case("rt42_nopbc", 108, "c: COORDINATION GROUPA=1-108 R_0=1 NOPBC", pos42, box42)
    run_case("rt42_pair", 108, "c: COORDINATION GROUPA=1-54 GROUPB=55-108 R_0=1 PAIR", pos42, box42)
    run_case("rt42_groups", 108, "c: COORDINATION GROUPA=1-30 GROUPB=20-108 SWITCH={EXP R_0=1 D_MAX=2.9}", pos42, box42)

    # --- synthetic water-like boxes (100 atoms/nm^3)
    n = 600
    L = (n / 100.0) ** (1 / 3)
    box_o = np.diag([L, L * 1.1, L * 0.95])
    pos_o = rng.random((n, 3)) @ box_o + 3.0 * rng.standard_normal((n, 3))  # unwrapped, far outside the cell too
    out["ortho_pos"], out["ortho_box"] = pos_o, box_o
    run_case("ortho_single", n, "c: COORDINATION GROUPA=1-600 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8}", pos_o, box_o)
    run_case("ortho_nlist", n, "c: COORDINATION GROUPA=1-600 SWITCH={RATIONAL R_0=0.3 D_MAX=0.5} NLIST NL_CUTOFF=0.6 NL_STRIDE=5", pos_o, box_o)
    run_case("ortho_cells", n, "c: COORDINATION GROUPA=1-600 SWITCH={RATIONAL R_0=0.3 D_MAX=0.5} NLISTCELLS NL_CUTOFF=0.6 NL_STRIDE=5", pos_o, box_o)
    run_case("ortho_two", n, "c: COORDINATION GROUPA=1-100 GROUPB=101-600 SWITCH={EXP R_0=0.2 D_MAX=0.9}", pos_o, box_o)
    box_t = L * np.array([[1.0, 0.0, 0.0], [0.2, 1.0, 0.0], [0.1, 0.3, 1.0]])
    pos_t = rng.random((n, 3)) @ box_t
    out["tri_pos"], out["tri_box"] = pos_t, box_t
    run_case("tri_two_cells", n, "c: COORDINATION GROUPA=1-100 GROUPB=101-600 SWITCH={EXP R_0=0.2 D_MAX=0.55} NLISTCELLS NL_CUTOFF=0.6 NL_STRIDE=1", pos_t, box_t)
    run_case("tri_two_nlist", n, "c: COORDINATION GROUPA=1-100 GROUPB=101-600 SWITCH={EXP R_0=0.2 D_MAX=0.55} NLIST NL_CUTOFF=0.6 NL_STRIDE=1", pos_t, box_t)
    run_case("tri_single", n, "c: COORDINATION GROUPA=1-600 SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.8}", pos_t, box_t)
    run_case("nobox_single", n, "c: COORDINATION GROUPA=1-600 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=0.9 NL_STRIDE=2", pos_t, None)
    run_case("nobox_cells", n, "c: COORDINATION GROUPA=1-600 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLISTCELLS NL_CUTOFF=0.9 NL_STRIDE=2", pos_t, None)
    out["cases_json"] = np.array(json.dumps(cases))

    # --- neighbour-list pair sets from the reference NeighborList class
    nl_cases = []
    for tag, style, n0, n1, do_pbc, cells, cut, pos, box in [
            ("nl_single_pbc", 2, n, 0, 1, 0, 0.6, pos_o, box_o), ("nl_single_nopbc", 2, n, 0, 0, 0, 0.6, pos_o, box_o),
            ("nl_single_cells", 2, n, 0, 1, 1, 0.6, pos_o, box_o), ("nl_single_cells_nopbc", 2, n, 0, 0, 1, 0.6, pos_o, box_o),
            ("nl_two_tri", 1, 100, 500, 1, 0, 0.6, pos_t, box_t), ("nl_two_tri_cells", 1, 100, 500, 1, 1, 0.6, pos_t, box_t),
            ("nl_pair_tri", 0, 300, 300, 1, 0, 0.9, pos_t, box_t), ("nl_single_nobox_cells", 2, n, 0, 1, 1, 0.9, pos_t, np.zeros((3, 3))),
            ("nl_single_tri_wide", 2, n, 0, 1, 0, 0.95, pos_t, box_t), ("nl_single_tri_cells_wide", 2, n, 0, 1, 1, 0.95, pos_t, box_t)]:
        nl = R.RefNeighborList(style, n0, n1, do_pbc, cells, cut, 2, box)
        nl.update(pos)
        pr = nl.pairs()
        out[tag + "_pairs"] = pr[np.lexsort((pr[:, 1], pr[:, 0]))].astype(np.uint32)
        nl_cases.append({"tag": tag, "style": style, "n0": n0, "n1": n1, "do_pbc": do_pbc, "cells": cells, "cutoff": cut,
                         "pos": "ortho_pos" if pos is pos_o else "tri_pos",
                         "box": "zero" if not np.any(box) else ("ortho_box" if box is box_o else "tri_box")})
    out["nl_cases_json"] = np.array(json.dumps(nl_cases))

    # --- switching function tables (calculate and calculateSqr)
    r = np.concatenate([np.linspace(0.0, 1.2, 61), [0.3, 0.30000000001, 0.29999999999, 0.8, 0.9]])
    sw_defs = [v for v in SWITCH_DEFS.values()] + [v + " NOSTRETCH" for v in SWITCH_DEFS.values()] + EXTRA_SWITCHES
    tab = np.zeros((len(sw_defs), len(r), 4))
    for i, d in enumerate(sw_defs):
        s = R.RefSwitch(d)
        for j, x in enumerate(r):
            tab[i, j, 0:2] = s.calculate(x)
            tab[i, j, 2:4] = s.calculate_sqr(x * x)
    out["switch_r"] = r
    out["switch_table"] = tab
    out["switch_defs_json"] = np.array(json.dumps(sw_defs))
    kw = [(6, 0, 0.3, 0.0), (8, 0, 0.25, 0.0), (6, 12, 0.3, 0.1), (4, 10, 0.3, 0.0), (7, 0, 1.0, 0.1), (5, 9, 0.4, 0.0)]
    tabk = np.zeros((len(kw), len(r), 4))
    for i, (nn, mm, r0, d0) in enumerate(kw):
        s = R.RefSwitch(nn=nn, mm=mm, r0=r0, d0=d0)
        for j, x in enumerate(r):
            tabk[i, j, 0:2] = s.calculate(x)
            tabk[i, j, 2:4] = s.calculate_sqr(x * x)
    out["switch_kw"] = np.array(kw, dtype=np.float64)
    out["switch_kw_table"] = tabk

    # --- minimum image tables: the six lattice families of regtest/basic/rt-make-1/main.cpp
    boxes = []
    for kind in range(6):
        for _ in range(4):
            b = np.zeros((3, 3))
            if kind == 0:
                b = np.diag(1 + 2 * rng.random(3))
            elif kind == 1:
                b = np.diag(1 + 2 * rng.random(3)); b[1, 0] = rng.random() - 0.5
            elif kind == 2:
                b = np.diag(1 + 2 * rng.random(3)); b[2, 0], b[2, 1] = rng.random(2) - 0.5
            elif kind == 3:
                b = np.diag(1 + 2 * rng.random(3)) + np.tril(rng.random((3, 3)) - 0.5, -1) * 2
            elif kind == 4:
                b = rng.random((3, 3)) * 2 - 1 + np.eye(3) * 2
            else:
                b = np.array([[6.0, -6, 0], [0, 6, -6], [-6, 6, 6]]) * (0.5 + rng.random())
            boxes.append(b)
    boxes = np.array(boxes)
    vec = (rng.random((len(boxes), 200, 3)) - 0.5) * 12.0
    dist = np.zeros_like(vec)
    reduced = np.zeros_like(boxes)
    for i, b in enumerate(boxes):
        p = R.RefPbc(b)
        dist[i] = p.distance(np.zeros_like(vec[i]), vec[i])
        reduced[i] = R.ref_lattice_reduce(b)
    out["pbc_boxes"], out["pbc_vec"], out["pbc_dist"], out["pbc_reduced"] = boxes, vec, dist, reduced
    xs = np.concatenate([rng.random(500) * 8 - 4, [0.5, -0.5, 1.5, -1.5, 0.49999999999999994, 2.5, -2.5, 0.0, 99.5, -99.5]])
    out["tools_pbc_x"] = xs
    out["tools_pbc_y"] = np.array([R.lib().ref_tools_pbc(float(x)) for x in xs])

    # --- LinkCells: cell of every atom + stencil of every cell, for the boxes of rt-make-CellLists
    lc_boxes = [np.diag([10.0, 10, 10]), np.array([[10.0, 10, 0], [0, 10, 0], [0, 0, 10]]),
                np.array([[10.0, 5, 3], [5, 10, 2], [3, 2, 10]]), box_o, box_t, np.zeros((3, 3))]
    for i, b in enumerate(lc_boxes):
        cut = 1.5 if i < 3 else 0.6
        pts = (rng.random((300, 3)) * 3 - 1) @ (b if np.any(b) else np.eye(3) * 4)
        lc = R.RefLinkCells(cut, pts, b)
        nc = lc.ncells()
        out["lc%d_box" % i], out["lc%d_pts" % i], out["lc%d_cut" % i] = b, pts, np.array(cut)
        out["lc%d_ncells" % i] = nc
        out["lc%d_cell" % i] = np.array([lc.find_cell(p) for p in pts], dtype=np.uint32)
        st = []
        for use_pbc in (1, 0):
            for cz in range(nc[2]):
                for cy in range(nc[1]):
                    for cx in range(nc[0]):
                        req = lc.required([cx, cy, cz], use_pbc)
                        st.append(np.concatenate([[use_pbc, cx, cy, cz, len(req)], req, -np.ones(27 - len(req))]))
        out["lc%d_stencil" % i] = np.array(st, dtype=np.int64)
    return out


def main():
    if not R.available() or not os.path.isdir(REF):
        sys.exit("needs /root/reference and oracle/_ref (make -C oracle ref)")
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "ref_regtest_kats.json"), "w") as f:
        json.dump(parse_kats(), f, indent=0, sort_keys=True)
    out = gen_outputs()
    np.savez_compressed(os.path.join(OUT, "ref_outputs.npz"), **out)
    for fn in os.listdir(OUT):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()

"""Shared test helpers: synthetic frames and the oracle-side evaluation of a COORDINATION input."""
import numpy as np

from oracle import oracle as O

SEED = 20261017


def water_box(n, density=100.0, seed=SEED, triclinic=False, jitter=None):
    """n uniform points in a periodic box at `density` atoms/nm^3 (cube|scale distribution).
    Returns positions (n,3) and box (3,3)."""
    rng = np.random.default_rng(seed)
    L = (n / density) ** (1.0 / 3.0)
    frac = rng.random((n, 3))
    if triclinic:
        box = L * np.array([[1.0, 0.0, 0.0], [0.2, 1.0, 0.0], [0.1, 0.3, 1.0]])
    else:
        box = np.diag([L, L, L])
    pos = frac @ box
    if jitter:
        pos = pos + jitter * rng.standard_normal(pos.shape)
    return pos, box


def oracle_eval(pos, box, style, n_a, n_b, switch, do_pbc=True, nl_mode="none", cutoff=1e30, stride=0,
                abs_index=None, nthreads=4, list_pos=None, fast_list=False, list_box=None, charges=None, types=None):
    """value, deriv, virial, pairs from the C oracle for one frame; list built on list_pos (default pos)"""
    pbc = O.make_pbc(np.zeros(9) if box is None else box)
    st = {"pair": O.NL_PAIR, "two": O.NL_TWOLIST, "single": O.NL_SINGLELIST}[style]
    use_cells = nl_mode == "cells"
    nl = O.NeighborList(st, n_a, n_b, do_pbc=do_pbc, use_cells=use_cells, cutoff=cutoff,
                        stride=(stride if nl_mode != "none" else 0))
    if nl_mode != "none":
        lpbc = pbc if list_box is None else O.make_pbc(list_box)  # the box of the step the list was built at
        nl.update(lpbc, pos if list_pos is None else list_pos, fast=(fast_list and nl_mode == "classic"))
    v, d, vir, npairs = O.coordination(nl, pbc, do_pbc, switch, pos, abs_index, nthreads=nthreads, charges=charges, types=types)
    pairs = nl.pairs() if nl_mode != "none" else None
    return v, d, vir, pairs, npairs


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def oracle_switch_from_kv(kv):
    """the switching function a COORDINATION input line selects (Coordination.cpp:128-157)"""
    if "SWITCH" in kv:
        return O.make_switch(kv["SWITCH"])
    if "D_MAX" in kv:  # additive top-level D_MAX (cudaCoord semantics)
        return O.make_switch("RATIONAL R_0=%s D_0=%s NN=%s MM=%s D_MAX=%s" %
                             (kv["R_0"], kv.get("D_0", "0.0"), kv.get("NN", "6"), kv.get("MM", "0"), kv["D_MAX"]))
    return O.make_switch(nn=int(kv.get("NN", 6)), mm=int(kv.get("MM", 0)), r0=float(kv["R_0"]), d0=float(kv.get("D_0", 0.0)))


def oracle_from_line(line, positions, box, list_positions=None, nthreads=1, fast_list=False, list_box=None,
                     charges=None, types=None):
    """evaluate a `c: COORDINATION ...` (or `c: DHENERGY ...`, with system `charges`) input line with the C oracle on
    full-system positions.  Returns dict(value, deriv (n,3) per requested atom, virial, pairs, atoms)"""
    from plumed2_b200.coordination import parse_atom_list, split_input_line
    _, action, kv, flags = split_input_line(line)
    assert action in ("COORDINATION", "DHENERGY", "GHBFIX")
    ga = parse_atom_list(kv["GROUPA"])
    gb = parse_atom_list(kv.get("GROUPB"))
    atoms = np.concatenate([ga, gb]).astype(np.uint32)
    style = "single" if gb.size == 0 else ("pair" if "PAIR" in flags else "two")
    mode = "classic" if "NLIST" in flags else ("cells" if "NLISTCELLS" in flags else "none")
    q, tt = None, None
    if action == "GHBFIX":  # types = (type per SYSTEM atom, ntypes, etas), e.g. from O.read_ghbfix_tables
        sw = O.make_ghbfix(float(kv["D_MAX"]), float(kv["D_0"]), float(kv["C"]))
        tt = (np.ascontiguousarray(np.asarray(types[0], dtype=np.uint32)[atoms]), types[1], types[2])
    elif action == "DHENERGY":  # keyword defaults of DHEnergy::registerKeywords, DHEnergy.cpp:77-80
        sw = O.make_dhenergy(float(kv.get("I", 1.0)), float(kv.get("TEMP", 300.0)), float(kv.get("EPSILON", 80.0)))
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64)[atoms])
    else:
        sw = oracle_switch_from_kv(kv)
    pos = np.ascontiguousarray(np.asarray(positions, dtype=np.float64)[atoms])
    lpos = None if list_positions is None else np.ascontiguousarray(np.asarray(list_positions, dtype=np.float64)[atoms])
    v, d, vir, pairs, npairs = oracle_eval(pos, box, style, int(ga.size), int(gb.size), sw, do_pbc="NOPBC" not in flags,
                                           nl_mode=mode, cutoff=float(kv.get("NL_CUTOFF", 1e30)),
                                           stride=int(kv.get("NL_STRIDE", 0)), abs_index=atoms, nthreads=nthreads,
                                           list_pos=lpos, fast_list=fast_list, list_box=list_box, charges=q, types=tt)
    return dict(value=v, deriv=d, virial=vir, pairs=pairs, atoms=atoms, npairs=npairs)


def sort_pairs(p):
    p = np.asarray(p)
    if p.shape[0] == 0:
        return p.reshape(0, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def scatter_to_system(natoms, atoms, deriv):
    """sum per-slot derivatives onto system atoms (what Colvar::apply does with a unit force)"""
    out = np.zeros((natoms, 3))
    np.add.at(out, atoms, deriv)
    return out

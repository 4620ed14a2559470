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
                abs_index=None, nthreads=4, list_pos=None, fast_list=False):
    """value, deriv, virial, pairs from the C oracle for one frame; list built on list_pos (default pos)"""
    pbc = O.make_pbc(np.zeros(9) if box is None else box)
    st = {"pair": O.NL_PAIR, "two": O.NL_TWOLIST, "single": O.NL_SINGLELIST}[style]
    use_cells = nl_mode == "cells"
    nl = O.NeighborList(st, n_a, n_b, do_pbc=do_pbc, use_cells=use_cells, cutoff=cutoff,
                        stride=(stride if nl_mode != "none" else 0))
    if nl_mode != "none":
        nl.update(pbc, pos if list_pos is None else list_pos, fast=(fast_list and nl_mode == "classic"))
    v, d, vir, npairs = O.coordination(nl, pbc, do_pbc, switch, pos, abs_index, nthreads=nthreads)
    pairs = nl.pairs() if nl_mode != "none" else None
    return v, d, vir, pairs, npairs


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale

"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle_coord.so (coord_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (plumed2_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_coord.so")

NL_PAIR, NL_TWOLIST, NL_SINGLELIST = 0, 1, 2
PBC_UNSET, PBC_ORTHO, PBC_GENERIC = 0, 1, 2
MAXSHIFT = 6

SW_NAMES = ["rationalfix12", "rationalfix10", "rationalfix8", "rationalfix6", "rationalfix4", "rationalfix2",
            "rational", "rationalFast", "rationalSimple", "rationalSimpleFast", "exponential", "gaussian",
            "fastgaussian", "smap", "cubic", "tanh", "cosinus", "nativeq", "lepton", "not_initialized"]


class Switch(C.Structure):
    _fields_ = [("type", C.c_int), ("d0", C.c_double), ("dmax", C.c_double), ("dmax_2", C.c_double),
                ("invr0", C.c_double), ("invr0_2", C.c_double), ("stretch", C.c_double), ("shift", C.c_double),
                ("nn", C.c_int), ("mm", C.c_int), ("preRes", C.c_double), ("preDfunc", C.c_double),
                ("preSecDev", C.c_double), ("nnf", C.c_int), ("mmf", C.c_int), ("preDfuncF", C.c_double),
                ("preSecDevF", C.c_double), ("a", C.c_int), ("b", C.c_int), ("c", C.c_double), ("d", C.c_double),
                ("beta", C.c_double), ("lambda_", C.c_double), ("ref", C.c_double)]


class Pbc(C.Structure):
    _fields_ = [("type", C.c_int), ("box", C.c_double * 9), ("invBox", C.c_double * 9),
                ("reduced", C.c_double * 9), ("invReduced", C.c_double * 9), ("nshift", C.c_int * 8),
                ("shifts", C.c_double * (8 * MAXSHIFT * 3))]


class LinkCells(C.Structure):
    _fields_ = [("nopbc", C.c_int), ("cutoff", C.c_double), ("origin", C.c_double * 3), ("mypbc", Pbc),
                ("ncells", C.c_uint * 3), ("nstride", C.c_uint * 3)]


def build():
    """(re)compile the C restatement if missing or stale"""
    src = os.path.join(_HERE, "coord_oracle.c")
    hdr = os.path.join(_HERE, "coord_oracle.h")
    if (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        up = C.POINTER(C.c_uint)
        L.orc_tools_pbc.restype = C.c_double
        L.orc_tools_pbc.argtypes = [C.c_double]
        L.orc_switch_set.restype = C.c_int
        L.orc_switch_set.argtypes = [C.POINTER(Switch), C.c_char_p, C.c_char_p, C.c_int]
        L.orc_switch_set_rational.argtypes = [C.POINTER(Switch), C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_switch_calculate.restype = C.c_double
        L.orc_switch_calculate.argtypes = [C.POINTER(Switch), C.c_double, dp]
        L.orc_switch_calculate_sqr.restype = C.c_double
        L.orc_switch_calculate_sqr.argtypes = [C.POINTER(Switch), C.c_double, dp]
        L.orc_lattice_reduce.argtypes = [dp]
        L.orc_pbc_set_box.argtypes = [C.POINTER(Pbc), dp]
        L.orc_pbc_distance.argtypes = [C.POINTER(Pbc), dp, dp, dp]
        L.orc_pbc_full_search.argtypes = [C.POINTER(Pbc), dp]
        L.orc_linkcells_setup.argtypes = [C.POINTER(LinkCells), C.c_double, dp, C.c_size_t, C.POINTER(Pbc)]
        L.orc_linkcells_find_cell.restype = C.c_uint
        L.orc_linkcells_find_cell.argtypes = [C.POINTER(LinkCells), dp]
        L.orc_linkcells_required.restype = C.c_uint
        L.orc_linkcells_required.argtypes = [C.POINTER(LinkCells), up, C.c_int, up]
        L.orc_nl_create.restype = C.c_void_p
        L.orc_nl_create.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_double, C.c_uint]
        L.orc_nl_free.argtypes = [C.c_void_p]
        L.orc_nl_update.argtypes = [C.c_void_p, C.POINTER(Pbc), dp]
        L.orc_nl_update_classic_cells.argtypes = [C.c_void_p, C.POINTER(Pbc), dp]
        L.orc_nl_size.restype = C.c_size_t
        L.orc_nl_size.argtypes = [C.c_void_p]
        L.orc_nl_pairs.restype = up
        L.orc_nl_pairs.argtypes = [C.c_void_p]
        L.orc_nl_prepare.argtypes = [C.c_void_p, C.c_long, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_coordination_calculate.restype = C.c_size_t
        L.orc_coordination_calculate.argtypes = [C.c_void_p, C.POINTER(Pbc), C.c_int, C.POINTER(Switch), dp, up,
                                                 C.c_size_t, C.c_uint, C.c_uint, C.c_int, dp, dp, dp]
        L.orc_dhenergy_setup.argtypes = [C.POINTER(Switch), C.c_double, C.c_double, C.c_double]
        L.orc_dhenergy_calculate.restype = C.c_size_t
        L.orc_dhenergy_calculate.argtypes = [C.c_void_p, C.POINTER(Pbc), C.c_int, C.POINTER(Switch), dp, up, dp,
                                             C.c_size_t, C.c_uint, C.c_uint, C.c_int, dp, dp, dp]
        L.orc_ghbfix_setup.argtypes = [C.POINTER(Switch), C.c_double, C.c_double, C.c_double]
        L.orc_ghbfix_calculate.restype = C.c_size_t
        L.orc_ghbfix_calculate.argtypes = [C.c_void_p, C.POINTER(Pbc), C.c_int, C.POINTER(Switch), dp, up, up, C.c_uint, dp,
                                           C.c_size_t, C.c_uint, C.c_uint, C.c_int, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint))


def tools_pbc(x):
    return lib().orc_tools_pbc(float(x))


def make_switch(definition=None, nn=None, mm=0, r0=None, d0=0.0):
    """definition: the SWITCH={...} string; or (nn,mm,r0,d0): the R_0/NN/MM/D_0 keyword form"""
    s = Switch()
    if definition is not None:
        err = C.create_string_buffer(512)
        rc = lib().orc_switch_set(C.byref(s), definition.encode(), err, 512)
        if rc != 0:
            raise ValueError(err.value.decode())
    else:
        lib().orc_switch_set_rational(C.byref(s), int(nn), int(mm), float(r0), float(d0))
    return s


def switch_calculate(s, r):
    df = C.c_double(0)
    v = lib().orc_switch_calculate(C.byref(s), float(r), C.byref(df))
    return v, df.value


def switch_calculate_sqr(s, r2):
    df = C.c_double(0)
    v = lib().orc_switch_calculate_sqr(C.byref(s), float(r2), C.byref(df))
    return v, df.value


def make_pbc(box):
    """box: 9 numbers row-major (box[i][j] = j-th component of i-th lattice vector), zeros = no box"""
    p = Pbc()
    b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
    lib().orc_pbc_set_box(C.byref(p), _dp(b))
    return p


def lattice_reduce(box):
    b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9)).copy()
    lib().orc_lattice_reduce(_dp(b))
    return b.reshape(3, 3)


def pbc_distance(p, v1, v2):
    a = np.ascontiguousarray(v1, dtype=np.float64)
    b = np.ascontiguousarray(v2, dtype=np.float64)
    d = np.zeros(3)
    lib().orc_pbc_distance(C.byref(p), _dp(a), _dp(b), _dp(d))
    return d


def pbc_full_search(p, d):
    x = np.ascontiguousarray(d, dtype=np.float64).copy()
    lib().orc_pbc_full_search(C.byref(p), _dp(x))
    return x


def pbc_shifts(p):
    """list of 8 arrays (n_o,3), octant index = 4*(s0>0)+2*(s1>0)+(s2>0)"""
    sh = np.frombuffer(p.shifts, dtype=np.float64).reshape(8, MAXSHIFT, 3)
    return [sh[o, :p.nshift[o]].copy() for o in range(8)]


def linkcells(cutoff, pos, pbc):
    lc = LinkCells()
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    lib().orc_linkcells_setup(C.byref(lc), float(cutoff), _dp(pos), pos.shape[0], C.byref(pbc))
    return lc


def linkcells_find_cell(lc, p):
    x = np.ascontiguousarray(p, dtype=np.float64)
    return lib().orc_linkcells_find_cell(C.byref(lc), _dp(x))


def linkcells_required(lc, celn, use_pbc=True):
    c = np.ascontiguousarray(celn, dtype=np.uint32)
    out = np.zeros(27, dtype=np.uint32)
    n = lib().orc_linkcells_required(C.byref(lc), _up(c), int(use_pbc), _up(out))
    return out[:n].copy()


class NeighborList:
    """NeighborList restatement; style in {NL_PAIR, NL_TWOLIST, NL_SINGLELIST}"""

    def __init__(self, style, n0, n1=0, do_pbc=True, use_cells=False, cutoff=1e30, stride=0):
        self.h = lib().orc_nl_create(style, n0, n1, int(do_pbc), int(use_cells), float(cutoff), int(stride))
        self.style, self.n0, self.n1 = style, n0, (0 if style == NL_SINGLELIST else n1)
        self.stride = stride
        self.firsttime, self.invalidate = True, True

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_nl_free(self.h)
            self.h = None

    def update(self, pbc, pos, fast=False):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        assert pos.shape[0] == self.n0 + self.n1
        (lib().orc_nl_update_classic_cells if fast else lib().orc_nl_update)(self.h, C.byref(pbc), _dp(pos))

    def size(self):
        return lib().orc_nl_size(self.h)

    def pairs(self):
        n = self.size()
        ptr = lib().orc_nl_pairs(self.h)
        if not ptr or n == 0:
            return np.zeros((0, 2), dtype=np.uint32)
        return np.ctypeslib.as_array(ptr, shape=(n, 2)).copy()

    def prepare(self, step, exchange_step=False):
        ft, inv = C.c_int(int(self.firsttime)), C.c_int(int(self.invalidate))
        lib().orc_nl_prepare(self.h, int(step), int(exchange_step), C.byref(ft), C.byref(inv))
        self.firsttime, self.invalidate = bool(ft.value), bool(inv.value)
        return self.invalidate


def make_dhenergy(I, T=300.0, epsilon=80.0):
    """DHENERGY pairing constants (DHEnergy.cpp:104-128, default units)"""
    s = Switch()
    lib().orc_dhenergy_setup(C.byref(s), float(I), float(T), float(epsilon))
    return s


def make_ghbfix(dmax, d0, c):
    """GHBFIX polynomial constants (GHBFIX.cpp:98-113)"""
    s = Switch()
    lib().orc_ghbfix_setup(C.byref(s), float(dmax), float(d0), float(c))
    return s


# energy units of src/tools/Units.cpp (kJ/mol = 1), for GHBFIX's ENERGY_UNITS keyword
ENERGY_UNITS = {"kj/mol": 1.0, "kcal/mol": 4.184, "j/mol": 0.001, "eV": 96.48530749925792, "Ha": 2625.499638}


def read_ghbfix_tables(types_file, params_file, energy_units="plumed"):
    """GHBFIX.cpp:115-172: types per absolute atom index and the n x n eta table.  std::map::operator[] semantics are kept:
    a type name of the parameter file that the types file does not know reads as index 0."""
    def fields(fn):
        rows = []
        for ln in open(fn):
            ln = ln.split("#")[0].split()
            if ln:
                rows.append(ln)
        return rows
    table, types = {}, []
    for (name,) in fields(types_file):
        if name not in table:
            table[name] = (max(table.values()) + 1) if table else 0
        types.append(table[name])
    n = max(types) + 1
    etas = np.zeros(n * n)
    for it, jt, eta in fields(params_file):
        etas[n * table.setdefault(it, 0) + table.setdefault(jt, 0)] = float(eta)
    if energy_units != "plumed":
        etas *= ENERGY_UNITS[energy_units]
    return np.array(types, dtype=np.uint32), n, etas


def coordination(nl, pbc, do_pbc, sw, pos, abs_index=None, rank=0, nranks=1, nthreads=1, charges=None, types=None):
    """CoordinationBase::calculate: returns value, deriv (n,3), virial (3,3), pairs iterated.
    charges: per requested atom, for a DHENERGY pairing (make_dhenergy)"""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n = pos.shape[0]
    if abs_index is None:
        abs_index = np.arange(n, dtype=np.uint32)
    abs_index = np.ascontiguousarray(abs_index, dtype=np.uint32)
    val = C.c_double(0)
    deriv = np.zeros((n, 3))
    vir = np.zeros(9)
    if types is not None:  # (types per requested atom, ntypes, etas)
        t = np.ascontiguousarray(types[0], dtype=np.uint32)
        e = np.ascontiguousarray(types[2], dtype=np.float64)
        assert t.shape[0] == n
        npairs = lib().orc_ghbfix_calculate(nl.h, C.byref(pbc), int(do_pbc), C.byref(sw), _dp(pos), _up(abs_index), _up(t),
                                            int(types[1]), _dp(e), n, rank, nranks, nthreads, C.byref(val), _dp(deriv),
                                            _dp(vir))
    elif charges is not None:
        q = np.ascontiguousarray(charges, dtype=np.float64)
        assert q.shape[0] == n
        npairs = lib().orc_dhenergy_calculate(nl.h, C.byref(pbc), int(do_pbc), C.byref(sw), _dp(pos), _up(abs_index),
                                              _dp(q), n, rank, nranks, nthreads, C.byref(val), _dp(deriv), _dp(vir))
    else:
        npairs = lib().orc_coordination_calculate(nl.h, C.byref(pbc), int(do_pbc), C.byref(sw), _dp(pos),
                                                  _up(abs_index), n, rank, nranks, nthreads, C.byref(val),
                                                  _dp(deriv), _dp(vir))
    return val.value, deriv, vir.reshape(3, 3), npairs

"""TEST INFRASTRUCTURE ONLY -- ctypes access to the REAL reference built under oracle/_ref by
oracle/Makefile (`make ref`, needs /root/reference at build time only).

  * tool-level entry points of libref_tools_capi.so (reference SwitchingFunction / Pbc / LinkCells /
    NeighborList objects), used to pin oracle/coord_oracle.c and to generate tests/golden/;
  * class Plumed: drives the reference's PlumedMain through plumed_cmd exactly like `plumed driver`
    (src/cltools/Driver.cpp:526-552, :1009-1039) with numpy arrays as the MD engine's buffers.
    With `LOAD FILE=<our plugin>.so` as first input line this is also the drop-in test harness.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
KERNEL = os.path.join(REF_DIR, "lib", "libplumedKernel.so")
CAPI = os.path.join(REF_DIR, "lib", "libref_tools_capi.so")
PLUMED_BIN = os.path.join(REF_DIR, "bin", "plumed")


def available():
    return os.path.exists(KERNEL) and os.path.exists(CAPI)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref not built: run `make -C oracle ref` where /root/reference exists")
        C.CDLL(KERNEL, mode=C.RTLD_GLOBAL)
        L = C.CDLL(CAPI)
        dp = C.POINTER(C.c_double)
        up = C.POINTER(C.c_uint)
        L.ref_tools_pbc.restype = C.c_double
        L.ref_tools_pbc.argtypes = [C.c_double]
        L.ref_switch_create.restype = C.c_void_p
        L.ref_switch_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_switch_create_rational.restype = C.c_void_p
        L.ref_switch_create_rational.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_switch_free.argtypes = [C.c_void_p]
        L.ref_switch_calculate.restype = C.c_double
        L.ref_switch_calculate.argtypes = [C.c_void_p, C.c_double, dp]
        L.ref_switch_calculate_sqr.restype = C.c_double
        L.ref_switch_calculate_sqr.argtypes = [C.c_void_p, C.c_double, dp]
        L.ref_switch_data.argtypes = [C.c_void_p, dp]
        L.ref_lattice_reduce.argtypes = [dp]
        L.ref_pbc_create.restype = C.c_void_p
        L.ref_pbc_create.argtypes = [dp]
        L.ref_pbc_free.argtypes = [C.c_void_p]
        L.ref_pbc_is_orthorombic.argtypes = [C.c_void_p]
        L.ref_pbc_distance.argtypes = [C.c_void_p, dp, dp, dp]
        L.ref_pbc_distance_many.argtypes = [C.c_void_p, dp, dp, dp, C.c_size_t]
        L.ref_pbc_full_search.argtypes = [C.c_void_p, dp]
        L.ref_linkcells_create.restype = C.c_void_p
        L.ref_linkcells_create.argtypes = [C.c_double, dp, C.c_size_t, dp]
        L.ref_linkcells_free.argtypes = [C.c_void_p]
        L.ref_linkcells_ncells.argtypes = [C.c_void_p, up]
        L.ref_linkcells_find_cell.restype = C.c_uint
        L.ref_linkcells_find_cell.argtypes = [C.c_void_p, dp]
        L.ref_linkcells_required.restype = C.c_uint
        L.ref_linkcells_required.argtypes = [C.c_void_p, up, C.c_int, up]
        L.ref_nl_create.restype = C.c_void_p
        L.ref_nl_create.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_double, C.c_uint, dp]
        L.ref_nl_free.argtypes = [C.c_void_p]
        L.ref_nl_update.argtypes = [C.c_void_p, dp]
        L.ref_nl_size.restype = C.c_size_t
        L.ref_nl_size.argtypes = [C.c_void_p]
        L.ref_nl_pairs.argtypes = [C.c_void_p, up]
        L.ref_plumed_create.restype = C.c_void_p
        L.ref_plumed_cmd.restype = C.c_int
        L.ref_plumed_cmd.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_plumed_finalize.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint))


# ------------------------------------------------------------------ tool-level
class RefSwitch:
    def __init__(self, definition=None, nn=None, mm=0, r0=None, d0=0.0):
        if definition is not None:
            err = C.create_string_buffer(1024)
            self.h = lib().ref_switch_create(definition.encode(), err, 1024)
            if not self.h:
                raise ValueError(err.value.decode())
        else:
            self.h = lib().ref_switch_create_rational(int(nn), int(mm), float(r0), float(d0))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_switch_free(self.h)
            self.h = None

    def calculate(self, r):
        df = C.c_double(0)
        v = lib().ref_switch_calculate(self.h, float(r), C.byref(df))
        return v, df.value

    def calculate_sqr(self, r2):
        df = C.c_double(0)
        v = lib().ref_switch_calculate_sqr(self.h, float(r2), C.byref(df))
        return v, df.value

    def data(self):
        out = np.zeros(7)
        lib().ref_switch_data(self.h, _dp(out))
        return dict(zip(["d0", "dmax", "dmax_2", "invr0", "invr0_2", "stretch", "shift"], out))


class RefPbc:
    def __init__(self, box):
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
        self.h = lib().ref_pbc_create(_dp(b))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_pbc_free(self.h)
            self.h = None

    def is_ortho(self):
        return bool(lib().ref_pbc_is_orthorombic(self.h))

    def distance(self, v1, v2):
        a = np.ascontiguousarray(v1, dtype=np.float64)
        b = np.ascontiguousarray(v2, dtype=np.float64)
        if a.ndim == 2:
            d = np.zeros_like(a)
            lib().ref_pbc_distance_many(self.h, _dp(a), _dp(b), _dp(d), a.shape[0])
            return d
        d = np.zeros(3)
        lib().ref_pbc_distance(self.h, _dp(a), _dp(b), _dp(d))
        return d

    def full_search(self, d):
        x = np.ascontiguousarray(d, dtype=np.float64).copy()
        lib().ref_pbc_full_search(self.h, _dp(x))
        return x


def ref_lattice_reduce(box):
    b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9)).copy()
    lib().ref_lattice_reduce(_dp(b))
    return b.reshape(3, 3)


class RefLinkCells:
    def __init__(self, cutoff, pos, box):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
        self.h = lib().ref_linkcells_create(float(cutoff), _dp(pos), pos.shape[0], _dp(b))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_linkcells_free(self.h)
            self.h = None

    def ncells(self):
        out = np.zeros(3, dtype=np.uint32)
        lib().ref_linkcells_ncells(self.h, _up(out))
        return out

    def find_cell(self, p):
        x = np.ascontiguousarray(p, dtype=np.float64)
        return lib().ref_linkcells_find_cell(self.h, _dp(x))

    def required(self, celn, use_pbc=True):
        c = np.ascontiguousarray(celn, dtype=np.uint32)
        out = np.zeros(27, dtype=np.uint32)
        n = lib().ref_linkcells_required(self.h, _up(c), int(use_pbc), _up(out))
        return out[:n].copy()


class RefNeighborList:
    def __init__(self, style, n0, n1, do_pbc, use_cells, cutoff, stride, box):
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
        self.h = lib().ref_nl_create(style, n0, n1, int(do_pbc), int(use_cells), float(cutoff), int(stride), _dp(b))
        if not self.h:
            raise RuntimeError("reference NeighborList construction failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_nl_free(self.h)
            self.h = None

    def update(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        lib().ref_nl_update(self.h, _dp(pos))

    def size(self):
        return lib().ref_nl_size(self.h)

    def pairs(self):
        n = self.size()
        out = np.zeros((n, 2), dtype=np.uint32)
        if n:
            lib().ref_nl_pairs(self.h, _up(out))
        return out


# ------------------------------------------------------------------ plumed_cmd level
class PlumedError(RuntimeError):
    pass


class Plumed:
    """The reference PlumedMain driven like an MD engine (double precision, natural units).

    p = Plumed(natoms, ["c: COORDINATION GROUPA=1-100 R_0=0.3", "RESTRAINT ARG=c AT=0 SLOPE=1"])
    out = p.calc(step, positions(n,3), box(3,3))  ->  dict(forces=(n,3), virial=(3,3), bias=float)
    p.value("c") reads a scalar through the reference's own `setMemoryForData` route.
    """

    def __init__(self, natoms, lines, log="/dev/null", watch=()):
        L = lib()
        self.L = L
        self.p = L.ref_plumed_create()
        self.natoms = int(natoms)
        self._keep = []
        self.cmd("setRealPrecision", C.c_int(8))
        self.cmd("setMDEngine", b"b200-test")
        self.cmd("setNatoms", C.c_int(self.natoms))
        self.cmd("setTimestep", C.c_double(1.0))
        self.cmd("setLogFile", log.encode() if isinstance(log, str) else log)
        self.cmd("init", None)
        for ln in lines:
            self.cmd("readInputLine", ln.encode())
        self.watch = {}
        for name in watch:  # ActionToGetData ("GET"), src/core/ActionToGetData.cpp
            buf = np.zeros(1)
            self.cmd("readInputLine", ("grab_%s: GET ARG=%s" % (name, name)).encode())
            self.cmd("setMemoryForData " + name, buf)
            self.watch[name] = buf
        self.masses = np.ones(self.natoms)
        self.charges = np.zeros(self.natoms)

    def cmd(self, key, val):
        if val is None:
            ptr = None
        elif isinstance(val, (bytes, bytearray)):
            buf = C.create_string_buffer(bytes(val))
            self._keep.append(buf)
            ptr = C.cast(buf, C.c_void_p)
        elif isinstance(val, np.ndarray):
            ptr = C.c_void_p(val.ctypes.data)
        else:  # a ctypes scalar
            self._keep.append(val)
            ptr = C.cast(C.pointer(val), C.c_void_p)
        err = C.create_string_buffer(4096)
        rc = self.L.ref_plumed_cmd(self.p, key.encode(), ptr, err, 4096)
        if rc != 0:
            raise PlumedError(err.value.decode(errors="replace"))

    def calc(self, step, pos, box=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        assert pos.shape == (self.natoms, 3)
        forces = np.zeros_like(pos)
        virial = np.zeros((3, 3))
        b = np.zeros((3, 3)) if box is None else np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(3, 3))
        bias = C.c_double(0)
        self.cmd("setStep", C.c_int(int(step)))
        self.cmd("setPositions", pos)
        self.cmd("setMasses", self.masses)
        self.cmd("setCharges", self.charges)
        self.cmd("setBox", b)
        self.cmd("setForces", forces)
        self.cmd("setVirial", virial)
        self.cmd("calc", None)
        self.cmd("getBias", bias)
        self._keep = self._keep[-64:]
        return dict(forces=forces, virial=virial, bias=bias.value)

    def value(self, name):
        return float(self.watch[name][0])

    def close(self):
        if self.p:
            self.L.ref_plumed_finalize(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host-side mirror of the reference's COORDINATION action interface on top of the C ABI.

`Coordination` takes the same keyword set as the reference action
(src/colvar/CoordinationBase.cpp:30-40, src/colvar/Coordination.cpp:115-126):
GROUPA GROUPB PAIR NLIST NLISTCELLS NL_CUTOFF NL_STRIDE NOPBC SERIAL SWITCH R_0 NN MM D_0, plus the
additive top-level D_MAX of the CUDA prototype (plugins/cudaCoord/src/Coordination.cu:1008-1011, :1469-1508),
raises the reference's error texts for bad keyword combinations, and follows the action life cycle
prepare() -> calculate() of src/core/PlumedMain.cpp:1162-1326.  All numerics happen in libb200coord.so.

The C++ twin of this file, used inside a real PLUMED, is plumed2_b200/csrc/plugin/CoordinationB200.cpp.
"""
import ctypes as C
import re

import numpy as np

from . import capi


class PlumedInputError(ValueError):
    """what Action::error() / plumed_assert would raise in the reference"""


# ------------------------------------------------------------------ input-line parsing
def parse_atom_list(text):
    """PLUMED atom-list syntax (serials, 1-based): '1-100', '1,4,9', '1-100:2', mixtures thereof.
    Returns 0-based absolute indices in the order given (src/tools/Tools.cpp interpretRanges)."""
    if text is None:
        return np.zeros(0, dtype=np.uint32)
    if isinstance(text, (list, tuple, np.ndarray)):
        return np.asarray(text, dtype=np.int64).astype(np.uint32)
    out = []
    for tok in str(text).replace("{", " ").replace("}", " ").replace(",", " ").split():
        m = re.fullmatch(r"(\d+)-(\d+)(?::(-?\d+))?", tok)
        if m:
            a, b = int(m.group(1)), int(m.group(2))
            st = int(m.group(3)) if m.group(3) else 1
            if st == 0:
                raise PlumedInputError("interpreting ranges: stride cannot be zero in " + tok)
            if st > 0:
                out.extend(range(a, b + 1, st))
            else:
                out.extend(range(a, b - 1, st))
        elif re.fullmatch(r"\d+", tok):
            out.append(int(tok))
        else:
            raise PlumedInputError("cannot interpret atom list element '%s'" % tok)
    arr = np.asarray(out, dtype=np.int64)
    if arr.size and arr.min() < 1:
        raise PlumedInputError("atom serials start from 1")
    return (arr - 1).astype(np.uint32)


def split_input_line(line):
    """'lab: COORDINATION K=V FLAG K={a b}' -> (label, action, {K: V}, [FLAGS]); braces group words
    (src/tools/Tools.cpp getWords)"""
    words, cur, depth = [], "", 0
    for ch in line.strip():
        if ch == "{":
            depth += 1
            if depth == 1:
                continue
        elif ch == "}":
            depth -= 1
            if depth == 0:
                continue
            if depth < 0:
                raise PlumedInputError("unmatched } in input line")
        if ch.isspace() and depth == 0:
            if cur:
                words.append(cur)
            cur = ""
        else:
            cur += ch
    if depth != 0:
        raise PlumedInputError("unmatched { in input line")
    if cur:
        words.append(cur)
    label = None
    if words and words[0].endswith(":"):
        label = words.pop(0)[:-1]
    if not words:
        raise PlumedInputError("empty input line")
    action = words.pop(0)
    kv, flags = {}, []
    for w in words:
        if "=" in w:
            k, v = w.split("=", 1)
            if k == "LABEL":
                label = v
            else:
                kv[k] = v
        else:
            flags.append(w)
    return label, action, kv, flags


_KEYS = {"GROUPA", "GROUPB", "NL_CUTOFF", "NL_STRIDE", "SWITCH", "R_0", "NN", "MM", "D_0", "D_MAX"}
_FLAGS = {"PAIR", "NLIST", "NLISTCELLS", "NOPBC", "SERIAL", "NUMERICAL_DERIVATIVES"}


class Coordination:
    """COORDINATION on one B200.  positions/box follow PLUMED conventions: (natoms,3) array in nm of ALL
    system atoms (the action gathers the ones it requested), box = 3x3 row vectors or None."""

    def __init__(self, GROUPA, GROUPB=None, PAIR=False, NLIST=False, NLISTCELLS=False, NL_CUTOFF=None,
                 NL_STRIDE=None, NOPBC=False, SERIAL=False, SWITCH=None, R_0=None, NN=6, MM=0, D_0=0.0, D_MAX=None,
                 label="c", device=-1, rank=0, nranks=1, precision=capi.FP64, DHENERGY=None, GHBFIX=None):
        """DHENERGY: None for COORDINATION, else dict(I=..., TEMP=..., EPSILON=...) -> the sibling action DHENERGY
        (src/colvar/DHEnergy.cpp): same groups / lists, Debye-Hueckel pairing, needs set_charges()"""
        self.label = label
        ga = parse_atom_list(GROUPA)
        gb = parse_atom_list(GROUPB)
        if ga.size == 0:
            raise PlumedInputError("GROUPA: no atoms specified")
        # --- CoordinationBase ctor, CoordinationBase.cpp:49-83
        if NLIST and NLISTCELLS:
            raise PlumedInputError("Please activate only one of the two version of the NL")
        if NLISTCELLS and PAIR:
            raise PlumedInputError("Pair is not compatible with the CELLS implementation of the NL")
        doneigh = bool(NLIST or NLISTCELLS)
        nl_cut, nl_st = 0.0, 0
        if doneigh:
            nl_cut = float(NL_CUTOFF) if NL_CUTOFF is not None else 0.0
            if nl_cut <= 0.0:
                raise PlumedInputError("NL_CUTOFF should be explicitly specified and positive")
            nl_st = int(NL_STRIDE) if NL_STRIDE is not None else 0
            if nl_st <= 0:
                raise PlumedInputError("NL_STRIDE should be explicitly specified and positive")
        if PAIR and ga.size != gb.size:
            raise PlumedInputError("when using PAIR option, the two groups should have the same number of elements\n"
                                   "the groups you specified have size %d and %d" % (ga.size, gb.size))
        # --- Coordination ctor, Coordination.cpp:128-157 (+ cudaCoord's top-level D_MAX rewrite)
        if GHBFIX is not None:  # GHBFIX ctor, GHBFIX.cpp:98-113; dict(D_MAX, D_0, C); types/etas: set_types()
            self.switch = capi.pairing_ghbfix(float(GHBFIX["D_MAX"]), float(GHBFIX["D_0"]), float(GHBFIX["C"]))
        elif DHENERGY is not None:  # DHEnergy ctor, DHEnergy.cpp:104-128 (default units)
            self.switch = capi.pairing_dhenergy(float(DHENERGY["I"]), float(DHENERGY.get("TEMP", 300.0)),
                                                float(DHENERGY["EPSILON"]))
        elif SWITCH:
            try:
                self.switch = capi.switch_parse(str(SWITCH))
            except capi.B200CoordError as e:
                if e.code == capi.ERR_UNSUPPORTED:
                    raise
                raise PlumedInputError("problem reading SWITCH keyword : " + e.message)
        else:
            r0 = float(R_0) if R_0 is not None else 0.0
            if r0 <= 0.0:
                raise PlumedInputError("R_0 should be explicitly specified and positive")
            if D_MAX is not None:
                self.switch = capi.switch_parse("RATIONAL R_0=%r D_0=%r NN=%d MM=%d D_MAX=%r" %
                                                (r0, float(D_0), int(NN), int(MM), float(D_MAX)))
            else:
                self.switch = capi.switch_rational(int(NN), int(MM), r0, float(D_0))
        self.group_a, self.group_b = ga, gb
        self.atoms = np.concatenate([ga, gb]).astype(np.uint32)  # NeighborList::getFullAtomList order
        self.n = int(self.atoms.size)
        self.pbc = not NOPBC
        self.serial = bool(SERIAL)
        if gb.size == 0:
            style = capi.STYLE_SINGLELIST
        else:
            style = capi.STYLE_PAIR if PAIR else capi.STYLE_TWOLIST
        self.style = style
        self.nl_mode = capi.NL_CELLS if NLISTCELLS else (capi.NL_CLASSIC if NLIST else capi.NL_NONE)
        self.nl_cutoff, self.nl_stride = nl_cut, nl_st
        cfg = capi.Config(capi.ABI_VERSION, int(device), int(precision), style, int(ga.size), int(gb.size),
                          int(self.pbc), self.nl_mode, nl_cut, nl_st, int(rank), int(nranks))
        self._cfg = cfg
        self._ctx = C.c_void_p()
        L = capi.lib()
        absidx = np.ascontiguousarray(self.atoms)
        capi.check(L.b200coord_create(C.byref(cfg), C.byref(self.switch),
                                      absidx.ctypes.data_as(C.POINTER(C.c_uint)), C.byref(self._ctx)))
        self._L = L
        self._pos = np.zeros((self.n, 3))
        self.derivatives = np.zeros((self.n, 3))
        self.virial = np.zeros((3, 3))
        self.value = 0.0
        self._zero_box = np.zeros(9)
        self.description = capi.switch_describe(self.switch)

    # ---- construction from a PLUMED input line
    @classmethod
    def from_input(cls, line, **kw):
        label, action, kv, flags = split_input_line(line)
        if action == "GHBFIX":  # GHBFIX::registerKeywords, GHBFIX.cpp:91-103; the TYPES / PARAMS files are read by
            # set_types_from_files (the mirror does not open files behind the caller's back)
            allowed = {"GROUPA", "GROUPB", "NL_CUTOFF", "NL_STRIDE", "D_MAX", "D_0", "C", "TYPES", "PARAMS", "ENERGY_UNITS"}
            for k in kv:
                if k not in allowed:
                    raise PlumedInputError("cannot understand the following words from the input line : " + k)
            for f in flags:
                if f not in _FLAGS:
                    raise PlumedInputError("cannot understand the following words from the input line : " + f)
            for k in ("GROUPA", "TYPES", "PARAMS", "D_MAX", "D_0", "C"):
                if k not in kv:
                    raise PlumedInputError("%s is compulsory" % k)
            args = dict(GROUPA=kv["GROUPA"], GROUPB=kv.get("GROUPB"), PAIR="PAIR" in flags, NLIST="NLIST" in flags,
                        NLISTCELLS="NLISTCELLS" in flags, NL_CUTOFF=kv.get("NL_CUTOFF"), NL_STRIDE=kv.get("NL_STRIDE"),
                        NOPBC="NOPBC" in flags, SERIAL="SERIAL" in flags, label=label or "c",
                        GHBFIX=dict(D_MAX=float(kv["D_MAX"]), D_0=float(kv["D_0"]), C=float(kv["C"])))
            args.update(kw)
            obj = cls(**args)
            obj.ghbfix_files = (kv["TYPES"], kv["PARAMS"], kv.get("ENERGY_UNITS", "plumed"))
            return obj
        if action == "DHENERGY":  # DHEnergy::registerKeywords, DHEnergy.cpp:77-90: CoordinationBase + I, TEMP, EPSILON
            allowed = {"GROUPA", "GROUPB", "NL_CUTOFF", "NL_STRIDE", "I", "TEMP", "EPSILON"}
            for k in kv:
                if k not in allowed:
                    raise PlumedInputError("cannot understand the following words from the input line : " + k)
            for f in flags:
                if f not in _FLAGS:
                    raise PlumedInputError("cannot understand the following words from the input line : " + f)
            if "GROUPA" not in kv:
                raise PlumedInputError("GROUPA: no atoms specified")
            args = dict(GROUPA=kv["GROUPA"], GROUPB=kv.get("GROUPB"), PAIR="PAIR" in flags, NLIST="NLIST" in flags,
                        NLISTCELLS="NLISTCELLS" in flags, NL_CUTOFF=kv.get("NL_CUTOFF"), NL_STRIDE=kv.get("NL_STRIDE"),
                        NOPBC="NOPBC" in flags, SERIAL="SERIAL" in flags, label=label or "c",
                        DHENERGY=dict(I=float(kv.get("I", 1.0)), TEMP=float(kv.get("TEMP", 300.0)),
                                      EPSILON=float(kv.get("EPSILON", 80.0))))
            args.update(kw)
            return cls(**args)
        if action != "COORDINATION":
            raise PlumedInputError("this mirror only implements COORDINATION, DHENERGY and GHBFIX, got " + action)
        for k in kv:
            if k not in _KEYS:
                raise PlumedInputError("cannot understand the following words from the input line : " + k)
        for f in flags:
            if f not in _FLAGS:
                raise PlumedInputError("cannot understand the following words from the input line : " + f)
        if "GROUPA" not in kv:
            raise PlumedInputError("GROUPA: no atoms specified")
        args = dict(GROUPA=kv["GROUPA"], GROUPB=kv.get("GROUPB"), PAIR="PAIR" in flags, NLIST="NLIST" in flags,
                    NLISTCELLS="NLISTCELLS" in flags, NL_CUTOFF=kv.get("NL_CUTOFF"), NL_STRIDE=kv.get("NL_STRIDE"),
                    NOPBC="NOPBC" in flags, SERIAL="SERIAL" in flags, SWITCH=kv.get("SWITCH"), R_0=kv.get("R_0"),
                    NN=int(kv.get("NN", 6)), MM=int(kv.get("MM", 0)), D_0=float(kv.get("D_0", 0.0)),
                    D_MAX=kv.get("D_MAX"), label=label or "c")
        args.update(kw)
        return cls(**args)

    # ---- action life cycle
    def prepare(self, step, exchange_step=False):
        """CoordinationBase::prepare: returns True when the next calculate() rebuilds the list"""
        flag = C.c_int(0)
        capi.check(self._L.b200coord_prepare(self._ctx, int(step), int(bool(exchange_step)), C.byref(flag)), self._ctx)
        return bool(flag.value)

    def set_charges(self, charges):
        """charges of ALL system atoms (ActionAtomistic::getCharge); required for DHENERGY"""
        q = np.ascontiguousarray(np.asarray(charges, dtype=np.float64)[self.atoms])
        capi.check(self._L.b200coord_set_charges(self._ctx, q.ctypes.data_as(C.POINTER(C.c_double))), self._ctx)

    def set_types(self, types_of_all_atoms, ntypes, etas):
        """GHBFIX: interaction type per SYSTEM atom (typesTable[absolute index]) and the ntypes x ntypes eta table"""
        t = np.ascontiguousarray(np.asarray(types_of_all_atoms, dtype=np.uint32)[self.atoms])
        e = np.ascontiguousarray(np.asarray(etas, dtype=np.float64).reshape(-1))
        capi.check(self._L.b200coord_set_types(self._ctx, t.ctypes.data_as(C.POINTER(C.c_uint)), int(ntypes),
                                               e.ctypes.data_as(C.POINTER(C.c_double))), self._ctx)

    def _set_box(self, box):
        b = self._zero_box if box is None else np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
        capi.check(self._L.b200coord_set_box(self._ctx, b.ctypes.data_as(C.POINTER(C.c_double))), self._ctx)

    def gather(self, positions):
        """ActionAtomistic::retrieveAtoms: requested atoms in GROUPA-then-GROUPB order"""
        positions = np.asarray(positions, dtype=np.float64)
        np.take(positions, self.atoms, axis=0, out=self._pos)
        return self._pos

    def calculate(self, positions, box=None, gathered=False):
        """CoordinationBase::calculate.  Sets .value, .derivatives (n,3), .virial (3,3); returns value."""
        self._set_box(box)
        pos = np.ascontiguousarray(positions, dtype=np.float64) if gathered else self.gather(positions)
        if pos.shape != (self.n, 3):
            raise ValueError("positions must be (%d,3) when gathered=True" % self.n)
        val = C.c_double(0)
        capi.check(self._L.b200coord_calculate(self._ctx, pos.ctypes.data_as(C.c_void_p), C.byref(val),
                                               self.derivatives.ctypes.data_as(C.c_void_p),
                                               self.virial.ctypes.data_as(C.POINTER(C.c_double))), self._ctx)
        self.value = val.value
        return self.value

    def update_list(self, positions, box=None, gathered=False):
        """NeighborList::update right now"""
        self._set_box(box)
        pos = np.ascontiguousarray(positions, dtype=np.float64) if gathered else self.gather(positions)
        capi.check(self._L.b200coord_update_list(self._ctx, pos.ctypes.data_as(C.c_void_p)), self._ctx)

    def apply(self, natoms, force_on_value=1.0):
        """Colvar::apply: forces on the system atoms and the virial contribution for a force on the value"""
        f = np.zeros((natoms, 3))
        np.add.at(f, self.atoms, force_on_value * self.derivatives)
        return f, force_on_value * self.virial

    # ---- inspection
    def stats(self):
        s = capi.Stats()
        capi.check(self._L.b200coord_get_stats(self._ctx, C.byref(s)), self._ctx)
        return {k: (list(getattr(s, k)) if k == "ncells" else getattr(s, k)) for k, _ in capi.Stats._fields_}

    def neighbor_pairs(self):
        """current list as (npairs,2) indices into the requested-atom array, sorted"""
        n = C.c_ulonglong(0)
        capi.check(self._L.b200coord_nl_pairs(self._ctx, None, 0, C.byref(n)), self._ctx)
        out = np.zeros((n.value, 2), dtype=np.uint32)
        if n.value:
            capi.check(self._L.b200coord_nl_pairs(self._ctx, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)),
                       self._ctx)
        return out

    def comm_init(self, unique_id):
        capi.check(self._L.b200coord_comm_init(self._ctx, bytes(unique_id)), self._ctx)

    def peer_export(self):
        """IPC handles of this rank's derivative-row buffers (fused sweep + exchange over NVLink)"""
        buf = C.create_string_buffer(capi.PEER_HANDLE_BYTES)
        capi.check(self._L.b200coord_peer_export(self._ctx, buf), self._ctx)
        return buf.raw

    def peer_attach(self, handles_of_all_ranks):
        """handles_of_all_ranks: list of the peer_export() blobs in rank order"""
        blob = b"".join(bytes(h) for h in handles_of_all_ranks)
        capi.check(self._L.b200coord_peer_attach(self._ctx, blob), self._ctx)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.b200coord_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(n, rank, nranks):
    """the equal-chunk partition the library uses for i-atom rows and for position slices:
    chunk = ceil(n/nranks), rank r owns [r*chunk, min(n,(r+1)*chunk))  (cf. CoordinationBase.cpp:168-170)"""
    chunk = (n + nranks - 1) // nranks
    lo = min(n, chunk * rank)
    return lo, min(n, lo + chunk) - lo


def comm_unique_id():
    buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
    capi.check(capi.lib().b200coord_comm_unique_id(buf))
    return buf.raw

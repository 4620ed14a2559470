"""CPU tests of the drop-in boundary: libb200coord.so loads, exports every symbol include/b200coord.h
declares, the host-side set-up code (switch parsing, stretch) agrees with the oracle field by field, and the
action mirror raises the reference's keyword errors.  No compute call is made here (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import plumed2_b200 as P
from oracle import oracle as O
from plumed2_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFS = ["RATIONAL R_0=0.3", "RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8", "RATIONAL R_0=0.3 NN=8 D_MAX=0.8",
        "RATIONAL R_0=0.3 NN=12 D_MAX=0.9", "RATIONAL R_0=0.3 NN=2", "RATIONAL R_0=0.3 NN=4 MM=10 D_MAX=0.9",
        "RATIONAL R_0=0.3 NN=5 MM=11 D_MAX=0.9", "RATIONAL R_0=0.3 NN=5 D_MAX=0.9", "RATIONAL R_0=0.3 NN=14 D_MAX=0.9",
        "RATIONAL R_0=0.3 D_0=0.1 D_MAX=0.9", "RATIONAL R_0=0.3 D_MAX=0.9 NOSTRETCH", "EXP R_0=0.2 D_MAX=0.9",
        "EXP R_0=0.8 D_0=0.5 D_MAX=2.6", "GAUSSIAN R_0=1.0 D_0=0.0 D_MAX=2.6", "GAUSSIAN R_0=1.0 D_0=0.3 D_MAX=2.6",
        "SMAP R_0=1.3 A=3 B=2 D_MAX=2.6", "CUBIC D_MAX=2.6 D_0=0.6", "TANH R_0=1.3 D_MAX=2.6", "COSINUS R_0=2.6",
        "Q R_0=1.0 D_0=0.3 BETA=5.0 LAMBDA=1.0 REF=1.3 D_MAX=2.6", "{RATIONAL R_0=0.3 STRETCH D_MAX=1.0}"]

FIELDS = ["type", "d0", "dmax", "dmax_2", "invr0", "invr0_2", "stretch", "shift", "nn", "mm", "preRes", "preDfunc",
          "preSecDev", "nnf", "mmf", "preDfuncF", "preSecDevF", "a", "b", "c", "d", "beta", "lambda_", "ref"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200coord.h")).read()
    declared = sorted(set(re.findall(r"\b(b200coord_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libb200coord.so does not export " + name
    assert sorted(capi.EXPORTED) == declared
    assert capi.lib().b200coord_abi_version() == capi.ABI_VERSION


def test_plugin_library_registers_coordination():
    """the PLUMED-facing .so carries the action registration and links to the C ABI library"""
    plugin = os.path.join(os.path.dirname(capi.LIB_PATH), "libb200coord_plumed.so")
    if not os.path.exists(plugin):
        pytest.skip("plugin not built (needs the reference headers at build time)")
    blob = open(plugin, "rb").read()
    assert b"COORDINATION" in blob and b"libb200coord.so" in blob and b"NL_CUTOFF" in blob


@pytest.mark.parametrize("definition", DEFS)
def test_switch_parse_matches_oracle(definition):
    a = capi.switch_parse(definition)
    b = O.make_switch(definition)
    for f in FIELDS:
        assert getattr(a, f) == getattr(b, f), (definition, f, getattr(a, f), getattr(b, f))


@pytest.mark.parametrize("nn,mm,r0,d0", [(6, 0, 0.3, 0.0), (8, 0, 0.25, 0.0), (6, 12, 0.3, 0.1), (4, 10, 0.3, 0.0),
                                          (7, 0, 1.0, 0.1), (5, 9, 0.4, 0.0), (12, 24, 0.5, 0.0)])
def test_switch_keyword_form_matches_oracle(nn, mm, r0, d0):
    a = capi.switch_rational(nn, mm, r0, d0)
    b = O.make_switch(nn=nn, mm=mm, r0=r0, d0=d0)
    for f in FIELDS:
        assert getattr(a, f) == getattr(b, f), (f, getattr(a, f), getattr(b, f))
    assert "dmax" in capi.switch_describe(a)


def test_switch_parse_errors():
    for bad, code in [("", capi.ERR_PARSE), ("FOO R_0=1", capi.ERR_PARSE), ("RATIONAL", capi.ERR_PARSE),
                      ("RATIONAL R_0=1 BAR=2", capi.ERR_PARSE), ("SMAP R_0=1", capi.ERR_PARSE),
                      ("RATIONAL R_0=abc", capi.ERR_PARSE), ("CUSTOM FUNC=1/(1+x^6) R_0=1", capi.ERR_UNSUPPORTED)]:
        with pytest.raises(capi.B200CoordError) as e:
            capi.switch_parse(bad)
        assert e.value.code == code, bad
    with pytest.raises(capi.B200CoordError):
        capi.switch_rational(6, 0, 0.0, 0.0)
    s = capi.Switch()
    err = C.create_string_buffer(256)
    assert capi.lib().b200coord_switch_parse(b"RATIONAL R_0=1 BAR=2", C.byref(s), err, 256) == capi.ERR_PARSE
    assert b"rogue keywords" in err.value and b"BAR=2" in err.value


def test_atom_list_parsing():
    assert P.parse_atom_list("1-5").tolist() == [0, 1, 2, 3, 4]
    assert P.parse_atom_list("1,3,9").tolist() == [0, 2, 8]
    assert P.parse_atom_list("1-10:3").tolist() == [0, 3, 6, 9]
    assert P.parse_atom_list("2-4,10,20-22").tolist() == [1, 2, 3, 9, 19, 20, 21]
    with pytest.raises(P.PlumedInputError):
        P.parse_atom_list("a-b")


@pytest.mark.parametrize("line,msg", [
    ("c: COORDINATION GROUPA=1-10 R_0=0.3 NLIST NLISTCELLS NL_CUTOFF=1 NL_STRIDE=2", "only one of the two version"),
    ("c: COORDINATION GROUPA=1-10 GROUPB=11-20 R_0=0.3 PAIR NLISTCELLS NL_CUTOFF=1 NL_STRIDE=2", "Pair is not compatible"),
    ("c: COORDINATION GROUPA=1-10 R_0=0.3 NLIST NL_STRIDE=2", "NL_CUTOFF should be explicitly specified and positive"),
    ("c: COORDINATION GROUPA=1-10 R_0=0.3 NLIST NL_CUTOFF=1.0", "NL_STRIDE should be explicitly specified and positive"),
    ("c: COORDINATION GROUPA=1-10 R_0=0.3 NLIST NL_CUTOFF=-1.0 NL_STRIDE=2", "NL_CUTOFF should be"),
    ("c: COORDINATION GROUPA=1-10", "R_0 should be explicitly specified and positive"),
    ("c: COORDINATION GROUPA=1-10 R_0=-2", "R_0 should be explicitly specified and positive"),
    ("c: COORDINATION GROUPA=1-10 GROUPB=11-15 R_0=0.3 PAIR", "same number of elements"),
    ("c: COORDINATION GROUPA=1-10 SWITCH={RATIONAL R_0=0.3 FOO=1}", "problem reading SWITCH keyword"),
    ("c: COORDINATION GROUPA=1-10 R_0=0.3 BOGUS=1", "cannot understand"),
    ("c: COORDINATION R_0=0.3", "GROUPA"),
    ("c: DISTANCE ATOMS=1,2", "only implements COORDINATION"),
])
def test_keyword_errors_mirror_reference(line, msg):
    """the checks of CoordinationBase.cpp:69-83 and Coordination.cpp:131-152 fire before any device work"""
    with pytest.raises(P.PlumedInputError) as e:
        P.Coordination.from_input(line)
    assert msg in str(e.value)


def test_custom_switch_is_refused_not_emulated():
    with pytest.raises(capi.B200CoordError) as e:
        P.Coordination.from_input("c: COORDINATION GROUPA=1-10 SWITCH={CUSTOM FUNC=1/(1+x^6) R_0=1}")
    assert e.value.code == capi.ERR_UNSUPPORTED


def test_input_line_splitting():
    lab, act, kv, flags = P.coordination.split_input_line(
        "cn: COORDINATION GROUPA=1-100 GROUPB=101-200 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10 NOPBC")
    assert lab == "cn" and act == "COORDINATION" and kv["SWITCH"] == "RATIONAL R_0=0.3 D_MAX=0.8"
    assert set(flags) == {"NLIST", "NOPBC"} and kv["NL_STRIDE"] == "10"


def test_create_fails_loudly_without_a_gpu():
    """no CPU fallback: on a machine without a CUDA device context creation is an error"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.B200CoordError) as e:
        P.Coordination.from_input("c: COORDINATION GROUPA=1-100 R_0=0.3")
    assert e.value.code == capi.ERR_CUDA


def test_create_rejects_bad_configs():
    sw = capi.switch_rational(6, 0, 0.3, 0.0)
    out = C.c_void_p()
    L = capi.lib()
    absidx = np.arange(10, dtype=np.uint32)
    ap = absidx.ctypes.data_as(C.POINTER(C.c_uint))
    def cfg(**kw):
        base = dict(abi_version=capi.ABI_VERSION, device=-1, precision=0, style=2, n_group_a=10, n_group_b=0, pbc=1,
                    nl_mode=0, nl_cutoff=0.0, nl_stride=0, rank=0, nranks=1)
        base.update(kw)
        return capi.Config(*[base[k] for k, _ in capi.Config._fields_])
    for bad in [cfg(abi_version=99), cfg(style=7), cfg(nl_mode=2, style=0, n_group_a=5, n_group_b=5, nl_cutoff=1.0, nl_stride=1),
                cfg(nl_mode=1, nl_cutoff=0.0, nl_stride=1), cfg(nl_mode=1, nl_cutoff=1.0, nl_stride=0),
                cfg(style=0, n_group_a=4, n_group_b=6), cfg(n_group_a=0), cfg(style=2, n_group_b=3),
                cfg(precision=2), cfg(precision=-1)]:  # B200COORD_FP64 = 0 and B200COORD_FP32 = 1 are the two modes
        rc = L.b200coord_create(C.byref(bad), C.byref(sw), ap, C.byref(out))
        assert rc == capi.ERR_INVALID and not out.value
        assert len(L.b200coord_last_error(None)) > 0


def test_coupling_table_rejects_what_is_not_device_memory():
    """b200coord_coupling_publish takes device pointers of the stated device only; a name nobody published is an error
    for lookup and withdraw (host logic of the device-resident MD coupling, no kernels involved)"""
    L = capi.lib()
    host = np.zeros(30)
    hp = host.ctypes.data_as(C.c_void_p)
    assert L.b200coord_coupling_publish(b"x", 0, hp, hp, 10) == capi.ERR_INVALID
    assert L.b200coord_coupling_publish(b"", 0, hp, hp, 10) == capi.ERR_INVALID
    assert L.b200coord_coupling_publish(b"x", 0, None, None, 10) == capi.ERR_INVALID
    dev, n = C.c_int(-5), C.c_size_t(0)
    assert L.b200coord_coupling_lookup(b"x", C.byref(dev), None, None, C.byref(n)) != 0
    assert dev.value == -5 and b"no coupling named x" in L.b200coord_last_error(None)
    assert L.b200coord_coupling_withdraw(b"x") != 0

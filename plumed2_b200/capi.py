"""ctypes binding of libb200coord.so -- one Python function per entry point of include/b200coord.h.

No numerics happen here and there is no fallback: if the CUDA library is missing or cannot load, importing
`lib()` raises, and every non-zero return code of the C ABI becomes a B200CoordError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200coord.so")

ABI_VERSION = 1
OK, ERR_INVALID, ERR_CUDA, ERR_PARSE, ERR_UNSUPPORTED, ERR_NCCL, ERR_STATE = range(7)
STYLE_PAIR, STYLE_TWOLIST, STYLE_SINGLELIST = 0, 1, 2
NL_NONE, NL_CLASSIC, NL_CELLS = 0, 1, 2
FP64, FP32 = 0, 1
UNIQUE_ID_BYTES = 128
PEER_HANDLE_BYTES = 256

SW_NAMES = ["rationalfix12", "rationalfix10", "rationalfix8", "rationalfix6", "rationalfix4", "rationalfix2",
            "rational", "rationalFast", "rationalSimple", "rationalSimpleFast", "exponential", "gaussian",
            "fastgaussian", "smap", "cubic", "tanh", "cosinus", "nativeq", "lepton", "not_initialized"]

# every symbol include/b200coord.h declares (tests check the library exports all of them)
EXPORTED = ["b200coord_abi_version", "b200coord_switch_parse", "b200coord_switch_rational",
            "b200coord_switch_describe", "b200coord_create", "b200coord_destroy", "b200coord_last_error",
            "b200coord_set_box", "b200coord_prepare", "b200coord_update_list", "b200coord_calculate",
            "b200coord_calculate_device", "b200coord_get_stats", "b200coord_nl_pairs",
            "b200coord_comm_unique_id", "b200coord_comm_init", "b200coord_host_alloc", "b200coord_host_free",
            "b200coord_coupling_publish", "b200coord_coupling_withdraw", "b200coord_coupling_lookup",
            "b200coord_coupled_set_index", "b200coord_calculate_coupled", "b200coord_apply_coupled",
            "b200coord_coupled_derivatives", "b200coord_submit", "b200coord_collect",
            "b200coord_device_alloc", "b200coord_device_free", "b200coord_memcpy_h2d", "b200coord_memcpy_d2h",
            "b200coord_device_synchronize", "b200coord_enqueue_device", "b200coord_stream_mark",
            "b200coord_stream_elapsed_ms", "b200coord_calculate_distributed", "b200coord_my_slice",
            "b200coord_measure_fp64_peak", "b200coord_peer_export", "b200coord_peer_attach",
            "b200coord_pairing_dhenergy", "b200coord_set_charges", "b200coord_pairing_ghbfix",
            "b200coord_set_types", "b200coord_device_count", "b200coord_enqueue_device_distributed",
            "b200coord_nl_pairs_device", "b200coord_peer_attach_local", "b200coord_group_create",
            "b200coord_group_destroy", "b200coord_group_size", "b200coord_group_context", "b200coord_group_last_error",
            "b200coord_group_set_box", "b200coord_group_prepare", "b200coord_group_set_charges",
            "b200coord_group_set_types", "b200coord_group_calculate"]


class B200CoordError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200coord error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


class Switch(C.Structure):
    _fields_ = [("type", C.c_int), ("d0", C.c_double), ("dmax", C.c_double), ("dmax_2", C.c_double),
                ("invr0", C.c_double), ("invr0_2", C.c_double), ("stretch", C.c_double), ("shift", C.c_double),
                ("nn", C.c_int), ("mm", C.c_int), ("preRes", C.c_double), ("preDfunc", C.c_double),
                ("preSecDev", C.c_double), ("nnf", C.c_int), ("mmf", C.c_int), ("preDfuncF", C.c_double),
                ("preSecDevF", C.c_double), ("a", C.c_int), ("b", C.c_int), ("c", C.c_double), ("d", C.c_double),
                ("beta", C.c_double), ("lambda_", C.c_double), ("ref", C.c_double)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("precision", C.c_int), ("style", C.c_int),
                ("n_group_a", C.c_uint), ("n_group_b", C.c_uint), ("pbc", C.c_int), ("nl_mode", C.c_int),
                ("nl_cutoff", C.c_double), ("nl_stride", C.c_int), ("rank", C.c_int), ("nranks", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("nl_size", C.c_ulonglong), ("pair_evals", C.c_ulonglong), ("kernel_launches", C.c_ulonglong),
                ("rebuilds", C.c_ulonglong), ("last_sweep_ms", C.c_float), ("last_build_ms", C.c_float),
                ("last_h2d_ms", C.c_float), ("last_d2h_ms", C.c_float), ("ncells", C.c_uint * 3),
                ("pbc_type", C.c_int), ("sweep_ms_sum", C.c_float), ("sweep_count", C.c_uint),
                ("build_ms_sum", C.c_float), ("build_count", C.c_uint), ("f32_search", C.c_int),
                ("super_builds", C.c_ulonglong), ("filter_rebuilds", C.c_ulonglong), ("build_ms_max", C.c_float)]


_lib = None


def lib():
    """load libb200coord.so (in-tree build) and declare the prototypes; raises if it is missing"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    L.b200coord_abi_version.restype = C.c_int
    L.b200coord_switch_parse.argtypes = [C.c_char_p, C.POINTER(Switch), C.c_char_p, C.c_size_t]
    L.b200coord_switch_rational.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Switch)]
    L.b200coord_switch_describe.argtypes = [C.POINTER(Switch), C.c_char_p, C.c_size_t]
    L.b200coord_pairing_dhenergy.argtypes = [C.c_double] * 6 + [C.POINTER(Switch)]
    L.b200coord_set_charges.argtypes = [C.c_void_p, dp]
    L.b200coord_pairing_ghbfix.argtypes = [C.c_double] * 3 + [C.POINTER(Switch)]
    L.b200coord_set_types.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_uint, dp]
    L.b200coord_create.argtypes = [C.POINTER(Config), C.POINTER(Switch), C.POINTER(C.c_uint), C.POINTER(C.c_void_p)]
    L.b200coord_destroy.argtypes = [C.c_void_p]
    L.b200coord_destroy.restype = None
    L.b200coord_last_error.argtypes = [C.c_void_p]
    L.b200coord_last_error.restype = C.c_char_p
    L.b200coord_set_box.argtypes = [C.c_void_p, dp]
    L.b200coord_prepare.argtypes = [C.c_void_p, C.c_long, C.c_int, C.POINTER(C.c_int)]
    L.b200coord_update_list.argtypes = [C.c_void_p, C.c_void_p]
    L.b200coord_calculate.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_void_p, dp]
    L.b200coord_calculate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200coord_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.b200coord_nl_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    L.b200coord_comm_unique_id.argtypes = [C.c_char_p]
    L.b200coord_comm_init.argtypes = [C.c_void_p, C.c_char_p]
    L.b200coord_coupling_publish.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.b200coord_coupling_withdraw.argtypes = [C.c_char_p]
    L.b200coord_coupling_lookup.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_size_t)]
    L.b200coord_coupled_set_index.argtypes = [C.c_void_p, C.c_void_p]
    L.b200coord_calculate_coupled.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b200coord_apply_coupled.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    L.b200coord_coupled_derivatives.argtypes = [C.c_void_p, C.c_void_p]
    L.b200coord_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200coord_collect.argtypes = [C.c_void_p]
    L.b200coord_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.b200coord_host_free.argtypes = [C.c_void_p]
    L.b200coord_device_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.b200coord_device_free.argtypes = [C.c_void_p]
    L.b200coord_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.b200coord_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.b200coord_device_synchronize.argtypes = []
    L.b200coord_enqueue_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200coord_stream_mark.argtypes = [C.c_void_p, C.c_int]
    L.b200coord_stream_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.b200coord_calculate_distributed.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_void_p, dp]
    L.b200coord_my_slice.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    L.b200coord_measure_fp64_peak.argtypes = [C.c_int, dp]
    L.b200coord_device_count.argtypes = [C.POINTER(C.c_int)]
    L.b200coord_enqueue_device_distributed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200coord_nl_pairs_device.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    L.b200coord_group_create.argtypes = [C.POINTER(Config), C.POINTER(Switch), C.POINTER(C.c_uint), C.POINTER(C.c_int), C.c_int,
                                         C.POINTER(C.c_void_p)]
    L.b200coord_group_destroy.argtypes = [C.c_void_p]
    L.b200coord_group_destroy.restype = None
    L.b200coord_group_size.argtypes = [C.c_void_p]
    L.b200coord_group_context.argtypes = [C.c_void_p, C.c_int]
    L.b200coord_group_context.restype = C.c_void_p
    L.b200coord_group_last_error.argtypes = [C.c_void_p]
    L.b200coord_group_last_error.restype = C.c_char_p
    L.b200coord_group_set_box.argtypes = [C.c_void_p, dp]
    L.b200coord_group_prepare.argtypes = [C.c_void_p, C.c_long, C.c_int, C.POINTER(C.c_int)]
    L.b200coord_group_set_charges.argtypes = [C.c_void_p, dp]
    L.b200coord_group_set_types.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_uint, dp]
    L.b200coord_group_calculate.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_void_p, dp]
    L.b200coord_peer_export.argtypes = [C.c_void_p, C.c_char_p]
    L.b200coord_peer_attach.argtypes = [C.c_void_p, C.c_char_p]
    if L.b200coord_abi_version() != ABI_VERSION:
        raise ImportError("libb200coord ABI version mismatch")
    _lib = L
    return L


def check(rc, ctx=None):
    if rc != OK:
        msg = lib().b200coord_last_error(ctx)
        raise B200CoordError(rc, msg.decode(errors="replace") if msg else "")


def switch_parse(definition):
    s = Switch()
    err = C.create_string_buffer(1024)
    rc = lib().b200coord_switch_parse(definition.encode(), C.byref(s), err, 1024)
    if rc != OK:
        raise B200CoordError(rc, err.value.decode(errors="replace"))
    return s


def switch_rational(nn, mm, r0, d0):
    s = Switch()
    check(lib().b200coord_switch_rational(int(nn), int(mm), float(r0), float(d0), C.byref(s)))
    return s


def pairing_dhenergy(ionic_strength, temp, epsilon, energy_unit=1.0, length_unit=1.0, charge_unit=1.0):
    s = Switch()
    check(lib().b200coord_pairing_dhenergy(float(ionic_strength), float(temp), float(epsilon), float(energy_unit),
                                           float(length_unit), float(charge_unit), C.byref(s)))
    return s


def pairing_ghbfix(dmax, d0, c):
    s = Switch()
    check(lib().b200coord_pairing_ghbfix(float(dmax), float(d0), float(c), C.byref(s)))
    return s


def switch_describe(s):
    buf = C.create_string_buffer(512)
    lib().b200coord_switch_describe(C.byref(s), buf, 512)
    return buf.value.decode()

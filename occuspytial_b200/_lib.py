"""ctypes binding of the engine's C ABI (include/occu_b200.h).

The shared library is built in-tree for sm_100a by ``__graft_entry__.build()`` /
``make -C occuspytial_b200/csrc``.  There is deliberately no fallback: if the library is missing
or a call fails, an exception is raised -- the samplers never run on a CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OCCU_B200_LIB') or os.path.join(_HERE, 'liboccu_b200.so')

ABI_VERSION = 2
LOGIT_RSR, PROBIT_RSR, LOGIT_ICAR = 0, 1, 2
SITE_NOT_SURVEYED, SITE_NOT_OBS, SITE_OBS = 0, 1, 2
F_ALPHA, F_BETA, F_ETA, F_TAU, F_Z, F_OMEGA, F_PSI, F_EPS, F_ZPROB = range(9)


class Problem(C.Structure):
    """occu_problem_t"""
    _fields_ = [
        ('abi_version', C.c_int32), ('model', C.c_int32), ('n', C.c_int32), ('p', C.c_int32),
        ('r', C.c_int32), ('q', C.c_int32), ('V', C.c_int32), ('chains', C.c_int32),
        ('chain_offset', C.c_int32), ('debug', C.c_int32), ('pg_fp64', C.c_int32), ('icar_bw', C.c_int32),
        ('gram_int8', C.c_int32), ('seed', C.c_uint64),
        ('X', C.c_void_p), ('K', C.c_void_p), ('Qrsr', C.c_void_p), ('Qband', C.c_void_p),
        ('W', C.c_void_p), ('y', C.c_void_p), ('visit_ptr', C.c_void_p), ('site_class', C.c_void_p),
        ('tau_rate', C.c_double), ('tau_shape', C.c_double),
        ('a_mu', C.c_void_p), ('a_prec', C.c_void_p), ('b_mu', C.c_void_p), ('b_prec', C.c_void_p),
    ]


class DebugOut(C.Structure):
    """occu_debug_out_t"""
    _fields_ = [(name, C.c_void_p) for name in (
        'eta_prec', 'eta_b', 'eta_chol', 'eta_mean', 'beta_prec', 'beta_b', 'alpha_prec', 'alpha_b', 'tau_rate')]


# every symbol include/occu_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'occu_last_error': (C.c_char_p, []),
    'occu_abi_version': (C.c_int, []),
    'occu_workspace_bytes': (C.c_size_t, [C.POINTER(Problem)]),
    'occu_create': (C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    'occu_destroy': (None, [C.c_void_p]),
    'occu_set_field': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'occu_get_field': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'occu_get_field_chain': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'occu_reset': (C.c_int, [C.c_void_p, C.c_void_p]),
    'occu_draw_start': (C.c_int, [C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_void_p]),
    'occu_sweep': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'occu_sweep_enqueue': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'occu_finish': (C.c_int, [C.c_void_p, C.c_void_p]),
    'occu_launch_count': (C.c_longlong, [C.c_void_p]),
    'occu_set_profiling': (C.c_int, [C.c_void_p, C.c_int]),
    'occu_get_kernel_ms': (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    'occu_debug_gaussian_half': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(DebugOut),
                                           C.c_void_p]),
    'occu_debug_site_half': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'occu_pg1': (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    'occu_truncnorm': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p,
                                 C.c_void_p]),
    'occu_gamma': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_void_p]),
    'occu_precision_mvnorm_batched': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                C.c_void_p, C.c_void_p]),
    'occu_ensure_sums_to_zero_batched': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'occu_gemm_tn': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_void_p]),
    'occu_chain_summaries': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


class EngineError(RuntimeError):
    """Raised when the CUDA engine is missing or reports an error."""


def load():
    """Load liboccu_b200.so (once) and attach the prototypes.  Raises EngineError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C occuspytial_b200/csrc`. There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.occu_abi_version() != ABI_VERSION:
        raise EngineError('liboccu_b200.so ABI version mismatch: rebuild the library')
    _lib = lib
    return lib


def check(code, what):
    """Negative return codes are hard errors; positive ones are per-chain numerical failures."""
    if code < 0:
        msg = load().occu_last_error().decode('utf-8', 'replace')
        raise EngineError(f'{what}: {msg}')
    return code

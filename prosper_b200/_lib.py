"""ctypes binding of include/prosper_b200.h (the C ABI is the product boundary).

Loading fails loudly when the library has not been built -- there is no fallback path.
"""
import ctypes as C
import os

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)

# model kinds / flags (keep in sync with the header)
MODEL_BSC, MODEL_MCA, MODEL_MMCA, MODEL_TSC, MODEL_DSC, MODEL_GSC = range(6)
PASS_SELECT = 1
PASS_REUSE_SCORES = 2
PASS_DEFER_STATS = 4
N_STAGES = 10          # PET_N_STAGES
MAX_HPRIME, MAX_GAMMA = 16, 8     # engine limits (gl_kernel.cuh)


class PetError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "prosper_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("device", C.c_int32),
                ("D", C.c_int64), ("H", C.c_int64), ("Hprime", C.c_int64), ("gamma", C.c_int64),
                ("n_states", C.c_int32), ("states", c_double_p), ("chunk_rows", C.c_int64)]


class Params(C.Structure):
    _fields_ = [("W", C.c_void_p), ("ldW", C.c_int64), ("pi_host", c_double_p), ("n_pi", C.c_int32),
                ("sigma", C.c_double), ("mu", C.c_void_p)]


class Anneal(C.Structure):
    _fields_ = [("T", C.c_double), ("Ncut_factor", C.c_double), ("anneal_prior", C.c_int32)]


class StatsLayout(C.Structure):
    _fields_ = [("total", C.c_int64),
                ("off_Wp", C.c_int64), ("rows_Wp", C.c_int64), ("cols_Wp", C.c_int64), ("ld_Wp", C.c_int64),
                ("off_Wq", C.c_int64), ("rows_Wq", C.c_int64), ("cols_Wq", C.c_int64), ("ld_Wq", C.c_int64),
                ("off_scalars", C.c_int64), ("n_scalars", C.c_int64)]


class GSCParams(C.Structure):
    _fields_ = [("W", C.c_void_p), ("ldW", C.c_int64), ("pi_host", c_double_p), ("mu_host", c_double_p),
                ("psi_sq_host", c_double_p), ("sigma_sq_host", c_double_p), ("sigma_sq_type", C.c_int32)]


class GSCLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("total", "ld", "off_A", "off_Mssz", "off_Mout", "off_ss", "off_szsz",
                                        "off_sum_s", "off_sum_sz2", "off_ysq", "off_scalars", "off_yyT", "ld_yyT")]


# every symbol the header declares: name -> (restype, argtypes)
SIGNATURES = {
    "pet_abi_version": (C.c_int, []),
    "pet_last_error": (C.c_char_p, []),
    "pet_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "pet_destroy": (None, [C.c_void_p]),
    "pet_num_states": (C.c_int64, [C.c_void_p]),
    "pet_num_columns": (C.c_int64, [C.c_void_p]),
    "pet_state_matrix": (C.c_int, [C.c_void_p, c_double_p]),
    "pet_set_data": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "pet_set_chunk_target": (C.c_int, [C.c_void_p, C.c_int64]),
    "pet_num_data": (C.c_int64, [C.c_void_p]),
    "pet_select_hprimes": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p]),
    "pet_set_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pet_e_step": (C.c_int, [C.c_void_p, C.POINTER(Anneal), C.POINTER(Params), C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_posterior_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pet_log_denominators": (C.c_int, [C.c_void_p, C.POINTER(Anneal), C.POINTER(Params), C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "pet_log_denominators_ptr": (C.c_void_p, [C.c_void_p]),
    "pet_kth_largest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "pet_m_step_stats": (C.c_int, [C.c_void_p, C.POINTER(Anneal), C.POINTER(Params), C.c_void_p, C.c_int64,
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pet_stats_layout_get": (C.c_int, [C.c_void_p, C.POINTER(StatsLayout)]),
    "pet_m_step_solve": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int32), C.c_void_p]),
    "pet_gsc_layout_get": (C.c_int, [C.c_void_p, C.POINTER(GSCLayout)]),
    "pet_gsc_select": (C.c_int, [C.c_void_p, C.POINTER(GSCParams), C.c_void_p, C.c_void_p]),
    "pet_gsc_compute_lpj": (C.c_int, [C.c_void_p, C.POINTER(GSCParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pet_gsc_e_step": (C.c_int, [C.c_void_p, C.POINTER(Anneal), C.POINTER(GSCParams), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pet_gsc_stats": (C.c_int, [C.c_void_p, C.POINTER(Anneal), C.POINTER(GSCParams), C.c_int32, C.c_void_p, C.c_void_p]),
    "pet_generate_data": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                    c_double_p, c_double_p, C.c_double, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_int64, C.c_void_p]),
    "pet_normal_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_double, C.c_uint64, C.c_void_p]),
    "pet_gather_rows": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_col_centered_sumsq": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pet_data_sum": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pet_mix_posterior": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p,
                                    C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_rowdot": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_rowop": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double,
                            C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_colsum": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pet_dgemm_kk": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                               C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p]),
    "pet_dgemm_mn": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                               C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "pet_ozaki_gemm_kk": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pet_ozaki_gemm_mn": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pet_ozaki_last_ms": (C.c_double, []),
    "pet_spd_solve_right": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "pet_spd_solve_work_doubles": (C.c_int64, [C.c_int64, C.c_int64]),
    "pet_gemm_path": (C.c_int32, [C.c_void_p]),
    "pet_gemm_slices": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pet_set_state_kernel": (C.c_int, [C.c_void_p, C.c_int32]),
    "pet_state_kernel_path": (C.c_int32, [C.c_void_p]),
    "pet_stage_times_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "pet_enable_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "pet_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


def library_path():
    return _build.LIBPATH


def load():
    """Return the loaded CDLL (cached); raises if the library was never built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("prosper_b200: %s is missing -- run `python -m prosper_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PetError(rc, (load().pet_last_error() or b"").decode("utf-8", "replace"))
    return rc

"""ctypes binding of libgpdoemd_b200.so (include/gpdoemd_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GPDOEMD_B200_LIB') or os.path.join(_HERE, 'libgpdoemd_b200.so')   # env: A/B builds

KERNEL_IDS = {'RBF': 0, 'Exponential': 1, 'Matern32': 2, 'Matern52': 3, 'Cosine': 4, 'RatQuad': 5}
CRITERION_IDS = {'HR': 0, 'BH': 1, 'BF': 2, 'AW': 3, 'JR': 4}
FAMILIES = ('eval', 'tri', 'gram', 'marginal', 'criteria', 'argmax', 'setup', 'tri_head', 'comm')
NFAMILIES = len(FAMILIES)
NCCL_ID_BYTES = 128

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_void_pp = C.POINTER(C.c_void_p)


class GpDesc(C.Structure):
    _fields_ = [('kernel', C.c_int32), ('n', C.c_int32), ('dim_x', C.c_int32), ('dim_p', C.c_int32),
                ('n_x_active', C.c_int32), ('x_active', c_int32_p), ('var_x', C.c_double), ('ls_x', c_double_p),
                ('power_x', C.c_double), ('var_p', C.c_double), ('ls_p', c_double_p), ('power_p', C.c_double),
                ('Z', c_double_p), ('alpha', c_double_p), ('L', c_double_p), ('factor_kind', C.c_int32)]


class GpdError(RuntimeError):
    pass


_SIGNATURES = {
    # name: (restype, argtypes)
    'gpd_create': (C.c_int, [C.c_int, c_void_pp]),
    'gpd_destroy': (None, [C.c_void_p]),
    'gpd_last_error': (C.c_char_p, []),
    'gpd_set_scratch_bytes': (C.c_int, [C.c_void_p, C.c_uint64]),
    'gpd_set_option': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    'gpd_launch_count': (C.c_uint64, [C.c_void_p]),
    'gpd_profile_enable': (C.c_int, [C.c_void_p, C.c_int]),
    'gpd_profile_read': (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_uint64), c_double_p]),
    'gpd_timer_start': (C.c_int, [C.c_void_p]),
    'gpd_timer_stop': (C.c_int, [C.c_void_p, c_double_p]),
    'gpd_sync': (C.c_int, [C.c_void_p]),
    'gpd_probe_fp64_peak': (C.c_int, [C.c_void_p, c_double_p]),
    'gpd_flush_l2': (C.c_int, [C.c_void_p]),
    'gpd_gp_upload': (C.c_int, [C.c_void_p, C.POINTER(GpDesc), c_void_pp]),
    'gpd_gp_build': (C.c_int, [C.c_void_p, C.POINTER(GpDesc), c_double_p, C.c_double, c_void_pp, c_double_p, c_double_p,
                               c_double_p]),
    'gpd_gp_free': (None, [C.c_void_p]),
    'gpd_model_create': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_void_pp,
                                   c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_void_pp]),
    'gpd_model_free': (None, [C.c_void_p]),
    'gpd_model_set_pmean_sigma': (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    'gpd_model_set_pstar': (C.c_int, [C.c_void_p, c_double_p]),
    'gpd_predict': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    'gpd_dmu_dp': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p]),
    'gpd_d2mu_dp2': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p]),
    'gpd_ds2_dp': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p]),
    'gpd_d2s2_dp2': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p]),
    'gpd_hooks': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int32] + [c_double_p] * 6),
    'gpd_marginal': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int, c_double_p, c_double_p]),
    'gpd_marginal_assemble': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int32] + [c_double_p] * 7),
    'gpd_criterion': (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_int32,
                                c_double_p, c_double_p, c_double_p]),
    'gpd_gauss_loglik': (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_int32,
                                   c_double_p, c_double_p]),
    'gpd_argmax': (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.POINTER(C.c_int64)]),
    'gpd_score': (C.c_int, [C.c_void_p, c_void_pp, C.c_int32, c_double_p, C.c_int64, C.c_int, C.c_int, c_double_p,
                            c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int64), C.c_int64]),
    'gpd_score_dev': (C.c_int, [C.c_void_p, c_void_pp, C.c_int32, C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p,
                                c_double_p, C.c_void_p, c_double_p, C.POINTER(C.c_int64), C.c_int64]),
    'gpd_score_multi': (C.c_int, [C.c_void_p, c_void_pp, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int, c_int32_p,
                                  C.c_int32, c_double_p, c_double_p, C.c_void_p, C.c_int32, c_double_p,
                                  C.POINTER(C.c_int64), C.c_int64]),
    'gpd_marginal_dx': (C.c_int, [C.c_void_p, c_double_p, C.c_int64] + [c_double_p] * 4),
    'gpd_marginals': (C.c_int, [C.c_void_p, c_void_pp, C.c_int32, c_double_p, C.c_int64, C.c_int, c_double_p, c_double_p,
                                c_void_pp]),
    'gpd_criterion_resident': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_double_p, c_double_p, c_double_p]),
    'gpd_resident_free': (None, [C.c_void_p]),
    'gpd_nccl_version': (C.c_int, [c_int32_p]),
    'gpd_nccl_unique_id': (C.c_int, [C.c_char_p]),
    'gpd_comm_init': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, c_void_pp]),
    'gpd_comm_adopt': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, c_void_pp]),
    'gpd_comm_destroy': (None, [C.c_void_p]),
    'gpd_score_sharded': (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int,
                                    c_int32_p, C.c_int32, c_double_p, c_double_p, C.c_void_p, C.c_int32, c_double_p,
                                    C.POINTER(C.c_int64), C.c_int64]),
    'gpd_dev_alloc': (C.c_int, [C.c_void_p, C.c_uint64, c_void_pp]),
    'gpd_dev_free': (C.c_int, [C.c_void_p, C.c_void_p]),
    'gpd_dev_upload': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    'gpd_dev_download': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    'gpd_host_alloc_pinned': (C.c_int, [C.c_void_p, C.c_uint64, c_void_pp]),
    'gpd_host_free_pinned': (C.c_int, [C.c_void_p, C.c_void_p]),
    'gpd_dev_fill_uniform': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_int64,
                                       c_double_p, c_double_p]),
    'gpd_debug_check_scratch': (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'gpd_layout_a_index': (C.c_int32, [C.c_int32] * 3),
    'gpd_layout_a_row': (C.c_int32, [C.c_int32] * 2),
    'gpd_layout_a_col': (C.c_int32, [C.c_int32] * 2),
    'gpd_layout_b_index': (C.c_int32, [C.c_int32] * 2),
    'gpd_layout_tile_offset': (C.c_int64, [C.c_int32] * 4),
    'gpd_binary_combo_weight': (C.c_int32, [C.c_int32] * 2),
}

_lib = None


def _preload_nccl():
    """One NCCL per process.  The C library binds NCCL with dlopen("libnccl.so.2") on first use; if that picked the
    system copy and PyTorch (whose libtorch_cuda.so needs the newer NCCL bundled in the `nvidia-nccl-cu12` wheel) were
    imported afterwards, the loader would hand torch the older, already loaded soname.  So the wheel's copy, when
    there is one, is loaded first; a process that already holds a libnccl.so.2 keeps it."""
    try:
        import importlib.util
        spec = importlib.util.find_spec('nvidia.nccl')
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, 'lib', 'libnccl.so.2')
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return cand
    except Exception:
        pass
    return None


def load():
    """Load the shared library and declare every prototype of include/gpdoemd_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    _preload_nccl()
    if not os.path.exists(LIB_PATH):
        raise GpdError('%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       '(or `make -C gpdoemd_b200/csrc`); there is no CPU fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


ERR_SINGULAR = 6


def check(rc):
    if rc == ERR_SINGULAR:      # the reference's np.linalg.inv raises this for a singular covariance
        raise np.linalg.LinAlgError(load().gpd_last_error().decode('utf-8', 'replace'))
    if rc != 0:
        raise GpdError(load().gpd_last_error().decode('utf-8', 'replace'))


def dptr(a):
    """float64 C-contiguous array -> double* (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_double_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)

"""ctypes binding of libpmgrav.so — the C ABI declared in include/pmgrav.h.

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libpmgrav.so')

PM_GRID_F64, PM_GRID_F32 = 0, 1
PM_TAP_REAL, PM_TAP_FOURIER, PM_TAP_FORCE = 0, 1, 2


class PMError(RuntimeError):
    """A libpmgrav call returned a negative pm_status (the reference would abort(), commons.py:1002)."""

    def __init__(self, status, message):
        super().__init__(f'libpmgrav error {status}: {message}')
        self.status = status


class KickParams(Structure):
    """struct pm_kick_params"""
    _fields_ = [
        ('order', c_int), ('diff_order', c_int), ('deconv_order', c_int), ('interlace', c_int),
        ('contribution', c_double), ('prefactor', c_double), ('gauss', c_double), ('kick_factor', c_double),
    ]


# name: (restype, argtypes).  Every symbol of include/pmgrav.h is listed; tests check the set.
SIGNATURES = {
    'pm_version': (c_char_p, []),
    'pm_last_error': (c_char_p, []),
    'pm_launch_count': (c_int64, []),
    'pm_device_count': (c_int, []),
    'pm_create': (c_int, [POINTER(c_void_p), c_int, c_double, c_int, c_int, c_int, c_int, c_void_p]),
    'pm_destroy': (c_int, [c_void_p]),
    'pm_set_stream': (c_int, [c_void_p, c_void_p]),
    'pm_sync': (c_int, [c_void_p]),
    'pm_local_shape': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    'pm_device_bytes': (c_int64, [c_void_p]),
    'pm_comm_unique_id': (c_int, [c_void_p]),
    'pm_comm_init': (c_int, [c_void_p, c_void_p]),
    'pm_allreduce_sum': (c_int, [c_void_p, c_void_p, c_int]),
    'pm_ipc_get_handle': (c_int, [c_void_p, c_void_p]),
    'pm_ipc_open_peers': (c_int, [c_void_p, c_void_p]),
    'pm_grid_zero': (c_int, [c_void_p]),
    'pm_deposit': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_double, POINTER(c_double)]),
    'pm_halo_add': (c_int, [c_void_p]),
    'pm_halo_fill': (c_int, [c_void_p]),
    'pm_halo_fill_for': (c_int, [c_void_p, c_int, c_int, c_int]),
    'pm_fft_forward': (c_int, [c_void_p]),
    'pm_fft_backward': (c_int, [c_void_p]),
    'pm_kspace_potential': (c_int, [c_void_p, c_double, c_int, c_double, c_double]),
    'pm_fourier_operate': (c_int, [c_void_p, c_int, POINTER(c_double), c_double, c_int, c_int]),
    'pm_solve_fused': (c_int, [c_void_p, c_double, c_int, c_double]),
    'pm_solve_fused_stage': (c_int, [c_void_p, c_double, c_int, c_double, c_int]),
    'pm_fused_solve_available': (c_int, [c_void_p]),
    'pm_set_fused_solve': (c_int, [c_void_p, c_int]),
    'pm_check_async_error': (c_int, [c_void_p]),
    'pm_power_k2': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'pm_slab_save': (c_int, [c_void_p]),
    'pm_slab_accumulate': (c_int, [c_void_p]),
    'pm_slab_restore': (c_int, [c_void_p]),
    'pm_diff': (c_int, [c_void_p, c_int, c_int]),
    'pm_gather': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_double, POINTER(c_double)]),
    'pm_gather_kick': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_double, POINTER(c_double), c_void_p]),
    'pm_gather_kick_drift': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_double, POINTER(c_double), c_void_p, c_double]),
    'pm_drift': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double]),
    'pm_sum_mom2': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'pm_sort_particles': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    'pm_exchange': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64]),
    'pm_exchange_rungs': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64]),
    'pm_shortrange': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, POINTER(c_double), c_int, c_double,
                              c_void_p, c_int, c_double, c_void_p]),
    'pm_shortrange_stats': (c_int, [c_void_p, c_int, POINTER(c_int64), POINTER(c_int64)]),
    'pm_apply_dmom': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, POINTER(c_double), c_int, c_int]),
    'pm_assign_rungs': (c_int, [c_void_p, c_void_p, c_int64, c_double, c_int, c_void_p, c_void_p, POINTER(c_int64)]),
    'pm_flag_rung_jumps': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_double, c_double,
                                   POINTER(c_double), c_int, POINTER(c_int)]),
    'pm_apply_rung_jumps': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, POINTER(c_int64)]),
    'pm_ic_lattice': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int64, c_int64, POINTER(c_int64)]),
    'pm_ic_potential': (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(c_double), c_double]),
    'pm_ic_nongaussianity': (c_int, [c_void_p, c_double]),
    'pm_ic_displace': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_double, c_double]),
    'pm_ic_wrap': (c_int, [c_void_p, c_void_p, c_int64]),
    'pm_real_export': (c_int, [c_void_p, c_void_p]),
    'pm_ic_2lpt_source': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pm_lpt_accumulate': (c_int, [c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_int]),
    'pm_real_import': (c_int, [c_void_p, c_void_p]),
    'pm_fourier_resize': (c_int, [c_void_p, c_void_p]),
    'pm_fourier_copy_modes': (c_int, [c_void_p, c_void_p, c_int, POINTER(c_double), c_double, c_int, c_int, c_int]),
    'pm_kick_long': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(KickParams), c_void_p]),
    'pm_kick_drift': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(KickParams), c_double, c_void_p]),
    'pm_kick_long_host': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(KickParams), c_double, POINTER(c_double)]),
    'pm_tap_size': (c_int64, [c_void_p, c_int]),
    'pm_get_grid': (c_int, [c_void_p, c_int, c_void_p]),
    'pm_set_grid': (c_int, [c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libpmgrav.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing. Build it with `python -m concept_b200.build` '
            f'(or __graft_entry__.build()); concept_b200 has no CPU fallback.'
        )
    # If PyTorch is around, import it first so that its bundled libnccl.so.2 / libcufft are the ones
    # already mapped when our library resolves the same sonames.
    try:
        import torch  # noqa: F401
    except Exception:
        pass
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status < 0:
        raise PMError(status, load().pm_last_error().decode('utf-8', 'replace'))
    return status


def vec3(shift):
    """None or a length-3 sequence → POINTER(c_double)"""
    if shift is None:
        return None
    return (c_double*3)(*[float(s) for s in shift])

"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/pm_oracle.c (the C/OpenMP restatement).
Used by tests/ and by bench.py's cpu_baseline / --impl reference legs; never by concept_b200/."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libpm_oracle.so')
_lib = None
_dp = ctypes.c_void_p


def build():
    subprocess.run(['make', '-C', HERE], check=True, stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        lib = ctypes.CDLL(LIB)
        lib.pmo_num_threads.restype = ctypes.c_int
        lib.pmo_sum_mom2.restype = ctypes.c_double
        lib.pmo_sum_mom2.argtypes = [_dp, ctypes.c_int64]
        lib.pmo_deposit.argtypes = [_dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, _dp, _dp]
        lib.pmo_fft_forward.argtypes = [_dp, ctypes.c_int, _dp]
        lib.pmo_fft_backward.argtypes = [_dp, ctypes.c_int, _dp]
        lib.pmo_kspace_potential.argtypes = [_dp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        lib.pmo_diff.argtypes = [_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp]
        lib.pmo_gather.argtypes = [_dp, ctypes.c_int, ctypes.c_double, _dp, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp]
        lib.pmo_drift.argtypes = [_dp, _dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double]
        lib.pmo_kick_long.argtypes = [_dp, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp, _dp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def num_threads():
    return load().pmo_num_threads()


def set_num_threads(n=None):
    """Use n OpenMP threads (default: every host core), whatever OMP_NUM_THREADS says — torchrun sets it to 1."""
    load().pmo_set_num_threads(int(n or os.cpu_count() or 1))
    return num_threads()


class Workspace:
    """Scratch grids for pmo_kick_long (allocated once, like the reference's cached buffers)."""
    def __init__(self, G):
        self.G = G
        self.real = np.empty(2*G**3, dtype=np.float64)
        self.slab = np.empty(G*G*(G//2 + 1), dtype=np.complex128)


def kick_long(pos, mom, *, mass, boxsize, gridsize, order, G_Newton, dt_rho_over_dt1, dt_kick, diff_order=2,
              deconvolve=True, r_scale=0.0, work=None):
    """In-place on mom (C-contiguous float64 (N,3)).  Same scalars as concept_b200.pmsolver.make_kick_params."""
    G = int(gridsize)
    work = work or Workspace(G)
    contribution = dt_rho_over_dt1
    contribution *= mass
    contribution *= float(G)**(-3)*(G/boxsize)**3
    s = load().pmo_kick_long(_p(pos), _p(mom), pos.shape[0], G, boxsize, order, diff_order, order*(2 if deconvolve else 0),
                             contribution, -boxsize**2*G_Newton/np.pi, (2*np.pi/boxsize*r_scale)**2 if r_scale else 0.0,
                             mass*(-dt_kick), _p(work.real), _p(work.slab))
    if s:
        raise ValueError(f'pmo_kick_long returned {s}')
    return mom


def drift(pos, mom, dt_over_mass, boxsize):
    load().pmo_drift(_p(pos), _p(mom), pos.shape[0], dt_over_mass, boxsize)
    return pos


def sum_mom2(mom):
    return load().pmo_sum_mom2(_p(mom), mom.shape[0])


def deposit(pos, G, boxsize, order, contribution, shift=None):
    rho = np.zeros((G, G, G))
    sh = None if shift is None else np.asarray(shift, dtype=np.float64)
    load().pmo_deposit(_p(pos), pos.shape[0], G, boxsize, order, contribution, None if sh is None else _p(sh), _p(rho))
    return rho

"""PMContext — Python handle on one libpmgrav context (one grid size, one GPU/slab).

It is the analogue of the reference's cached FFTW slab + plans (get_fftw_slab, mesh.py:3769-3866)
plus the named scratch grids ('grid_updownstream', 'slab_global', 'force_downstream';
communication.py:1666-1723).  All heavy lifting happens inside libpmgrav.so; particle arrays are
torch CUDA tensors (float64, shape (N, 3), contiguous — the AoS layout of Component.pos/.mom,
species.py:1411-1431) and only their device pointers cross the C ABI.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import KickParams, PM_GRID_F32, PM_GRID_F64, PM_TAP_FORCE, PM_TAP_FOURIER, PM_TAP_REAL, check, vec3

BCC_SHIFT = (-0.5, -0.5, -0.5)   # Lattice.shift_amount for cell-centred grids (mesh.py:85)


def _ptr(t, n=None):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise TypeError('expected a contiguous CUDA tensor')
    return ctypes.c_void_p(t.data_ptr())


def _particles(t):
    if t.dtype != torch.float64 or t.dim() != 2 or t.shape[1] != 3:
        raise TypeError('particle arrays must be float64 of shape (N, 3)')
    return _ptr(t)


class PMContext:
    def __init__(self, gridsize, boxsize, dtype='f64', rank=0, nranks=1, device=None, stream=None):
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.torch_device = torch.device('cuda', self.device)
        self.gridsize = int(gridsize)
        self.boxsize = float(boxsize)
        self.dtype = {'f64': PM_GRID_F64, 'f32': PM_GRID_F32, 'float64': PM_GRID_F64, 'float32': PM_GRID_F32}[str(dtype)]
        self.rank, self.nranks = int(rank), int(nranks)
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self._h = ctypes.c_void_p()
        check(self.lib.pm_create(ctypes.byref(self._h), self.gridsize, self.boxsize, self.dtype,
                                 self.rank, self.nranks, self.device, ctypes.c_void_p(stream)))
        nx, x0, nj, j0 = (ctypes.c_int64() for _ in range(4))
        check(self.lib.pm_local_shape(self._h, *(ctypes.byref(v) for v in (nx, x0, nj, j0))))
        self.nx_local, self.x_start, self.nj_local, self.j_start = nx.value, x0.value, nj.value, j0.value
        self._sum = torch.zeros(8, dtype=torch.float64, device=f'cuda:{self.device}')

    # -- life cycle ----------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.pm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        check(self.lib.pm_set_stream(self._h, ctypes.c_void_p(stream)))

    def sync(self):
        check(self.lib.pm_sync(self._h))

    @property
    def device_bytes(self):
        return self.lib.pm_device_bytes(self._h)

    # -- communicator ----------------------------------------------------------------
    def comm_init(self, unique_id_bytes):
        buf = ctypes.create_string_buffer(bytes(unique_id_bytes), 128)
        check(self.lib.pm_comm_init(self._h, buf))

    @staticmethod
    def comm_unique_id():
        lib = _lib.load()
        buf = ctypes.create_string_buffer(128)
        check(lib.pm_comm_unique_id(buf))
        return buf.raw

    def connect(self, bcast, allgather, is_master):
        """Join the ranks of one job: NCCL communicator (id broadcast from the master) and CUDA-IPC
        peer mappings of every rank's slab (handles all-gathered in rank order).  `bcast(obj)` and
        `allgather(obj)` are host-side helpers (torch.distributed object collectives)."""
        uid = bcast(PMContext.comm_unique_id() if is_master else None)
        self.comm_init(uid)
        self.ipc_open_peers(allgather(self.ipc_handle()))

    def ipc_handle(self):
        buf = ctypes.create_string_buffer(64)
        check(self.lib.pm_ipc_get_handle(self._h, buf))
        return buf.raw

    def ipc_open_peers(self, handles):
        """handles: list of 64-byte handles in rank order (all-gathered by the host code)"""
        blob = b''.join(bytes(h) for h in handles)
        check(self.lib.pm_ipc_open_peers(self._h, ctypes.create_string_buffer(blob, len(blob))))

    def allreduce_sum(self, t):
        check(self.lib.pm_allreduce_sum(self._h, _ptr(t), t.numel()))

    # -- mesh operators ----------------------------------------------------------------
    def grid_zero(self):
        check(self.lib.pm_grid_zero(self._h))

    def deposit(self, pos, order, contribution, shift=None):
        check(self.lib.pm_deposit(self._h, _particles(pos), pos.shape[0], int(order), float(contribution), vec3(shift)))

    def halo_add(self):
        check(self.lib.pm_halo_add(self._h))

    def halo_fill(self):
        check(self.lib.pm_halo_fill(self._h))

    def halo_fill_for(self, order, diff_order, interlace=False):
        check(self.lib.pm_halo_fill_for(self._h, int(order), int(diff_order), int(bool(interlace))))

    def fft_forward(self):
        check(self.lib.pm_fft_forward(self._h))

    def fft_backward(self):
        check(self.lib.pm_fft_backward(self._h))

    def kspace_potential(self, prefactor, deconv_order, gauss=0.0, scale=1.0):
        check(self.lib.pm_kspace_potential(self._h, float(prefactor), int(deconv_order), float(gauss), float(scale)))

    def solve_fused(self, prefactor, deconv_order, gauss=0.0):
        check(self.lib.pm_solve_fused(self._h, float(prefactor), int(deconv_order), float(gauss)))

    def solve_fused_stage(self, prefactor, deconv_order, gauss, stage):
        check(self.lib.pm_solve_fused_stage(self._h, float(prefactor), int(deconv_order), float(gauss), int(stage)))

    @property
    def hand_fft_available(self):
        """True when pm_solve_fused runs the hand-written slab transform (G ∈ {128, 256, 512, 1024})."""
        return self.gridsize in (128, 256, 512, 1024) and self.fused_solve_available

    @property
    def fused_solve_available(self):
        return bool(self.lib.pm_fused_solve_available(self._h))

    # PM_SOLVE_* of include/pmgrav.h
    SOLVE_MODES = {'unfused': 0, 'auto': 1, 'cufft2d': 2, 'fft2_split': 3, 'fft2_l2': 4}

    def set_fused_solve(self, mode):
        """mode: bool (False = three-call path, True = best fused implementation) or a SOLVE_MODES key."""
        if isinstance(mode, str):
            mode = self.SOLVE_MODES[mode]
        check(self.lib.pm_set_fused_solve(self._h, int(mode)))

    def check_async_error(self):
        check(self.lib.pm_check_async_error(self._h))

    def fourier_operate(self, deconv_order=0, shift=None, scale=1.0, diff_dim=-1, from_saved=False):
        check(self.lib.pm_fourier_operate(self._h, int(deconv_order), vec3(shift), float(scale), int(diff_dim), int(from_saved)))

    def power_k2(self, k2_max, power, count=None):
        """power[k²] += |mode|² (and count[k²] += 1) over the sparse half-space; device tensors of k2_max + 1 entries
        (float64 / int64)."""
        check(self.lib.pm_power_k2(self._h, int(k2_max), _ptr(power), _ptr(count)))

    def slab_save(self):
        check(self.lib.pm_slab_save(self._h))

    def slab_accumulate(self):
        check(self.lib.pm_slab_accumulate(self._h))

    def slab_restore(self):
        check(self.lib.pm_slab_restore(self._h))

    def diff(self, dim, order):
        check(self.lib.pm_diff(self._h, int(dim), int(order)))

    def gather(self, which, pos, mom, order, dim, factor, shift=None):
        check(self.lib.pm_gather(self._h, int(which), _particles(pos), _particles(mom), pos.shape[0], int(order),
                                 int(dim), float(factor), vec3(shift)))

    def gather_kick(self, pos, mom, order, diff_order, factor, shift=None, sum_mom2=None):
        check(self.lib.pm_gather_kick(self._h, _particles(pos), _particles(mom), pos.shape[0], int(order),
                                      int(diff_order), float(factor), vec3(shift), _ptr(sum_mom2)))

    def gather_kick_drift(self, pos, mom, order, diff_order, factor, dt_over_mass, shift=None, sum_mom2=None):
        check(self.lib.pm_gather_kick_drift(self._h, _particles(pos), _particles(mom), pos.shape[0], int(order),
                                            int(diff_order), float(factor), vec3(shift), _ptr(sum_mom2), float(dt_over_mass)))

    # -- particle operators ------------------------------------------------------------
    def drift(self, pos, mom, dt_over_mass):
        check(self.lib.pm_drift(self._h, _particles(pos), _particles(mom), pos.shape[0], float(dt_over_mass)))

    def sum_mom2(self, mom, out=None):
        """Σ mom² over the local particles → python float (synchronises)."""
        acc = self._sum[:1] if out is None else out
        if out is None:
            acc.zero_()
        check(self.lib.pm_sum_mom2(self._h, _particles(mom), mom.shape[0], _ptr(acc)))
        if out is not None:
            return None
        value = float(acc.item())
        # the host synchronises here anyway (once per base step through measure('v_rms')): a good place to notice a
        # hand-written transform that gave up on a tile dependency and left an invalid potential behind
        self.check_async_error()
        return value

    def sort_particles(self, pos, mom, ids=None, n=None):
        """Stable reorder of the first n particles by grid cell (tile_sort analogue)."""
        n = pos.shape[0] if n is None else int(n)
        check(self.lib.pm_sort_particles(self._h, _particles(pos), _particles(mom), _ptr(ids), n))

    def exchange(self, pos, mom, ids, n, Δmom=None, rung_indices=None, rung_indices_jumped=None):
        """Slab migration.  pos/mom(/ids) are capacity-sized buffers; returns the new local count.  The P³M state of a
        component (Δmom, rung_indices, rung_indices_jumped) migrates along when given."""
        n_io = ctypes.c_int64(int(n))
        if Δmom is None and rung_indices is None:
            check(self.lib.pm_exchange(self._h, _particles(pos), _particles(mom), _ptr(ids), ctypes.byref(n_io), pos.shape[0]))
        else:
            capacity = min(t.shape[0] for t in (pos, mom, Δmom, rung_indices, rung_indices_jumped) if t is not None)
            check(self.lib.pm_exchange_rungs(self._h, _particles(pos), _particles(mom), _ptr(ids), _ptr(Δmom), _ptr(rung_indices),
                                             _ptr(rung_indices_jumped), ctypes.byref(n_io), capacity))
        return n_io.value

    # -- initial conditions (csrc/pm_ic.cu) ------------------------------------------------
    def ic_lattice(self, pos, mom, ids, shift, index_bgn, id_bgn):
        """preinitialize_particles for this rank's slab of the gridsize³ lattice; returns the local count."""
        n_local = ctypes.c_int64(0)
        check(self.lib.pm_ic_lattice(self._h, _particles(pos), _particles(mom), _ptr(ids), vec3(shift), int(index_bgn),
                                     int(id_bgn), ctypes.byref(n_local)))
        return n_local.value

    def ic_potential(self, noise, amplitudes, k2_max, shift=None, lap_factor=1.0):
        """working slab = amplitudes[k²]·noise·phase(shift)·(−lap_factor/k_f²)/k²  (realize_grid + laplacian_inverse)"""
        if noise.dtype != torch.float64 or noise.numel() != 2*self.gridsize*self.nj_local*(self.gridsize//2 + 1):
            raise TypeError('noise must hold the float64 (re, im) pairs of the local Fourier slab')
        if amplitudes.dtype != torch.float64 or amplitudes.numel() < k2_max + 1:
            raise TypeError('amplitudes must be a float64 table with k2_max + 1 entries')
        check(self.lib.pm_ic_potential(self._h, _ptr(noise), _ptr(amplitudes), int(k2_max), vec3(shift), float(lap_factor)))

    def ic_nongaussianity(self, f_nl):
        """real-space grid: x += f_nl·x²  (realize_grid, ic.py:766-771)"""
        check(self.lib.pm_ic_nongaussianity(self._h, float(f_nl)))

    def ic_displace(self, pos, mom, index_bgn, dim, pos_factor=1.0, mom_factor=0.0):
        check(self.lib.pm_ic_displace(self._h, None if pos is None else _particles(pos), None if mom is None else _particles(mom),
                                      int(index_bgn), int(dim), float(pos_factor), float(mom_factor)))

    def ic_wrap(self, pos, n):
        check(self.lib.pm_ic_wrap(self._h, _particles(pos), int(n)))

    def real_export(self, out=None):
        """The real-space grid without padding as a device tensor (nx_local, G, G)."""
        if out is None:
            out = torch.empty((self.nx_local, self.gridsize, self.gridsize), dtype=torch.float64, device=self.torch_device)
        check(self.lib.pm_real_export(self._h, _ptr(out)))
        return out

    def ic_2lpt_source(self, d00, d11, d22, d01, d12, d02):
        n = self.nx_local*self.gridsize**2
        for t in (d00, d11, d22, d01, d12, d02):
            if t.dtype != torch.float64 or t.numel() != n:
                raise TypeError('second-derivative grids must be float64 of shape (nx_local, G, G)')
        check(self.lib.pm_ic_2lpt_source(self._h, *(_ptr(t) for t in (d00, d11, d22, d01, d12, d02))))

    def lpt_accumulate(self, acc, factor, a, b, third=None, assign=False):
        """acc (= | +=) factor·a·b[·third] on compact device grids (one term of an LPT source)"""
        n = acc.numel()
        for t in (acc, a, b) + ((third, ) if third is not None else ()):
            if t.dtype != torch.float64 or t.numel() != n:
                raise TypeError('LPT grids must be float64 tensors of equal size')
        check(self.lib.pm_lpt_accumulate(self._h, _ptr(acc), n, float(factor), _ptr(a), _ptr(b), _ptr(third), int(assign)))

    def real_import(self, grid):
        if grid.dtype != torch.float64 or grid.numel() != self.nx_local*self.gridsize**2:
            raise TypeError('expected a float64 device grid of shape (nx_local, G, G)')
        check(self.lib.pm_real_import(self._h, _ptr(grid)))

    def fourier_resize_into(self, other):
        """Copy this context's Fourier slab into `other`'s (another grid size): modes |k| < min(G, G')/2."""
        check(self.lib.pm_fourier_resize(self._h, other._h))

    def fourier_copy_modes_into(self, other, deconv_order=0, shift=None, scale=1.0, src_saved=False, dst_saved=False,
                                accumulate=False):
        """copy_modes onto a context of another grid size (see pm_fourier_copy_modes in include/pmgrav.h)"""
        check(self.lib.pm_fourier_copy_modes(self._h, other._h, int(deconv_order), vec3(shift), float(scale), int(src_saved),
                                             int(dst_saved), int(accumulate)))

    # -- whole kick -------------------------------------------------------------------
    def kick_long(self, pos, mom, params, sum_mom2=None):
        check(self.lib.pm_kick_long(self._h, _particles(pos), _particles(mom), pos.shape[0], ctypes.byref(params), _ptr(sum_mom2)))

    def kick_drift(self, pos, mom, params, dt_over_mass, sum_mom2=None):
        """kick_long + Component.drift of the same particles in one call (drift fused into the gather/kick kernel)."""
        check(self.lib.pm_kick_drift(self._h, _particles(pos), _particles(mom), pos.shape[0], ctypes.byref(params),
                                     float(dt_over_mass), _ptr(sum_mom2)))

    def kick_long_host(self, pos_np, mom_np, params, dt_over_mass=0.0, want_sum=False):
        """Host (numpy, C-contiguous float64 (N,3)) buffers in and out; returns Σmom² or None."""
        for a in (pos_np, mom_np):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 2 and a.shape[1] == 3):
                raise TypeError('host particle arrays must be C-contiguous float64 (N, 3)')
        s = ctypes.c_double(0.0)
        check(self.lib.pm_kick_long_host(self._h, pos_np.ctypes.data_as(ctypes.c_void_p), mom_np.ctypes.data_as(ctypes.c_void_p),
                                         pos_np.shape[0], ctypes.byref(params), float(dt_over_mass),
                                         ctypes.byref(s) if want_sum else None))
        return s.value if want_sum else None

    # -- taps -------------------------------------------------------------------------
    def get_grid(self, which=PM_TAP_REAL):
        n = self.lib.pm_tap_size(self._h, int(which))
        out = np.empty(n, dtype=np.float64)
        check(self.lib.pm_get_grid(self._h, int(which), out.ctypes.data_as(ctypes.c_void_p)))
        G = self.gridsize
        if which == PM_TAP_FOURIER:
            return out.view(np.complex128).reshape(G, self.nj_local, G//2 + 1)
        return out.reshape(self.nx_local, G, G)

    def set_grid(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == (self.nx_local, self.gridsize, self.gridsize)
        check(self.lib.pm_set_grid(self._h, arr.ctypes.data_as(ctypes.c_void_p)))


def make_kick_params(*, mass, boxsize, gridsize, order, G_Newton, dt_rho_over_dt1, dt_kick,
                     diff_order=2, deconvolve=True, interlace=False, r_scale=0.0):
    """Scalars of one long-range kick, formed exactly as the reference does on the host:
    contribution  mesh.py:1550-1573 with fft_factor = G⁻³ (mesh.py:582)
    deconv_order  interactions.py:2069-2080 (up + down promoted to the global slab)
    prefactor     interactions.py:2105  ℝ[-boxsize**2*G_Newton/π]
    gauss         interactions.py:2112  (2π/boxsize·scale)²
    kick_factor   interactions.py:2386  mass·(−ᔑdt[ᔑdt_key])
    """
    G = int(gridsize)
    contribution = dt_rho_over_dt1
    contribution *= mass
    contribution_factor = float(G)**(-3)*(G/boxsize)**3
    contribution *= contribution_factor
    p = KickParams()
    p.order = int(order)
    p.diff_order = int(diff_order)
    p.deconv_order = int(order)*(2 if deconvolve else 0)
    p.interlace = int(bool(interlace))
    p.contribution = contribution
    p.prefactor = -boxsize**2*G_Newton/np.pi
    p.gauss = (2*np.pi/boxsize*r_scale)**2 if r_scale else 0.0
    p.kick_factor = mass*(-dt_kick)
    return p

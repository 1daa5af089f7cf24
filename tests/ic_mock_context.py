"""TEST INFRASTRUCTURE ONLY — numpy models of the libpmgrav entry points (include/pmgrav.h) that the host mirror
sequences, one rank, fp64, so that the CPU suite can run the *host code* of concept_b200 — parameters, orchestration,
factors and signs handed to the kernels, time loop, rung scheduling — against the reference's golden vectors
without a GPU.  The kernels themselves are checked by the `-m gpu` tests.

  MockContext          the initial-condition entry points (pm_ic_*, pm_fourier_operate, pm_fft_*, pm_kspace_potential,
                       pm_real_export/import, pm_fourier_resize, pm_lpt_accumulate) as plain numpy
  HostKernelContext    the same, with the element code of csrc/pm_ic_ops.cuh / pm_copy_ops.cuh itself, compiled for the
                       CPU (tests/ic_host_harness.cu), on buffers laid out as on the device
  MeshMockContext      + the mesh operators particle_mesh / powerspec sequence (deposit, deconvolution, gather, …),
                       built from oracle/pm_oracle.py, and pm_fourier_copy_modes (device code on the CPU)
  PMKickMockContext    + pm_kick_long, pm_drift, pm_sum_mom2: what main.timeloop needs
  ShortRangeFakeLib    the P³M entry points shortrange.py calls with raw pointers

Slab layout as on the device with one rank: complex [i][j][kk] (PM_TAP_FOURIER), real [i][j][k].
"""
import numpy as np
import torch


class MockContext:
    torch_device = torch.device('cpu')
    device = 0

    def __init__(self, gridsize, boxsize):
        self.gridsize, self.boxsize = int(gridsize), float(boxsize)
        G = self.gridsize
        self.nx_local = self.nj_local = G
        self.x_start = self.j_start = 0
        self.fourier = None
        self.real = None
        self.saved = None
        i = np.fft.fftfreq(G, 1/G).astype(np.int64)
        self.k = (i[:, None, None], i[None, :, None], np.arange(G//2 + 1, dtype=np.int64)[None, None, :])
        nyq = G//2
        self.live = np.broadcast_to((np.abs(self.k[0]) < nyq) & (np.abs(self.k[1]) < nyq) & (self.k[2] < nyq),
                                    (G, G, nyq + 1)).copy()
        self.live_nonzero = self.live.copy()
        self.live_nonzero[0, 0, 0] = False
        self.k2 = self.k[0]**2 + self.k[1]**2 + self.k[2]**2

    # -- pm_ic_lattice
    def ic_lattice(self, pos, mom, ids, shift, index_bgn, id_bgn):
        G = self.gridsize
        shift = shift or (0, 0, 0)
        cell = self.boxsize/G
        ax = [((0.5 + shift[d]) + np.arange(G))*cell for d in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing='ij')
        n = G**3
        pos[index_bgn:index_bgn + n] = torch.from_numpy(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1))
        mom[index_bgn:index_bgn + n] = 0
        if ids is not None:
            ids[index_bgn:index_bgn + n] = id_bgn + torch.arange(n)
        return n

    # -- pm_ic_potential
    def ic_potential(self, noise, amplitudes, k2_max, shift=None, lap_factor=1.0):
        G = self.gridsize
        noise = noise.numpy().view(np.complex128).reshape(G, G, G//2 + 1)
        amplitudes = amplitudes.numpy()
        v = noise.copy()
        if shift is not None and tuple(shift) != (0, 0, 0):
            θ = sum(self.k[d]*(-2*np.pi/G*(-shift[d])) for d in range(3))
            v = v*(np.cos(θ) + 1j*np.sin(θ))
        kf = 2*np.pi/self.boxsize
        k2 = np.where(self.live_nonzero, self.k2, 1)
        v = amplitudes[np.minimum(k2, k2_max)]*v
        if lap_factor != 0:
            v = v*((-lap_factor/kf**2)/k2)
        self.fourier = np.where(self.live_nonzero, v, 0)
        self.real = None

    def ic_nongaussianity(self, f_nl):
        assert self.real is not None
        self.real = self.real + f_nl*self.real**2

    def slab_save(self):
        self.saved = self.fourier.copy()

    # -- pm_fourier_operate (deconv_order 0, no shift)
    def fourier_operate(self, deconv_order=0, shift=None, scale=1.0, diff_dim=-1, from_saved=False):
        assert deconv_order == 0 and shift is None
        src = self.saved if from_saved else self.fourier
        assert src is not None
        v = src*scale
        if diff_dim >= 0:
            v = v*1j*(2*np.pi/self.boxsize)*self.k[diff_dim]
        self.fourier = np.where(self.live, v, 0)
        self.real = None

    # -- pm_kspace_potential (deconv_order 0, gauss 0)
    def kspace_potential(self, prefactor, deconv_order, gauss=0.0, scale=1.0):
        assert deconv_order == 0 and gauss == 0 and self.fourier is not None
        k2 = np.where(self.live_nonzero, self.k2, 1)
        self.fourier = np.where(self.live_nonzero, self.fourier*(scale*prefactor/k2), 0)

    def fft_backward(self):
        G = self.gridsize
        assert self.fourier is not None
        self.real = np.fft.irfftn(self.fourier, s=(G, G, G), axes=(0, 1, 2))*float(G)**3
        self.fourier = None

    def fft_forward(self):
        assert self.real is not None
        self.fourier = np.fft.rfftn(self.real, axes=(0, 1, 2))
        self.real = None

    # -- pm_ic_displace
    def ic_displace(self, pos, mom, index_bgn, dim, pos_factor=1.0, mom_factor=0.0):
        assert self.real is not None
        ψ = torch.from_numpy(self.real.ravel().copy())
        n = ψ.numel()
        if pos is not None:
            pos[index_bgn:index_bgn + n, dim] += pos_factor*ψ
        if mom is not None:
            mom[index_bgn:index_bgn + n, dim] += mom_factor*ψ

    def ic_wrap(self, pos, n):
        pos[:n] = torch.from_numpy(np.mod(pos[:n].numpy(), self.boxsize))

    def real_export(self, out=None):
        if out is None:
            return torch.from_numpy(self.real.copy())
        out.copy_(torch.from_numpy(self.real))
        return out

    def ic_2lpt_source(self, d00, d11, d22, d01, d12, d02):
        a, b, c, e, f, h = (t.numpy() for t in (d00, d11, d22, d01, d12, d02))
        self.real = -(a*b) - b*c - c*a + e*e + f*f + h*h
        self.fourier = None

    def lpt_accumulate(self, acc, factor, a, b, third=None, assign=False):
        v = a*b if third is None else (a*b)*third
        if assign:
            acc.copy_(factor*v)
        else:
            acc += factor*v

    def real_import(self, grid):
        self.real, self.fourier = grid.numpy().copy(), None

    def fourier_resize_into(self, other):
        Gs, Gd = self.gridsize, other.gridsize
        n = min(Gs, Gd)//2
        out = np.zeros((Gd, Gd, Gd//2 + 1), dtype=complex)
        idx = np.arange(-n + 1, n)
        out[np.ix_(idx % Gd, idx % Gd, np.arange(n))] = self.fourier[np.ix_(idx % Gs, idx % Gs, np.arange(n))]
        other.fourier = out
        other.real = None


class HostKernelContext(MockContext):
    """MockContext with the operations of csrc/pm_ic.cu executed by the *device code itself*, compiled for the CPU
    (tests/ic_host_harness.cu over csrc/pm_ic_ops.cuh), on buffers laid out as on the device: one allocation of
    G·G·(G+2) doubles that holds the padded real grid [i][j][G+2] or, in place, the Fourier slab [i][j][G/2+1].
    The FFTs and the pre-existing k-space kernels (pm_fourier_operate, pm_kspace_potential; GPU-tested elsewhere)
    stay numpy."""
    lib = None

    def __init__(self, gridsize, boxsize):
        super().__init__(gridsize, boxsize)
        G = self.gridsize
        self.buf = np.zeros(G*G*(G + 2))

    # views of the in-place buffer
    def _real_view(self):
        G = self.gridsize
        return self.buf.reshape(G, G, G + 2)

    def _fourier_view(self):
        G = self.gridsize
        return self.buf.view(np.complex128).reshape(G, G, G//2 + 1)

    def _sync_from_buf(self, fourier):
        if fourier:
            self.fourier, self.real = self._fourier_view().copy(), None
        else:
            self.real, self.fourier = self._real_view()[:, :, :self.gridsize].copy(), None

    def _sync_to_buf(self):
        if self.fourier is not None:
            self._fourier_view()[...] = self.fourier
        else:
            self.buf[:] = np.nan          # padding is garbage on the device too
            self._real_view()[:, :, :self.gridsize] = self.real

    @staticmethod
    def _p(a):
        import ctypes
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def ic_lattice(self, pos, mom, ids, shift, index_bgn, id_bgn):
        import ctypes
        G = self.gridsize
        shift = shift or (0, 0, 0)
        n = G**3
        host = pos.numpy()
        assert host.flags.c_contiguous
        self.lib.h_lattice(self._p(host), G, G, ctypes.c_double(0 + 0.5 + shift[0]), ctypes.c_double(0 + 0.5 + shift[1]),
                           ctypes.c_double(0 + 0.5 + shift[2]), ctypes.c_double(self.boxsize/G), ctypes.c_int64(index_bgn))
        mom[index_bgn:index_bgn + n] = 0
        ids[index_bgn:index_bgn + n] = id_bgn + torch.arange(n)
        return n

    def ic_potential(self, noise, amplitudes, k2_max, shift=None, lap_factor=1.0):
        import ctypes
        G = self.gridsize
        kf = 2*np.pi/self.boxsize
        th = np.array([-2*np.pi/G*(-(shift[d] if shift is not None else 0.0)) for d in range(3)])
        rotate = int(shift is not None and any(s != 0 for s in shift))
        self.buf[:] = np.nan
        self.lib.h_potential(self._p(noise.numpy()), self._p(self.buf), G, G, 0, self._p(amplitudes.numpy()), int(k2_max),
                             self._p(th), rotate, ctypes.c_double(-lap_factor/(kf*kf)))
        self._sync_from_buf(fourier=True)

    def ic_displace(self, pos, mom, index_bgn, dim, pos_factor=1.0, mom_factor=0.0):
        import ctypes
        G = self.gridsize
        self._sync_to_buf()
        self.lib.h_displace(self._p(None if pos is None else pos.numpy()), self._p(None if mom is None else mom.numpy()),
                            self._p(self.buf), G, G + 2, G, ctypes.c_int64(index_bgn), int(dim), ctypes.c_double(pos_factor),
                            ctypes.c_double(mom_factor))

    def ic_nongaussianity(self, f_nl):
        import ctypes
        G = self.gridsize
        self._sync_to_buf()
        self.lib.h_nongaussianity(self._p(self.buf), G, G + 2, G, ctypes.c_double(f_nl))
        self._sync_from_buf(fourier=False)

    def ic_wrap(self, pos, n):
        import ctypes
        self.lib.h_wrap(self._p(pos.numpy()), ctypes.c_int64(3*n), ctypes.c_double(self.boxsize))

    def ic_2lpt_source(self, d00, d11, d22, d01, d12, d02):
        G = self.gridsize
        self.buf[:] = np.nan
        self.lib.h_source(self._p(self.buf), *(self._p(t.numpy()) for t in (d00, d11, d22, d01, d12, d02)), G, G + 2, G)
        self._sync_from_buf(fourier=False)

    def lpt_accumulate(self, acc, factor, a, b, third=None, assign=False):
        import ctypes
        self.lib.h_lpt_accumulate(self._p(acc.numpy()), ctypes.c_int64(acc.numel()), ctypes.c_double(factor), self._p(a.numpy()),
                                  self._p(b.numpy()), self._p(None if third is None else third.numpy()), int(assign))

    def real_import(self, grid):
        G = self.gridsize
        self.buf[:] = np.nan
        self.lib.h_import(self._p(self.buf), self._p(grid.numpy()), G, G + 2, G)
        self._sync_from_buf(fourier=False)

    def real_export(self, out=None):
        G = self.gridsize
        self._sync_to_buf()
        if out is None:
            out = torch.empty((G, G, G), dtype=torch.float64)
        self.lib.h_export(self._p(self.buf), self._p(out.numpy()), G, G + 2, G)
        return out

    def fourier_resize_into(self, other):
        self._sync_to_buf()
        other.buf[:] = np.nan
        self.lib.h_resize(self._p(self.buf), self._p(other.buf), self.gridsize, other.gridsize)
        other._sync_from_buf(fourier=True)


class MeshMockContext(HostKernelContext):
    """HostKernelContext plus a numpy model (built from oracle/pm_oracle.py) of the mesh entry points that
    concept_b200.interactions.particle_mesh sequences — pm_grid_zero, pm_deposit, pm_fourier_operate with deconvolution
    and lattice phase, pm_kspace_potential, pm_slab_accumulate, pm_gather, pm_gather_kick — and pm_fourier_copy_modes
    executed by the device code compiled for the CPU.  Used to check the host orchestration of component-specific
    grid sizes against the reference's golden vectors without a GPU."""

    def __init__(self, gridsize, boxsize):
        super().__init__(gridsize, boxsize)
        G = self.gridsize
        k = np.where(np.arange(G) >= G//2, np.arange(G) - G, np.arange(G))
        self.tab_x = k*(np.pi/G) + 2.220446049250313e-16
        self.tab_sin = np.sin(self.tab_x)

    def grid_zero(self):
        G = self.gridsize
        self.real, self.fourier = np.zeros((G, G, G)), None

    def deposit(self, pos, order, contribution, shift=None):
        from oracle import pm_oracle as O
        assert self.real is not None
        self.real = self.real + O.deposit(pos.numpy(), self.boxsize, self.gridsize, order, contribution, shift or (0.0, 0.0, 0.0))

    def halo_add(self):
        pass

    def halo_fill(self):
        pass

    def slab_accumulate(self):
        self.saved = self.saved + self.fourier

    def slab_restore(self):
        self.fourier, self.real = self.saved.copy(), None

    def fourier_operate(self, deconv_order=0, shift=None, scale=1.0, diff_dim=-1, from_saved=False):
        from oracle import pm_oracle as O
        G = self.gridsize
        src = self.saved if from_saved else self.fourier
        assert src is not None
        v = src*(O.deconv_factor(G, deconv_order)*scale)
        if shift is not None and tuple(shift) != (0, 0, 0):
            v = v*O.interlace_phase(G, shift)
        if diff_dim >= 0:
            v = v*1j*(2*np.pi/self.boxsize)*self.k[diff_dim]
        self.fourier, self.real = np.where(self.live, v, 0), None

    def kspace_potential(self, prefactor, deconv_order, gauss=0.0, scale=1.0):
        from oracle import pm_oracle as O
        G = self.gridsize
        assert self.fourier is not None
        k2 = np.where(self.live_nonzero, self.k2, 1).astype(np.float64)
        factor = O.deconv_factor(G, deconv_order)*scale*(prefactor/k2)
        if gauss:
            factor = factor*np.exp(k2*(-gauss))
        self.fourier = np.where(self.live_nonzero, self.fourier*factor, 0)

    def power_k2(self, k2_max, power, count=None):
        from oracle import pm_oracle as O
        assert self.fourier is not None
        p2, c2 = O.power_by_k2(self.fourier, k2_max)
        power += torch.from_numpy(p2)
        if count is not None:
            count += torch.from_numpy(c2.astype(np.int64))

    def gather(self, which, pos, mom, order, dim, factor, shift=None):
        from oracle import pm_oracle as O
        assert which == 0 and self.real is not None
        mom[:, dim] += torch.from_numpy(O.gather(self.real, pos.numpy(), self.boxsize, order, shift or (0.0, 0.0, 0.0))*factor)

    def gather_kick(self, pos, mom, order, diff_order, factor, shift=None, sum_mom2=None):
        from oracle import pm_oracle as O
        assert self.real is not None
        for dim in range(3):
            force = O.diff_grid(self.real, dim, diff_order, self.boxsize/self.gridsize)
            mom[:, dim] += torch.from_numpy(O.gather(force, pos.numpy(), self.boxsize, order, shift or (0.0, 0.0, 0.0))*factor)

    def fourier_copy_modes_into(self, other, deconv_order=0, shift=None, scale=1.0, src_saved=False, dst_saved=False,
                                accumulate=False):
        import ctypes
        Gs, Gd = self.gridsize, other.gridsize
        assert Gs != Gd
        src = np.ascontiguousarray(self.saved if src_saved else self.fourier)
        if accumulate:
            dst = np.ascontiguousarray(other.saved if dst_saved else other.fourier).copy()
        else:
            dst = np.full((Gd, Gd, Gd//2 + 1), np.nan + 0j)
        th = np.array([-2*np.pi/Gs*(shift[d] if shift is not None else 0.0) for d in range(3)])
        rotate = int(shift is not None and any(s != 0 for s in shift))
        self.lib.h_copy_modes(self._p(src), self._p(dst), Gs, Gd, int(deconv_order), self._p(th), rotate,
                              ctypes.c_double(np.pi/Gd - np.pi/Gs), ctypes.c_double(scale), self._p(self.tab_x),
                              self._p(self.tab_sin), int(accumulate))
        if dst_saved:
            other.saved = dst
        else:
            other.fourier, other.real = dst, None


class PMKickMockContext(MeshMockContext):
    """MeshMockContext plus a numpy model of the whole-kick and particle entry points the PM time loop uses
    (pm_kick_long with the scalars of struct pm_kick_params, pm_drift, pm_sum_mom2), so that concept_b200.main.timeloop —
    the time-step controller, the kick/drift sequencing, dumps, static time-stepping — runs on the CPU."""

    def kick_long(self, pos, mom, params, sum_mom2=None):
        from oracle import pm_oracle as O
        G, L = self.gridsize, self.boxsize
        x = pos.numpy()
        shifts = [(0.0, 0.0, 0.0)] + ([(O.BCC_SHIFT,)*3] if params.interlace else [])
        nl = len(shifts)
        slab = 0
        for s in shifts:
            f = O.forward_fft(O.deposit(x, L, G, params.order, params.contribution, s))
            f[~O.mode_mask(G)] = 0
            slab = slab + (f*(1/nl)*O.interlace_phase(G, s) if nl > 1 else f)
        k2 = np.where(self.live_nonzero, self.k2, 1).astype(np.float64)
        factor = O.deconv_factor(G, params.deconv_order)*(params.prefactor/k2)
        if params.gauss:
            factor = factor*np.exp(k2*(-params.gauss))
        slab = np.where(self.live_nonzero, slab*factor, 0)
        kick = np.zeros_like(x)
        for s in shifts:
            f = slab*(1/nl)*O.interlace_phase(G, s) if nl > 1 else slab
            if params.diff_order == 0:
                for dim in range(3):
                    force = O.backward_fft(1j*(2*np.pi/L)*self.k[dim]*f, G)
                    kick[:, dim] += O.gather(force, x, L, params.order, s)*params.kick_factor
            else:
                phi = O.backward_fft(f, G)
                for dim in range(3):
                    kick[:, dim] += O.gather(O.diff_grid(phi, dim, params.diff_order, L/G), x, L, params.order, s)*params.kick_factor
        mom += torch.from_numpy(kick)

    def drift(self, pos, mom, dt_over_mass):
        from oracle import pm_oracle as O
        pos.copy_(torch.from_numpy(O.drift(pos.numpy(), mom.numpy(), dt_over_mass, self.boxsize)))

    def sort_particles(self, pos, mom, ids=None, n=None):
        """pm_sort_particles: stable re-ordering by grid cell (x plane, y row, z)"""
        import torch
        n = pos.shape[0] if n is None else int(n)
        G = self.gridsize
        cell = np.clip((pos[:n].numpy()*(G/self.boxsize)).astype(np.int64), 0, G - 1)
        order = torch.as_tensor(np.argsort((cell[:, 0]*G + cell[:, 1])*G + cell[:, 2], kind='stable'))
        pos[:n] = pos[:n][order]
        mom[:n] = mom[:n][order]
        if ids is not None:
            ids[:n] = ids[:n][order]

    def sum_mom2(self, mom, out=None):
        from oracle import pm_oracle as O
        return O.sum_mom2(mom.numpy())


class ShortRangeFakeLib:
    """TEST INFRASTRUCTURE — a numpy model of the P³M entry points that concept_b200.shortrange calls straight on
    `ctx.lib` with raw pointers (pm_shortrange, pm_apply_dmom, pm_assign_rungs, pm_flag_rung_jumps,
    pm_apply_rung_jumps; csrc/pm_shortrange.cu).  The tensors live in host memory here, so the pointers are mapped
    back to numpy arrays.  Everything else is forwarded to the CPU build of the device code (`harness`)."""

    def __init__(self, harness, boxsize):
        self._harness, self._boxsize = harness, boxsize

    def __getattr__(self, name):
        return getattr(self._harness, name)

    @staticmethod
    def _arr(ptr, n, ctype):
        import ctypes
        if isinstance(ptr, int):
            return np.ctypeslib.as_array((ctype*n).from_address(ptr))
        return np.ctypeslib.as_array(ptr, shape=(n,)) if not isinstance(ptr, ctypes.Array) else np.ctypeslib.as_array(ptr)

    @staticmethod
    def _get_rung(acc, current, rung_factor, n_rungs):
        acc2 = (acc**2).sum(axis=1)
        with np.errstate(divide='ignore'):
            rf = rung_factor + 0.25*np.log2(np.where(acc2 == 0, 1.0, acc2))
        rung = np.where(rf < 0, 0, np.where(rf > n_rungs - 1, n_rungs - 1, 1 + np.trunc(np.clip(rf, 0, n_rungs)).astype(np.int64)))
        return np.where(acc2 == 0, current, rung).astype(np.int8)

    def pm_shortrange(self, h, pos, n, rung, rung_jumped, lowest_active_rung, factors, nfactors, rng, table, tablesize, maxr2, dmom):
        import ctypes
        from oracle import pm_oracle as O
        x = self._arr(pos, 3*n, ctypes.c_double).reshape(n, 3)
        r = self._arr(rung, n, ctypes.c_int8)
        rj = self._arr(rung_jumped, n, ctypes.c_int8)
        f = self._arr(factors, nfactors, ctypes.c_double)
        tab = self._arr(table, tablesize, ctypes.c_double)
        out = self._arr(dmom, 3*n, ctypes.c_double).reshape(n, 3)
        active = r >= lowest_active_rung
        S = O.shortrange_sums(x, self._boxsize, rng, tab, maxr2, active)
        out[active] = S[active]*f[rj[active]][:, None]
        return 0

    def pm_apply_dmom(self, h, mom, dmom, n, rung, rung_jumped, lowest_active_rung, conv, nconv, apply):
        import ctypes
        m = self._arr(mom, 3*n, ctypes.c_double).reshape(n, 3)
        dm = self._arr(dmom, 3*n, ctypes.c_double).reshape(n, 3)
        r = self._arr(rung, n, ctypes.c_int8)
        rj = self._arr(rung_jumped, n, ctypes.c_int8)
        cv = self._arr(conv, nconv, ctypes.c_double)
        active = r >= lowest_active_rung
        if apply:
            m[active] += dm[active]
        dm[active] = dm[active]*cv[rj[active]][:, None]
        return 0

    def pm_assign_rungs(self, h, acc, n, rung_factor, n_rungs, rung, rung_jumped, counts):
        import ctypes
        a = self._arr(acc, 3*n, ctypes.c_double).reshape(n, 3)
        r = self._arr(rung, n, ctypes.c_int8)
        rj = self._arr(rung_jumped, n, ctypes.c_int8)
        new = self._get_rung(a, r, rung_factor, n_rungs)
        r[:] = new
        rj[:] = new
        for q in range(n_rungs):
            counts[q] = int((new == q).sum())
        return 0

    def pm_flag_rung_jumps(self, h, acc, n, rung, rung_jumped, lowest_active_rung, rf_up, rf_down, dt1, n_rungs, any_ref):
        import ctypes
        a = self._arr(acc, 3*n, ctypes.c_double).reshape(n, 3)
        r = self._arr(rung, n, ctypes.c_int8)
        rj = self._arr(rung_jumped, n, ctypes.c_int8)
        dt = self._arr(dt1, 3*n_rungs - 1, ctypes.c_double)
        considered = (r >= lowest_active_rung) & (dt[r] != 0)
        up = considered & (self._get_rung(a, r, rf_up, n_rungs) > r)
        down_index = np.minimum(r.astype(np.int64) + n_rungs, 3*n_rungs - 2)
        down = considered & ~up & (dt[down_index] != -1) & (self._get_rung(a, r, rf_down, n_rungs) < r)
        rj[up] = r[up] + 2*n_rungs
        rj[down] = r[down] + n_rungs
        any_ref._obj.value = int(up.any() or down.any())
        return 0

    def pm_apply_rung_jumps(self, h, n, rung, rung_jumped, n_rungs, counts):
        import ctypes
        r = self._arr(rung, n, ctypes.c_int8)
        rj = self._arr(rung_jumped, n, ctypes.c_int8)
        up, down = rj >= 2*n_rungs, (rj >= n_rungs) & (rj < 2*n_rungs)
        r[up] += 1
        r[down] -= 1
        rj[up | down] = r[up | down]
        for q in range(n_rungs):
            counts[q] = int((r == q).sum())
        return 0

"""Initial-condition generator of the host mirror: particle realisations from
primordial noise with 1LPT / 2LPT / 3LPT (optionally back-scaled, dealiased, with local non-Gaussianity).

Reference (ic.py): PseudoRandomNumberGenerator :67-232, get_amplitudes :542-627, realize_grid :670-782,
generate_primordial_noise :928-1163, realize_particles :1199-1399, carryout_1lpt :1447-1509,
carryout_2lpt :1539-1589, carryout_3lpt_a/b/c :1619-1893, handle_lpt_term :1895-2057, diff_ifft :2093-2108, preinitialize_particles
:2138-2247, displace_particles :2249-2283; mesh.py: fourier_curve_loop / fourier_curve_slice_loop
:2909-3044, get_fourier_curve_coords :3109-3160, laplacian_inverse :3422-3437, fourier_diff :3470-3510.

Division of labour: the primordial noise is drawn on the host — it is defined by sequential NumPy bit
streams (one per kj slice), so it cannot be generated out of order — but vectorised: one draw per
stream and slice instead of one Python call per mode.  Everything downstream (amplitudes × noise,
lattice phase, inverse Laplacian, Fourier differentiation, inverse FFT, displacement of the lattice
particles, 2LPT source, dealiasing resize, periodic wrap) runs in libpmgrav.so (csrc/pm_ic.cu); torch is
used for device memory only.

On several GPUs the realisation is replicated (every rank realises the deterministic particle set on a
private one-rank context and keeps its x-slab), see _get_context.

Not built: fluid realisations and the non-linear ("structure": "non-linear") realisations.
"""
import math
import os

import numpy as np
import torch

from . import commons, communication, linear, mesh
from .commons import abort, masterprint, masterwarn, π
from .integration import hubble


# ------------------------------------------------------------------------------------------ random numbers
class PseudoRandomNumberGenerator:
    """ic.py:67-232.  Same seeding (salt for integer seeds), same child streams (spawn) and the same
    sequences; draws are served from cached batches like the reference's, or as whole arrays (`*_array`)."""
    streams = {name: attr for name, attr in vars(np.random).items()
               if isinstance(attr, type) and issubclass(attr, np.random.BitGenerator) and attr is not np.random.BitGenerator}

    def __init__(self, seed=None, stream=None, cache_size=2**12, salt=True):
        stream = stream or commons.params.random_generator
        if salt and isinstance(seed, (int, np.integer)) and not isinstance(seed, bool):
            seed = int(seed) + int(π*1e+8) + 137           # magic number + lucky_seed_offset, ic.py:108-116
        if not isinstance(seed, np.random.SeedSequence):
            seed = np.random.SeedSequence(seed)
        self.seed, self.stream, self.cache_size = seed, stream, int(cache_size)
        bit_generator = self.streams.get(stream)
        if bit_generator is None and stream == 'PCG64DXSM':
            masterwarn(f'Pseudo-random bit generator "{stream}" not available in NumPy. Falling back to "PCG64".')
            stream = 'PCG64'
            bit_generator = self.streams.get(stream)
        if bit_generator is None:
            abort(f'Pseudo-random bit generator "{stream}" not available in NumPy. '
                  f'The available ones are {", ".join(self.streams)}.')
        self.bit_generator = bit_generator(self.seed)
        self.generator = np.random.Generator(self.bit_generator)
        self._cache = {'uniform': None, 'gaussian': None, 'rayleigh': None}
        self._index = dict.fromkeys(self._cache, self.cache_size - 1)

    def spawn(self, spawn_key=0):
        spawn_key = tuple(int(k) for k in (spawn_key if isinstance(spawn_key, (tuple, list)) else (spawn_key, )))
        seed = np.random.SeedSequence(self.seed.entropy, spawn_key=(self.seed.spawn_key + spawn_key))
        return type(self)(seed, self.stream, self.cache_size)

    def _draw(self, kind, size):
        if kind == 'uniform':
            return self.generator.uniform(0, 1, size=size)
        if kind == 'gaussian':
            return self.generator.normal(0, 1, size=size)
        return self.generator.rayleigh(1, size=size)

    def _next(self, kind):
        self._index[kind] += 1
        if self._index[kind] == self.cache_size:
            self._index[kind] = 0
            self._cache[kind] = self._draw(kind, self.cache_size)
        return self._cache[kind][self._index[kind]]

    def uniform(self, low=0, high=1):
        return low + self._next('uniform')*(high - low)

    def gaussian(self, scale=1):
        return self._next('gaussian')*scale

    def rayleigh(self, scale=1):
        return self._next('rayleigh')*scale

    # whole-array forms: the next `size` numbers of the same sequences (fresh generators only)
    def uniform_array(self, size, low=0, high=1):
        return low + self._draw('uniform', size)*(high - low)

    def rayleigh_array(self, size, scale=1):
        return self._draw('rayleigh', size)*scale


# ---------------------------------------------------------------------------- Fourier space-filling curve
def _icbrt(x):
    r = np.rint(np.cbrt(x.astype(np.float64))).astype(np.int64)
    r -= r**3 > x
    r += (r + 1)**3 <= x
    return r


def get_fourier_curve_coords(key):
    """mesh.py:3109-3160 for an int64 array of keys → (ki, kj, kk)"""
    key = np.asarray(key, dtype=np.int64).copy()
    g = _icbrt(2*key)
    g += g & 1
    g += np.where(g**2*(g//2 + 1) <= key, 2, 0)
    s = g//2
    key -= (g - 2)**2*s
    s_safe = np.maximum(s, 1)
    f0 = 2*s**2 - s
    f1 = 2*s**2 - 2*s + f0
    f2 = f0 + f1
    f3 = 2*s**2 + f2
    ki, kj, kk = np.zeros_like(key), np.zeros_like(key), np.zeros_like(key)
    done = np.zeros(key.shape, dtype=bool)

    def face(mask, k, a, b, c):
        m = mask & ~done
        ki[m], kj[m], kk[m] = a(k, s, s_safe)[m], b(k, s, s_safe)[m], c(k, s, s_safe)[m]
        done[m] = True
    face(key < f0, key, lambda k, s, q: s - 1, lambda k, s, q: -s + 1 + k//q, lambda k, s, q: k % q)
    face(key < f1, key - f0, lambda k, s, q: -s + 1 + k//q, lambda k, s, q: s - 1, lambda k, s, q: k % q)
    face(key < f2, key - f1, lambda k, s, q: -s, lambda k, s, q: -s + 1 + k//q, lambda k, s, q: k % q)
    face(key < f3, key - f2, lambda k, s, q: -s + k//q, lambda k, s, q: -s, lambda k, s, q: k % q)
    face(np.ones(key.shape, dtype=bool), key - f3, lambda k, s, q: -s + k//(2*q), lambda k, s, q: -s + k % (2*q),
         lambda k, s, q: s)
    return ki, kj, kk


def fourier_curve_slice_order(gridsize):
    """The (ki, kk) visited by fourier_curve_slice_loop (mesh.py:2984-3044), in order; identical for every
    j slice.  All points except those on Nyquist planes, origin included."""
    nyquist = gridsize//2
    keys = []
    for s in range(nyquist):
        for f in range(1 + (2 if s < nyquist - 1 else 0)):
            key_bgn = f**2 + (f == 2) + (1 + 6*f - (f == 1))*s + (5 + 4*f - (f == 2))*s**2 + 4*s**3
            num = (s + 1)*(1 + (f == 2))
            step = 1 + ((num - 1) if f == 2 else 0)
            keys.append(key_bgn + step*np.arange(num, dtype=np.int64))
    ki, kj, kk = get_fourier_curve_coords(np.concatenate(keys))
    assert not kj.any()
    return ki, kk


def generate_primordial_noise(gridsize, fixed_amplitude=False, phase_shift=0):
    """ic.py:928-1163 on the host.  Returns the noise as a complex array in the reference's transposed
    Fourier layout [j][i][kk], kk = 0 … gridsize/2: origin nullified, Nyquist planes zero (the reference
    leaves them untouched and nullifies them in realize_grid)."""
    p = commons.params
    G, nyquist = int(gridsize), int(gridsize)//2
    slab = np.zeros((G, G, nyquist + 1), dtype=np.complex128)
    prng_amplitudes_common = PseudoRandomNumberGenerator(p.random_seeds['primordial amplitudes'])
    prng_phases_common = PseudoRandomNumberGenerator(p.random_seeds['primordial phases'])
    scale = 1/math.sqrt(2)

    def polar(r, θ):
        if phase_shift:
            θ = θ + phase_shift
        return r*np.cos(θ) + 1j*(r*np.sin(θ))

    if p.primordial_noise_imprinting == 'simple':
        n_total = G*G*(nyquist + 1)
        n_nyquist = G**2 + nyquist*(2*G - 1)
        ki, kj, kk = get_fourier_curve_coords(np.arange(n_total - n_nyquist, dtype=np.int64))
        n = len(ki)
        r = np.ones(n) if fixed_amplitude else prng_amplitudes_common.rayleigh_array(n, scale)
        θ = prng_phases_common.uniform_array(n, -π, π)
        value = polar(r, θ)
        lower = (ki < 0) | ((ki == 0) & (kj < 0))
        direct = (kk != 0) | lower
        slab[kj[direct] % G, ki[direct] % G, kk[direct]] = value[direct]
        conj = (kk == 0) & lower
        slab[(-kj[conj]) % G, (-ki[conj]) % G, 0] = np.conj(value[conj])
    elif p.primordial_noise_imprinting == 'distributed':
        spawn_key_offset = 2**32
        ki, kk = fourier_curve_slice_order(G)
        n = len(ki)
        i_direct, i_conj = ki % G, (-ki) % G

        def fill_slice(j):
            # one kj slice: its four child streams are its own, its plane of the slab too — slices are independent
            kj = j - (G if j >= nyquist else 0)
            if kj == -nyquist:
                return
            streams = [common.spawn(spawn_key_offset + sign*kj) for common in (prng_amplitudes_common, prng_phases_common)
                       for sign in (+1, -1)]
            if fixed_amplitude:
                r = r_conj = np.ones(n)
            else:
                r, r_conj = streams[0].rayleigh_array(n, scale), streams[1].rayleigh_array(n, scale)
            θ, θ_conj = streams[2].uniform_array(n, -π, π), streams[3].uniform_array(n, -π, π)
            direct = (kk != 0) | (ki < 0) | ((ki == 0) & (kj < 0))
            slab[j, i_direct[direct], kk[direct]] = polar(r[direct], θ[direct])
            conj = (kk == 0) & ((ki < 0) | ((ki == 0) & (kj >= 0)))      # a few modes of the kk = 0 plane only
            slab[j, i_conj[conj], 0] = np.conj(polar(r_conj[conj], θ_conj[conj]))
        # the draws and the cos/sin of a slice are numpy calls that release the GIL: slices run on a few host threads
        # (every rank realises the whole noise, so the threads are shared out between the ranks of a node)
        from . import communication
        workers = max(1, min(16, (os.cpu_count() or 1)//max(1, communication.nprocs)))
        if workers > 1 and G >= 64:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=workers) as pool:
                list(pool.map(fill_slice, range(G)))
        else:
            for j in range(G):
                fill_slice(j)
    else:
        abort(f'primordial_noise_imprinting = "{p.primordial_noise_imprinting}" not implemented')
    slab[0, 0, 0] = 0
    return slab


# ------------------------------------------------------------------------------------------------ amplitudes
def get_primordial_curvature_perturbation(k):
    """linear.py:3329-3341: ζ(k) = π·√(2A_s)·k^(−3/2)·(k/pivot)^((n_s−1)/2)·exp(α_s/4·ln(k/pivot)²)"""
    ps = commons.params.primordial_spectrum
    return (π*math.sqrt(2*ps['A_s'])/ps['pivot']**((ps['n_s'] - 1)/2))*k**(ps['n_s']/2 - 2)*np.exp(
        ps['α_s']/4*(np.log(k) - math.log(ps['pivot']))**2)


# The two look-ups the realisation makes into linear theory (CLASS in the reference, linear.py:2587, :2730).
# They are module attributes so that a caller (or a test) can install tabulated CLASS output instead of the
# analytic stand-ins of concept_b200.linear.
compute_transfer = linear.compute_transfer
compute_cosmo = linear.compute_cosmo


def get_amplitudes(gridsize, component, a, a_next=-1, variable=-1, multi_index=None, factor=1):
    """ic.py:542-627 for the primordial structure: table over integer k² of T(a, k)·ζ(k)·L^(−3/2)·factor"""
    if variable not in (0, 1):
        abort(f'get_amplitudes() called with variable = {variable}')
    options = component.realization_options
    transfer_spline, _ = compute_transfer(component, variable, gridsize, multi_index, a, a_next, options.get('gauge', 'nbody'),
                                          weight=None, backscale=options['backscale']*(variable == 0))
    nyquist = gridsize//2
    k2_max = 3*(nyquist - 1)**2
    amplitudes = np.zeros(k2_max + 1)
    normalization = commons.params.boxsize**(-1.5)*factor
    k_magnitude = (2*π/commons.params.boxsize)*np.sqrt(np.arange(1, k2_max + 1))
    transfer = np.array([transfer_spline.eval(k) for k in k_magnitude]) if not hasattr(transfer_spline, 'eval_array') \
        else transfer_spline.eval_array(k_magnitude)
    amplitudes[1:] = transfer*get_primordial_curvature_perturbation(k_magnitude)*normalization
    return amplitudes


# ----------------------------------------------------------------------------------------------- realisation
LATTICE_SHIFTS = {   # Lattice.shifts_all, mesh.py:85-100 (cell-centred grids: shift_amount = −½)
    'sc': [(0, 0, 0)],
    'bcc': [(0, 0, 0), (-.5, -.5, -.5)],
    'fcc': [(0, 0, 0), (0, -.5, -.5), (-.5, 0, -.5), (-.5, -.5, 0)],
}
n_particles_realized = {'components_tally': 0, 'components_total': 0, 'particles_tally': 0}


def _icbrt_int(n):
    r = round(n**(1/3))
    return r if r**3 == n else -1


def preic_lattice(N):
    """species.py:1106-1117"""
    if _icbrt_int(N) > 0:
        return 'sc'
    if N % 2 == 0 and _icbrt_int(N//2) > 0:
        return 'bcc'
    if N % 4 == 0 and _icbrt_int(N//4) > 0:
        return 'fcc'
    return ''


_private_contexts = {}


def _get_context(gridsize):
    """The fp64 context the realisation works on.  On one rank it is the cached context of mesh.get_context (shared
    with the PM solver when the grid sizes coincide).  On several ranks the realisation is *replicated*: the slab
    FFT of a multi-rank context would need the reference's distributed noise/lattice bookkeeping, while the
    realisation is a deterministic, one-off, G³-sized job — so every rank runs it on a private single-rank context
    and keeps its own slab of particles afterwards."""
    if communication.nprocs == 1:
        return mesh.get_context(gridsize, 'f64')
    ctx = _private_contexts.get(int(gridsize))
    if ctx is None:
        from .pmsolver import PMContext
        ctx = _private_contexts[int(gridsize)] = PMContext(gridsize, commons.params.boxsize, dtype='f64', rank=0, nranks=1,
                                                           device=communication.local_rank)
    return ctx


def _free_private_contexts():
    for ctx in _private_contexts.values():
        ctx.close()
    _private_contexts.clear()


def _device_noise(ctx, noise):
    """Reference layout [j][i][kk] → the context's slab layout [i][j_local][kk] as device doubles."""
    local = noise[ctx.j_start:ctx.j_start + ctx.nj_local].transpose(1, 0, 2)
    return torch.from_numpy(np.ascontiguousarray(local).view(np.float64)).to(ctx.torch_device)


def realize_particles(component, a, components_all=None):
    """ic.py:1199-1399"""
    p = commons.params
    options = dict(p.realization_options)
    options.update(getattr(component, 'realization_options', None) or {})
    component.realization_options = options
    if options['lpt'] not in {1, 2, 3}:
        abort(f'realize_particles() called with attempted {options["lpt"]}LPT')
    if component.representation != 'particles':
        abort(f'realize_particles() called with non-particle component {component.name}')
    kind = preic_lattice(component.N)
    if not kind:
        abort(f'Cannot initialize particle component {component.name} with N = {component.N} on a lattice, '
              f'as neither of {{N, N/2, N/4}} is a cubic number')
    gridsize = _icbrt_int(component.N//{'sc': 1, 'bcc': 2, 'fcc': 4}[kind])
    if gridsize % 2:
        abort(f'The particle lattice of {component.name} has an odd size {gridsize}; the FFT grids must be even')
    shifts = LATTICE_SHIFTS[kind]
    if n_particles_realized['components_tally'] == 0:
        others = [c for c in (components_all or [component]) if c.representation == 'particles' and c.mass == -1]
        n_particles_realized['components_total'] = max(len(others), 1)
    if kind == 'sc':
        total, tally = n_particles_realized['components_total'], n_particles_realized['components_tally']
        if total == 2:
            shifts = [LATTICE_SHIFTS['bcc'][tally % 2]]
        elif total == 4:
            shifts = [LATTICE_SHIFTS['fcc'][tally % 4]]
        elif total != 1:
            masterwarn(f'{total} ∉ {{1, 2, 4}} particle components are to be initialized on simple cubic lattices. '
                       f'This leads to anisotropies in the initial conditions.')
    if component.mass == -1:
        component.mass = component.ϱ_bar*p.boxsize**3/component.N
    masterprint(f'Realising {len(shifts)}×{gridsize}³ particles of {component.name} ...')
    growth_factors = dict.fromkeys(('D1', 'f1', 'D2', 'f2') + (('D3a', 'f3a', 'D3b', 'f3b', 'D3c', 'f3c') if options['lpt'] >= 3 else ()),
                                   float('nan'))
    if options['backscale'] or options['lpt'] > 1:
        cosmoresults = compute_cosmo(class_call_reason='in order to get growth factors')
        for key in growth_factors:
            growth_factors[key] = float(getattr(cosmoresults, f'growth_fac_{key}')(a))
    ctx = _get_context(gridsize)
    ctx_dealias = ctx
    if options['dealias'] and options['lpt'] > 1:
        gridsize_dealias = (gridsize*3)//2
        gridsize_dealias += gridsize_dealias & 1
        ctx_dealias = _get_context(gridsize_dealias)
    component.N_local = 0
    component.resize(component.N)
    component.N_local = component.N
    noise = _device_noise(ctx, generate_primordial_noise(gridsize, p.primordial_amplitude_fixed, p.primordial_phase_shift))
    n_particles = gridsize**3
    index_bgn = 0
    id_bgn = n_particles_realized['particles_tally']
    for shift in shifts:
        n_local = ctx.ic_lattice(component.pos, component.mom, component.ids, shift, index_bgn, id_bgn)
        carryout_1lpt(component, ctx, noise, shift, gridsize, options, a, growth_factors, index_bgn)
        if options['lpt'] >= 2:
            second1 = carryout_2lpt(component, ctx, ctx_dealias, a, growth_factors, index_bgn)
            if options['lpt'] >= 3:
                carryout_3lpt(component, ctx, ctx_dealias, a, growth_factors, index_bgn, second1)
            del second1
        id_bgn += n_particles
        index_bgn += n_local
    n_particles_realized['particles_tally'] = id_bgn
    n_particles_realized['components_tally'] += 1
    ctx.ic_wrap(component.pos, component.N_local)      # ic.py:1396-1398
    component._ids_set = True
    if communication.nprocs > 1:
        # exchange(component) (ic.py:1399): every rank has realised the whole (deterministic) particle set on its
        # own GPU; keep the particles of this rank's x-slab and release the rest
        pos, mom, ids = component.pos[:component.N], component.mom[:component.N], component.ids[:component.N]
        component.pos = component.mom = component.ids = None
        component.N_allocated = component.N_local = 0
        component.set_particles(pos, mom, ids, distribute=True)
        del pos, mom, ids
        _free_private_contexts()
    masterprint('done')


def _mom_factor(component, a):
    """displace_particles (ic.py:2265-2271): mom = a·m(a)·u"""
    return a*(a**(-3*component.w_eff(a=a))*component.mass)


def _displace_from_saved(component, ctx, index_bgn, pos_factor, mom_factor, dims=((0, 0, 1), (1, 1, 1), (2, 2, 1))):
    """Ψ = sign·ℱ⁻¹[i·k_along·Φ] (diff_ifft, ic.py:2093-2108) from the saved potential, added to component `target` of
    the lattice particles (displace_particles), for every (target, along, sign) of `dims` — the gradient by
    default.  A factor of None leaves that array untouched."""
    for target, along, sign in dims:
        ctx.fourier_operate(scale=sign, diff_dim=along, from_saved=True)
        ctx.fft_backward()
        ctx.ic_displace(None if pos_factor is None else component.pos, None if mom_factor is None else component.mom,
                        index_bgn, target, pos_factor or 0.0, mom_factor or 0.0)


def carryout_1lpt(component, ctx, noise, shift, gridsize, options, a, growth_factors, index_bgn):
    """ic.py:1447-1509: ∇²Φ⁽¹⁾ = −δ.  Leaves Φ⁽¹⁾ (from δ) in the context's saved Fourier slab."""
    velocity_factor = a*hubble(a)*growth_factors['f1']
    for variable in range(1 - options['backscale'], -1, -1):      # first θ, then δ
        amplitudes = get_amplitudes(gridsize, component, a, variable=variable)
        amplitudes_dev = torch.from_numpy(amplitudes).to(noise.device)
        nongaussianity = options.get('nongaussianity', 0)*(variable == 0)      # velocities are unaffected (ic.py:1488-1490)
        if nongaussianity:
            # realize_grid (ic.py:766-776): to real space, x += f·x², back with the forward normalisation G⁻³ — folded,
            # with laplacian_inverse(…, 2·variable − 1), into the prefactor of pm_kspace_potential
            ctx.ic_potential(noise, amplitudes_dev, len(amplitudes) - 1, shift, lap_factor=0.0)
            ctx.fft_backward()
            ctx.ic_nongaussianity(nongaussianity)
            ctx.fft_forward()
            ctx.kspace_potential(-(2*variable - 1)*float(gridsize)**(-3)*(commons.params.boxsize/(2*π))**2, 0)
        else:
            ctx.ic_potential(noise, amplitudes_dev, len(amplitudes) - 1, shift, lap_factor=2*variable - 1)
        ctx.slab_save()
        if variable == 1:
            _displace_from_saved(component, ctx, index_bgn, None, _mom_factor(component, a))
        elif options['backscale']:
            _displace_from_saved(component, ctx, index_bgn, 1.0, velocity_factor*_mom_factor(component, a))
        else:
            _displace_from_saved(component, ctx, index_bgn, 1.0, None)


def carryout_2lpt(component, ctx, ctx_dealias, a, growth_factors, index_bgn):
    """ic.py:1539-1589 with handle_lpt_term (:1895-2057) and diff_ifft (:2093-2108):
    ∇²Φ⁽²⁾ = D⁽²⁾/(D⁽¹⁾)²·(−Φ,₀₀Φ,₁₁ − Φ,₁₁Φ,₂₂ − Φ,₂₂Φ,₀₀ + Φ,₀₁² + Φ,₁₂² + Φ,₂₀²), the products formed in real
    space — on a grid enlarged by 3/2 (Orszag) when dealiasing."""
    dealias = ctx_dealias is not ctx
    fft_factor = float(ctx_dealias.gridsize)**(-3)
    potential_factor = fft_factor*growth_factors['D2']/growth_factors['D1']**2
    velocity_factor = a*hubble(a)*growth_factors['f2']
    second = _second_derivatives(ctx, ctx_dealias)
    ctx_dealias.ic_2lpt_source(*(second[ij] for ij in ((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2))))
    ctx_dealias.fft_forward()
    if dealias:
        ctx_dealias.fourier_resize_into(ctx)
    # laplacian_inverse(Φ2, potential_factor): ×(−potential_factor/k_f²)/k² (mesh.py:3422-3437)
    ctx.kspace_potential(-potential_factor*(commons.params.boxsize/(2*π))**2, 0)
    ctx.slab_save()
    _displace_from_saved(component, ctx, index_bgn, 1.0, velocity_factor*_mom_factor(component, a))
    return second


def _second_derivatives(ctx, ctx_dealias):
    """Φ,ᵢⱼ in real space for the potential in ctx's saved Fourier slab (fourier_diff twice + diff_ifft,
    mesh.py:3470-3510, ic.py:2093-2108), on the dealiasing grid if there is one; keyed by (i, j) both ways."""
    second = {}
    for i, j in ((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)):
        ctx.fourier_operate(diff_dim=i, from_saved=True)
        ctx.fourier_operate(diff_dim=j)
        if ctx_dealias is not ctx:
            ctx.fourier_resize_into(ctx_dealias)
        ctx_dealias.fft_backward()
        second[i, j] = second[j, i] = ctx_dealias.real_export()
    return second


def _lpt_potential(ctx, ctx_dealias, terms, potential_factor):
    """handle_lpt_term (ic.py:1895-2057) for a list of (factor, [grid, grid(, grid)]) terms, then the forward
    transform and laplacian_inverse(…, potential_factor): the potential ends up in ctx's saved Fourier slab.
    Products are formed pairwise in the order written; with dealiasing an intermediate product is cut back to
    the cube |k| < G/2 before the next factor (forward, shrink, enlarge, backward: the two unnormalised
    transforms cost Gd⁻³, folded into the term's factor)."""
    dealias = ctx_dealias is not ctx
    fft_factor = float(ctx_dealias.gridsize)**(-3)
    acc = torch.empty((ctx_dealias.nx_local, ctx_dealias.gridsize, ctx_dealias.gridsize), dtype=torch.float64,
                      device=ctx_dealias.torch_device)
    tmp = torch.empty_like(acc) if dealias else None
    for n, (factor, grids) in enumerate(terms):
        if len(grids) == 2:
            ctx_dealias.lpt_accumulate(acc, factor, grids[0], grids[1], None, assign=(n == 0))
        elif not dealias:
            ctx_dealias.lpt_accumulate(acc, factor, grids[0], grids[1], grids[2], assign=(n == 0))
        else:
            ctx_dealias.lpt_accumulate(tmp, 1.0, grids[0], grids[1], None, assign=True)
            ctx_dealias.real_import(tmp)
            ctx_dealias.fft_forward()
            ctx_dealias.fourier_resize_into(ctx)
            ctx.fourier_resize_into(ctx_dealias)
            ctx_dealias.fft_backward()
            ctx_dealias.real_export(tmp)
            ctx_dealias.lpt_accumulate(acc, factor*fft_factor, tmp, grids[2], None, assign=(n == 0))
    ctx_dealias.real_import(acc)
    ctx_dealias.fft_forward()
    if dealias:
        ctx_dealias.fourier_resize_into(ctx)
    ctx.kspace_potential(-potential_factor*(commons.params.boxsize/(2*π))**2, 0)
    ctx.slab_save()


def carryout_3lpt(component, ctx, ctx_dealias, a, growth_factors, index_bgn, P1):
    """carryout_3lpt_a, _b and _c (ic.py:1619-1893).  P1: the second derivatives of Φ⁽¹⁾ (from carryout_2lpt);
    ctx's saved slab holds Φ⁽²⁾ on entry."""
    g = growth_factors
    fft_factor = float(ctx_dealias.gridsize)**(-3)
    mom_factor = _mom_factor(component, a)
    aH = a*hubble(a)
    P2 = _second_derivatives(ctx, ctx_dealias)
    # 'a' term: the determinant of the Hessian of Φ⁽¹⁾
    _lpt_potential(ctx, ctx_dealias, [
        (+1, [P1[0, 2], P1[0, 2], P1[1, 1]]), (-1, [P1[1, 1], P1[2, 2], P1[0, 0]]), (+1, [P1[0, 0], P1[1, 2], P1[1, 2]]),
        (-2, [P1[1, 2], P1[0, 2], P1[0, 1]]), (+1, [P1[0, 1], P1[0, 1], P1[2, 2]]),
    ], fft_factor*g['D3a']/g['D1']**3)
    _displace_from_saved(component, ctx, index_bgn, 1.0, aH*g['f3a']*mom_factor)
    # 'b' term
    _lpt_potential(ctx, ctx_dealias, [
        (-.5, [P1[2, 2], P2[0, 0]]), (-.5, [P2[0, 0], P1[1, 1]]), (-.5, [P1[1, 1], P2[2, 2]]), (-.5, [P2[2, 2], P1[0, 0]]),
        (-.5, [P1[0, 0], P2[1, 1]]), (-.5, [P2[1, 1], P1[2, 2]]),
        (+1, [P2[0, 2], P1[0, 2]]), (+1, [P2[0, 1], P1[0, 1]]), (+1, [P2[1, 2], P1[1, 2]]),
    ], fft_factor*g['D3b']/(g['D1']*g['D2']))
    _displace_from_saved(component, ctx, index_bgn, 1.0, aH*g['f3b']*mom_factor)
    # 'c' term: the transverse displacement, the curl of the vector potential A⁽³ᶜ⁾
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        _lpt_potential(ctx, ctx_dealias, [
            (+1, [P2[j, j], P1[j, k]]), (-1, [P1[j, k], P2[k, k]]), (-1, [P1[i, j], P2[i, k]]),
            (-1, [P1[j, j], P2[j, k]]), (+1, [P2[j, k], P1[k, k]]), (+1, [P2[i, j], P1[i, k]]),
        ], fft_factor*g['D3c']/(g['D1']*g['D2']))
        dims = []
        for target in range(3):
            if target == i:
                continue
            along = ({0, 1, 2} - {i, target}).pop()
            dims.append((target, along, 2*(along == (target + 1) % 3) - 1))
        _displace_from_saved(component, ctx, index_bgn, 1.0, aH*g['f3c']*mom_factor, dims)

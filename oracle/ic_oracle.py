"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy / plain Python loops) of the reference's particle
initial-condition generator, pinned to golden vectors made by the unmodified reference
(tests/golden/gen_golden_ic.py → tests/golden/ic_*.npz; tests/test_widen_ic.py).  Parity PINNED.

Nothing here is imported by the product (concept_b200/); only tests/ use it as the checker.

Follows (all under /root/reference/src/):
  ic.py:67-232      PseudoRandomNumberGenerator (seed salt, spawn, cached draws)
  ic.py:928-1163    generate_primordial_noise ('distributed' and 'simple' imprinting)
  mesh.py:2909-3044 fourier_curve_loop / fourier_curve_slice_loop, :3109-3160 get_fourier_curve_coords
  ic.py:542-627     get_amplitudes;  linear.py:3329-3341 get_primordial_curvature_perturbation
  ic.py:670-782     realize_grid (scalar realisations, lattice phase shift, local non-Gaussianity)
  mesh.py:3422-3437 laplacian_inverse, :3470-3510 fourier_diff, :3591-3622 nullify_modes
  ic.py:1447-1509   carryout_1lpt, :1539-1589 carryout_2lpt, :1619-1893 carryout_3lpt_a/b/c,
                    :1895-2057 handle_lpt_term (incl. dealiasing)
  ic.py:2138-2247   preinitialize_particles, :2249-2283 displace_particles, :1396-1398 wrap

Fourier slabs are complex arrays in the reference's transposed layout [j][i][kk], kk = 0 … G/2.
"""
import math

import numpy as np

π = math.pi


# ---------------------------------------------------------------------------------- random numbers
class PRNG:
    """ic.py:67-232.  Draws come in cached batches of `cache_size` like the reference's."""

    def __init__(self, seed, cache_size=2**12, salt=True):
        if salt and isinstance(seed, (int, np.integer)):
            seed = int(seed) + int(π*1e+8) + 137          # ic.py:108-116
        if not isinstance(seed, np.random.SeedSequence):
            seed = np.random.SeedSequence(seed)
        self.seed = seed
        self.cache_size = cache_size
        self.generator = np.random.Generator(np.random.PCG64DXSM(seed))
        self.cache_u = self.cache_r = None
        self.iu = self.ir = cache_size - 1

    def spawn(self, spawn_key):
        seed = np.random.SeedSequence(self.seed.entropy, spawn_key=self.seed.spawn_key + (int(spawn_key), ))
        return PRNG(seed, self.cache_size)

    def uniform(self, low, high):
        self.iu += 1
        if self.iu == self.cache_size:
            self.iu = 0
            self.cache_u = self.generator.uniform(0, 1, size=self.cache_size)
        return low + self.cache_u[self.iu]*(high - low)

    def rayleigh(self, scale):
        self.ir += 1
        if self.ir == self.cache_size:
            self.ir = 0
            self.cache_r = self.generator.rayleigh(1, size=self.cache_size)
        return self.cache_r[self.ir]*scale


# ------------------------------------------------------------------------ Fourier space-filling curve
def icbrt(x):
    r = int(round(x**(1/3)))
    while r**3 > x:
        r -= 1
    while (r + 1)**3 <= x:
        r += 1
    return r


def fourier_curve_coords(key):
    """mesh.py:3109-3160"""
    key = int(key)
    g = icbrt(2*key)
    g += g & 1
    g += -(g**2*(g//2 + 1) <= key) & 2
    s = g//2
    key -= (g - 2)**2*s
    f0 = 2*s**2 - s
    if key < f0:
        return s - 1, -s + 1 + key//s, key % s
    f1 = 2*s**2 - 2*s + f0
    if key < f1:
        key -= f0
        return -s + 1 + key//s, s - 1, key % s
    f2 = f0 + f1
    if key < f2:
        key -= f1
        return -s, -s + 1 + key//s, key % s
    f3 = 2*s**2 + f2
    if key < f3:
        key -= f2
        return -s + key//s, -s, key % s
    key -= f3
    return -s + key//(2*s), -s + key % (2*s), s


def fourier_curve_slice(G):
    """mesh.py:2984-3044: the (ki, kk) of the kj = 0 slice in curve order (same order for every slice)."""
    nyq = G//2
    out = []
    for s in range(nyq):
        for f in range(1 + (2 if s < nyq - 1 else 0)):
            key_bgn = f**2 + (f == 2) + (1 + 6*f - (f == 1))*s + (5 + 4*f - (f == 2))*s**2 + 4*s**3
            num = (s + 1)*(1 + (f == 2))
            step = 1 + ((num - 1) if f == 2 else 0)
            key = key_bgn - step
            for _ in range(num):
                key += step
                ki, kj, kk = fourier_curve_coords(key)
                assert kj == 0
                out.append((ki, kk))
    return out


def primordial_noise(G, seed_amplitudes=1000, seed_phases=2000, fixed_amplitude=False, phase_shift=0.0,
                     imprinting='distributed'):
    """ic.py:928-1163 on one rank.  Returns complex [j][i][G/2+1]; the Nyquist planes are left as they
    are (zero here), the origin is nullified."""
    nyq = G//2
    slab = np.zeros((G, G, nyq + 1), dtype=complex)
    amp_common = PRNG(seed_amplitudes)
    pha_common = PRNG(seed_phases)
    if imprinting == 'simple':
        n_total = G*G*(nyq + 1)
        n_nyquist = G**2 + nyq*(2*G - 1)
        for key in range(n_total - n_nyquist):
            ki, kj, kk = fourier_curve_coords(key)
            r = 1.0 if fixed_amplitude else amp_common.rayleigh(1/math.sqrt(2))
            θ = pha_common.uniform(-π, π)
            imprint, imprint_conj = True, False
            if kk == 0:
                lower = (ki < 0) or (ki == 0 and kj < 0)
                imprint = lower
                imprint_conj = lower
            if not imprint and not imprint_conj:
                continue
            if phase_shift:
                θ += phase_shift
            re, im = r*math.cos(θ), r*math.sin(θ)
            if imprint:
                slab[kj % G, ki % G, kk] = complex(re, im)
            if imprint_conj:
                slab[(-kj) % G, (-ki) % G, 0] = complex(re, -im)
    elif imprinting == 'distributed':
        offset = 2**32
        order = fourier_curve_slice(G)
        for j in range(G):
            kj = j - (G if j >= nyq else 0)
            if kj == -nyq:
                continue
            amp, amp_c = amp_common.spawn(offset + kj), amp_common.spawn(offset - kj)
            pha, pha_c = pha_common.spawn(offset + kj), pha_common.spawn(offset - kj)
            kj_conj = -kj
            for ki, kk in order:
                r = r_c = 1.0
                if not fixed_amplitude:
                    r = amp.rayleigh(1/math.sqrt(2))
                    r_c = amp_c.rayleigh(1/math.sqrt(2))
                θ = pha.uniform(-π, π)
                θ_c = pha_c.uniform(-π, π)
                imprint, imprint_conj = True, False
                if kk == 0:
                    lower = (ki < 0) or (ki == 0 and kj < 0)
                    lower_conj = (ki > 0) or (ki == 0 and kj_conj > 0)
                    imprint = lower
                    imprint_conj = not lower_conj
                if imprint:
                    if phase_shift:
                        θ += phase_shift
                    slab[j, ki % G, kk] = complex(r*math.cos(θ), r*math.sin(θ))
                if imprint_conj:
                    if phase_shift:
                        θ_c += phase_shift
                    # reflection of (ki, −kj) is (−ki, kj): lands in this very slice
                    slab[(-kj_conj) % G, (-ki) % G, 0] = complex(r_c*math.cos(θ_c), -r_c*math.sin(θ_c))
    else:
        raise ValueError(imprinting)
    slab[0, 0, 0] = 0
    return slab


# --------------------------------------------------------------------------------------- k-space helpers
def _k_grids(G):
    kj = np.fft.fftfreq(G, 1/G).astype(np.int64)[:, None, None]
    ki = np.fft.fftfreq(G, 1/G).astype(np.int64)[None, :, None]
    kk = np.arange(G//2 + 1, dtype=np.int64)[None, None, :]
    return ki, kj, kk


def _mask(G):
    """Modes visited by fourier_loop(skip_origin=True): everything but the Nyquist planes and the origin."""
    ki, kj, kk = _k_grids(G)
    m = (np.abs(ki) < G//2) & (np.abs(kj) < G//2) & (kk < G//2)
    m = np.broadcast_to(m, (G, G, G//2 + 1)).copy()
    m[0, 0, 0] = False
    return m


def zeta(k, A_s, n_s, alpha_s, pivot):
    """linear.py:3329-3341"""
    return (π*math.sqrt(2*A_s)/pivot**((n_s - 1)/2))*k**(n_s/2 - 2)*np.exp(alpha_s/4*(np.log(k) - math.log(pivot))**2)


def get_amplitudes(G, boxsize, transfer, primordial, factor=1.0):
    """ic.py:599-627 (primordial structure): table over integer k² of T(k)·ζ(k)·L^(−3/2)·factor"""
    nyq = G//2
    k2_max = 3*(nyq - 1)**2
    amp = np.zeros(k2_max + 1)
    normalization = boxsize**(-1.5)*factor
    k2 = np.arange(1, k2_max + 1)
    k = (2*π/boxsize)*np.sqrt(k2)
    amp[1:] = transfer(k)*zeta(k, **primordial)*normalization
    return amp


def realize_grid(noise, amplitudes, shift=(0, 0, 0)):
    """ic.py:670-782, scalar realisation in Fourier space.  `shift` is the particle lattice shift in grid
    units; realize_grid negates it (:692-695) and fourier_loop uses θ = −2π/G·k·shift' (mesh.py:2873-2888),
    so θ = +2π/G·(k·shift)."""
    G = noise.shape[0]
    ki, kj, kk = _k_grids(G)
    m = _mask(G)
    k2 = ki**2 + kj**2 + kk**2
    k2c = np.where(m, k2, 0)
    slab = amplitudes[np.minimum(k2c, len(amplitudes) - 1)]*noise
    if tuple(shift) != (0, 0, 0):
        θ = (2*π/G)*(ki*shift[0] + kj*shift[1] + kk*shift[2])
        slab = slab*(np.cos(θ) + 1j*np.sin(θ))
    return np.where(m, slab, 0)


def laplacian_inverse(slab, boxsize, factor=1.0):
    """mesh.py:3422-3437: ×(−factor/k_f²)/k² on the visited modes, the others untouched"""
    G = slab.shape[0]
    ki, kj, kk = _k_grids(G)
    m = _mask(G)
    k2 = np.where(m, ki**2 + kj**2 + kk**2, 1)
    kf = 2*π/boxsize
    return np.where(m, slab*((-factor/kf**2)/k2), slab)


def fourier_diff(slab, boxsize, dim0, dim1=-1, factor=1.0):
    """mesh.py:3470-3510: ×i·k_dim0 (and ×i·k_dim1) on the visited modes, zero elsewhere"""
    G = slab.shape[0]
    ks = _k_grids(G)
    m = _mask(G)
    kf = 2*π/boxsize
    out = slab*1j*(factor*kf)*ks[dim0] if dim1 == -1 else slab*(-1.0)*(factor*kf**2)*ks[dim0]*ks[dim1]
    return np.where(m, out, 0)


def backward(slab):
    """fft(slab, 'backward') (mesh.py:4012-4157): unnormalised c2r; real grid [i][j][k]"""
    G = slab.shape[0]
    return np.fft.irfftn(slab.transpose(1, 0, 2), s=(G, G, G), axes=(0, 1, 2))*float(G)**3


def forward(grid):
    """fft(grid, 'forward'): unnormalised r2c into the transposed layout [j][i][kk]"""
    return np.fft.rfftn(grid, axes=(0, 1, 2)).transpose(1, 0, 2).copy()


def resize_fourier(slab, G_new):
    """resize_grid(…, 'fourier') restricted to what the LPT code needs: copy the modes |k| < min(G, G')/2"""
    G = slab.shape[0]
    n = min(G, G_new)//2
    out = np.zeros((G_new, G_new, G_new//2 + 1), dtype=complex)
    idx = np.arange(-n + 1, n)
    out[np.ix_(idx % G_new, idx % G_new, np.arange(n))] = slab[np.ix_(idx % G, idx % G, np.arange(n))]
    return out


# ----------------------------------------------------------------------------------------- realisation
LATTICE_SHIFTS = {   # mesh.py:85-100 with cell_centered = True ⇒ shift_amount = −½
    1: [(0, 0, 0)],
    2: [(0, 0, 0), (-.5, -.5, -.5)],
    4: [(0, 0, 0), (0, -.5, -.5), (-.5, 0, -.5), (-.5, -.5, 0)],
}


def realize_particles(n, lattices, boxsize, a, H, mass, w_eff, noise, transfer_delta, transfer_theta, primordial,
                      backscale=False, lpt=1, dealias=False, growth=None, nongaussianity=0.0, shifts=None):
    """ic.py:1199-1399.  noise: complex [j][i][kk] from primordial_noise(n, …).  Returns pos, mom [N][3].
    shifts: the primitive lattices to realise (default: all of the sc / bcc / fcc lattice); a simple-cubic component
    that shares the box with one or three others gets a single primitive lattice of bcc / fcc (ic.py:1248-1253)."""
    G = n
    N1 = n**3
    pos = np.zeros((lattices*N1, 3))
    mom = np.zeros((lattices*N1, 3))
    cell = boxsize/G
    mom_factor = a*(a**(-3*w_eff)*mass)            # displace_particles, ic.py:2269-2271
    Gd = G
    if dealias:
        Gd = (G*3)//2
        Gd += Gd & 1
    for l, shift in enumerate(LATTICE_SHIFTS[lattices] if shifts is None else shifts):
        sl = slice(l*N1, (l + 1)*N1)
        # preinitialize_particles (ic.py:2138-2247), cell-centred
        ax = [(0.5 + shift[d] + np.arange(G))*cell for d in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing='ij')
        pos[sl] = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        # carryout_1lpt (ic.py:1447-1509)
        for variable in range(1 - backscale, -1, -1):
            amplitudes = get_amplitudes(G, boxsize, transfer_theta if variable == 1 else transfer_delta, primordial)
            slab = realize_grid(noise, amplitudes, shift)
            if nongaussianity and variable == 0:
                # realize_grid, ic.py:766-776: x += f·x² in real space, back with the forward normalisation G⁻³
                real = backward(slab)
                real = real + nongaussianity*real**2
                slab = forward(real)*float(G)**(-3)
            Φ1 = laplacian_inverse(slab, boxsize, factor=2*variable - 1)
            for d in range(3):
                ψ = backward(fourier_diff(Φ1, boxsize, d)).ravel()
                if variable == 0:
                    pos[sl, d] += ψ
                else:
                    mom[sl, d] += mom_factor*ψ
                if backscale:
                    mom[sl, d] += (a*H*growth['f1']*mom_factor)*ψ
        if lpt >= 2:
            # carryout_2lpt (ic.py:1539-1589); Φ1 is the δ potential (last loop iteration above)
            fft_factor = float(Gd)**(-3)
            potential_factor = fft_factor*growth['D2']/growth['D1']**2
            velocity_factor = a*H*growth['f2']

            def dd(i, j):
                s = fourier_diff(Φ1, boxsize, i, j)
                return backward(resize_fourier(s, Gd) if dealias else s)

            terms = [((0, 0), (1, 1), -1), ((1, 1), (2, 2), -1), ((2, 2), (0, 0), -1),
                     ((0, 1), (0, 1), +1), ((1, 2), (1, 2), +1), ((0, 2), (0, 2), +1)]
            Φ2 = 0
            for p, q, sign in terms:
                prod = dd(*p)*dd(*q)
                Φ2 = Φ2 + sign*(resize_fourier(forward(prod), G) if dealias else prod)
            if not dealias:
                Φ2 = forward(Φ2)
            Φ2 = laplacian_inverse(Φ2, boxsize, potential_factor)
            for d in range(3):
                ψ = backward(fourier_diff(Φ2, boxsize, d)).ravel()
                pos[sl, d] += ψ
                mom[sl, d] += (velocity_factor*mom_factor)*ψ
        if lpt >= 3:
            # carryout_3lpt_a / _b / _c (ic.py:1619-1893) through handle_lpt_term (:1895-2057): products are formed
            # pairwise in the order written; with dealiasing every intermediate product is cut back to the cube
            # |k| < G/2 (forward, nullify, backward — the two unnormalised transforms cost the factor Gd⁻³).
            def dd2(i, j):
                s = fourier_diff(Φ2, boxsize, i, j)
                return backward(resize_fourier(s, Gd) if dealias else s)

            def cut(prod):
                if not dealias:
                    return prod
                return backward(resize_fourier(resize_fourier(forward(prod), G), Gd))*fft_factor

            def potential(terms, potential_factor):
                total = 0
                for factor, grids in terms:
                    prod = grids[0]*grids[1]
                    for g in grids[2:]:
                        prod = cut(prod)*g
                    total = total + factor*(resize_fourier(forward(prod), G) if dealias else prod)
                if not dealias:
                    total = forward(total)
                return laplacian_inverse(total, boxsize, potential_factor)

            def displace(Φ, velocity_factor, dims=((0, 0, 1), (1, 1, 1), (2, 2, 1))):
                for target, along, sign in dims:
                    ψ = backward(fourier_diff(Φ, boxsize, along, factor=sign)).ravel()
                    pos[sl, target] += ψ
                    mom[sl, target] += (velocity_factor*mom_factor)*ψ

            P1 = {ij: dd(*ij) for ij in ((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2))}
            P2 = {ij: dd2(*ij) for ij in ((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2))}
            P1.update({(j, i): v for (i, j), v in list(P1.items())})
            P2.update({(j, i): v for (i, j), v in list(P2.items())})
            Φ3a = potential([(+1, [P1[0, 2], P1[0, 2], P1[1, 1]]), (-1, [P1[1, 1], P1[2, 2], P1[0, 0]]),
                             (+1, [P1[0, 0], P1[1, 2], P1[1, 2]]), (-2, [P1[1, 2], P1[0, 2], P1[0, 1]]),
                             (+1, [P1[0, 1], P1[0, 1], P1[2, 2]])], fft_factor*growth['D3a']/growth['D1']**3)
            displace(Φ3a, a*H*growth['f3a'])
            Φ3b = potential([(-.5, [P1[2, 2], P2[0, 0]]), (-.5, [P2[0, 0], P1[1, 1]]), (-.5, [P1[1, 1], P2[2, 2]]),
                             (-.5, [P2[2, 2], P1[0, 0]]), (-.5, [P1[0, 0], P2[1, 1]]), (-.5, [P2[1, 1], P1[2, 2]]),
                             (+1, [P2[0, 2], P1[0, 2]]), (+1, [P2[0, 1], P1[0, 1]]), (+1, [P2[1, 2], P1[1, 2]])],
                            fft_factor*growth['D3b']/(growth['D1']*growth['D2']))
            displace(Φ3b, a*H*growth['f3b'])
            for i in range(3):
                j, k = (i + 1) % 3, (i + 2) % 3
                A3c = potential([(+1, [P2[j, j], P1[j, k]]), (-1, [P1[j, k], P2[k, k]]), (-1, [P1[i, j], P2[i, k]]),
                                 (-1, [P1[j, j], P2[j, k]]), (+1, [P2[j, k], P1[k, k]]), (+1, [P2[i, j], P1[i, k]])],
                                fft_factor*growth['D3c']/(growth['D1']*growth['D2']))
                dims = []
                for jj in range(3):
                    if jj == i:
                        continue
                    kk = ({0, 1, 2} - {i, jj}).pop()
                    dims.append((jj, kk, 2*(kk == (jj + 1) % 3) - 1))
                displace(A3c, a*H*growth['f3c'], dims)
    pos = np.mod(pos, boxsize)
    return pos, mom

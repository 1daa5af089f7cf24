"""TEST INFRASTRUCTURE ONLY — CPU (numpy) restatement of CO*N*CEPT's particle-mesh
gravity hot path.  It is the *checker* for the CUDA path; the product
(concept_b200/) never imports it.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may use it.

Parity status: PINNED.  Every function below is checked in
tests/test_oracle_golden.py against tests/golden/*.npz, which hold outputs of the
unmodified reference itself run in its pure-Python mode (tests/golden/gen_golden.py,
oracle/ref_sandbox.py): final momenta after interactions.gravity('pm'|'p3m', …,
'long-range') for NGP/CIC/TSC/PCS, differentiation orders 1,2,4,6,8 and 'fourier',
interlacing, no-deconvolution, edge-positioned and clustered particles, plus taps
of the density, potential and force grids, and Component.drift.

All citations are file:line under /root/reference/src/.  Everything is float64
like the reference (C2np['double'] everywhere, e.g. mesh.py:3794).
"""
import numpy as np

MACHINE_EPS = float(np.finfo(np.float64).eps)  # commons.py:1814 machine_ϵ
NGHOSTS_REF = 2  # commons.py:4411-4432 with default P(k) options; enters only through (1±ε) fudge


# --------------------------------------------------------------------------
# Coordinates and weights
# --------------------------------------------------------------------------
def grid_coords(pos, boxsize, gridsize, shift=(0.0, 0.0, 0.0), for_gather=False, nghosts=NGHOSTS_REF):
    """Ghosted grid-unit coordinate of every particle, exactly as the reference
    forms it (single domain, domain_bgn = 0, cell_centered = True).

    deposit: mesh.py:1577-1606   offset = -(1+ε)(ng - ½ - shift)·h ; x = (pos - offset)·((1/h)(1-ε))
    gather : mesh.py:405-432     same with +shift (sign of the lattice shift reversed)
    Returns x[N,3] with  ng - ½ < x < G + ng - ½ .
    """
    pos = np.asarray(pos, dtype=np.float64)
    cellsize = boxsize/gridsize
    sgn = +1.0 if for_gather else -1.0
    out = np.empty_like(pos)
    scale = (1/cellsize)*(1 - MACHINE_EPS)
    for d in range(3):
        offset = 0.0 - (1 + MACHINE_EPS)*(nghosts - 0.5 + sgn*shift[d])*cellsize
        out[:, d] = (pos[:, d] - offset)*scale
    return out


def weights_1d(x, order):
    """Base index and the `order` weights along one axis.
    NGP mesh.py:5305-5308, CIC :5319-5324, TSC :5338-5348, PCS :5365-5379.
    x ≥ 0 always (ghosted frame) so int() == floor()."""
    x = np.asarray(x, dtype=np.float64)
    if order == 1:
        index = (x + 0.5).astype(np.int64)
        w = [np.ones_like(x)]
    elif order == 2:
        index = x.astype(np.int64)
        dist = x - index
        w = [1 - dist, dist]
    elif order == 3:
        index = (x + 0.5).astype(np.int64)
        dist = x - index
        index = index - 1
        dist2 = dist**2
        w0 = 0.125 + 0.5*(dist2 - dist)
        w1 = 0.75 - dist2
        w = [w0, w1, 1 - w0 - w1]
    elif order == 4:
        index = x.astype(np.int64) - 1
        dist = x - index
        tmp = 2 - dist
        tmp2 = tmp**2
        tmp3 = tmp*tmp2
        w0 = 1./6.*tmp3
        w2 = 2./3. - tmp2 + 0.5*tmp3
        w3 = 1./6.*(dist - 1)**3
        w = [w0, 1 - w0 - w2 - w3, w2, w3]
    else:
        raise ValueError(f'order = {order} ∉ {{1, 2, 3, 4}}')
    return index, w


# --------------------------------------------------------------------------
# Deposit (a2) and gather (a12)
# --------------------------------------------------------------------------
def deposit(pos, boxsize, gridsize, order, contribution, shift=(0.0, 0.0, 0.0), nghosts=NGHOSTS_REF):
    """interpolate_particles (mesh.py:1512-1636) + communicate_ghosts('+=')
    (communication.py:563-660) on one rank == periodic wrap.  `contribution` is the
    per-particle scalar  q·mass·(G/L)³·G⁻³  (mesh.py:1550-1573, 582).
    Weight = (wx·contribution)·wy·wz  (mesh.py:5143-5154).  Returns rho[G,G,G]."""
    G = int(gridsize)
    x = grid_coords(pos, boxsize, G, shift, for_gather=False, nghosts=nghosts)
    ix, wx = weights_1d(x[:, 0], order)
    iy, wy = weights_1d(x[:, 1], order)
    iz, wz = weights_1d(x[:, 2], order)
    rho = np.zeros(G**3, dtype=np.float64)
    for a in range(order):
        ia = (ix + a - nghosts) % G
        wa = wx[a]*contribution
        for b in range(order):
            ib = (iy + b - nghosts) % G
            wab = wa*wy[b]
            for c in range(order):
                ic = (iz + c - nghosts) % G
                rho += np.bincount((ia*G + ib)*G + ic, weights=wab*wz[c], minlength=G**3)
    return rho.reshape(G, G, G)


def gather(grid, pos, boxsize, order, shift=(0.0, 0.0, 0.0), nghosts=NGHOSTS_REF):
    """interpolate_domaingrid_to_particles (mesh.py:376-459): value = Σ grid[cell]·w,
    w = (wx·wy)·wz, accumulated in i,j,k-major order.  `grid` is ghost-free [G,G,G]."""
    G = grid.shape[0]
    x = grid_coords(pos, boxsize, G, shift, for_gather=True, nghosts=nghosts)
    ix, wx = weights_1d(x[:, 0], order)
    iy, wy = weights_1d(x[:, 1], order)
    iz, wz = weights_1d(x[:, 2], order)
    flat = grid.reshape(-1)
    value = np.zeros(len(pos), dtype=np.float64)
    for a in range(order):
        ia = (ix + a - nghosts) % G
        for b in range(order):
            ib = (iy + b - nghosts) % G
            wab = wx[a]*wy[b]
            for c in range(order):
                ic = (iz + c - nghosts) % G
                value += flat[(ia*G + ib)*G + ic]*(wab*wz[c])
    return value


# --------------------------------------------------------------------------
# k-space (a7, a8, a9, a10)
# --------------------------------------------------------------------------
def signed_wavenumbers(G):
    k = np.arange(G)
    return k - G*(k >= G//2)   # ki = i - G·[i ≥ G/2]; i = G/2 ↦ -G/2 (mesh.py:2720-2742)


def mode_mask(G):
    """True for modes visited by fourier_loop: all except the Nyquist planes ki = -G/2,
    kj = -G/2, kk = +G/2 (mesh.py:2720-2742, 2833; nullify_modes 'nyquist' :3591-3622).
    Natural rfftn layout [i, j, kk]."""
    ki = signed_wavenumbers(G)
    m1 = ki != -(G//2)
    mk = np.arange(G//2 + 1) != G//2
    return m1[:, None, None] & m1[None, :, None] & mk[None, None, :]


def deconv_factor(G, deconv_order):
    """[Π x_l/sin x_l]^D with x_l = k_l·π/G + ε (mesh.py:2775-2776, 2795-2798, 2848-2856),
    in the reference's operation order ((xi·xj)·xk)/((si·sj)·sk) then **D."""
    ki = signed_wavenumbers(G).astype(np.float64)
    kk = np.arange(G//2 + 1, dtype=np.float64)
    xi = ki*(np.pi/G) + MACHINE_EPS
    xk = kk*(np.pi/G) + MACHINE_EPS
    if not deconv_order:
        return np.ones((G, G, G//2 + 1))
    num = (xi[:, None, None]*xi[None, :, None])*xk[None, None, :]
    den = (np.sin(xi)[:, None, None]*np.sin(xi)[None, :, None])*np.sin(xk)[None, None, :]
    return (num/den)**deconv_order


def k2_grid(G):
    ki = signed_wavenumbers(G)
    kk = np.arange(G//2 + 1)
    return (ki[None, :, None]**2 + ki[:, None, None]**2) + kk[None, None, :]**2


def potential_factor(G, boxsize, G_Newton, deconv_order, r_scale=0.0, n_lattices=1):
    """Per-mode real factor applied in particle_mesh (interactions.py:2092-2118):
    factor = deconv/n_lattices · (−L²·G_N/π)/k² [· exp(−k²(2π r_s/L)²)], origin and
    Nyquist planes zero.  Natural layout [i, j, kk]."""
    factor = deconv_factor(G, deconv_order)*(1/n_lattices)
    k2 = k2_grid(G).astype(np.float64)
    k2[0, 0, 0] = 1.0
    pot = (-boxsize**2*G_Newton/np.pi)/k2
    if r_scale:
        pot = pot*np.exp(k2*(-(2*np.pi/boxsize*r_scale)**2))
    factor = factor*pot
    factor[0, 0, 0] = 0.0
    factor[~mode_mask(G)] = 0.0
    return factor


def interlace_phase(G, shift):
    """θ = −(2π/G)(ki·sx + kj·sy + kk·sz) (mesh.py:2873-2888); returns e^{iθ} [i,j,kk]."""
    ki = signed_wavenumbers(G).astype(np.float64)
    kk = np.arange(G//2 + 1, dtype=np.float64)
    theta = ((ki*(-2*np.pi/G*shift[0]))[:, None, None] + (ki*(-2*np.pi/G*shift[1]))[None, :, None]) \
        + (kk*(-2*np.pi/G*shift[2]))[None, None, :]
    return np.cos(theta) + 1j*np.sin(theta)


def forward_fft(rho):
    """Unnormalised r2c (mesh.py:4078 np.fft.rfftn norm='backward'); natural layout."""
    return np.fft.rfftn(rho)


def backward_fft(slab, G):
    """Unnormalised c2r (mesh.py:4131 irfftn norm='forward')."""
    return np.fft.irfftn(slab, s=(G, G, G), axes=(0, 1, 2), norm='forward')


def to_reference_slab_layout(slab_natural):
    """[i, j, kk] complex → the reference's transposed, interleaved double layout
    slab[j, i, 2·kk + {0: re, 1: im}] of shape (G, G, G+2) (fft.c:55-72; mesh.py:4081-4093)."""
    t = np.ascontiguousarray(slab_natural.transpose(1, 0, 2))
    return t.view(np.float64).reshape(t.shape[0], t.shape[1], 2*t.shape[2])


# --------------------------------------------------------------------------
# Finite differences (a11)
# --------------------------------------------------------------------------
FD_COEFFS = {  # mesh.py:4961-5015
    2: [1/2],
    4: [2/3, -1/12],
    6: [3/4, -3/20, 1/60],
    8: [4/5, -1/5, 4/105, -1/280],
}


def diff_grid(phi, dim, order, dx):
    """diff_domaingrid (mesh.py:4874-5030), periodic.  order 1 = forward one-sided."""
    if order == 1:
        return (1/dx)*(np.roll(phi, -1, dim) - phi)
    out = None
    for n, c in enumerate(FD_COEFFS[order], start=1):
        term = (abs(c)/dx)*(np.roll(phi, -n, dim) - np.roll(phi, n, dim))
        if out is None:
            out = term
        else:
            out = out + term if c > 0 else out - term
    return out


# --------------------------------------------------------------------------
# One long-range PM kick (a16 orchestration of a2…a12)
# --------------------------------------------------------------------------
BCC_SHIFT = -0.5  # Lattice.shift_amount for cell-centred grids (mesh.py:85)


def pm_kick(pos, mom, *, mass, boxsize, gridsize, order, G_Newton, dt_rho_over_dt1, dt_kick,
            diff_order=2, deconvolve=True, interlace=False, r_scale=0.0, taps=None):
    """interactions.gravity('pm'|'p3m', [c], [c], ᔑdt, 'long-range') for one particle
    component (interactions.py:2854-2961 → particle_mesh :1985-2335).

    dt_rho_over_dt1 = ᔑdt['a**(-3*w_eff-1)', name]/ᔑdt['1']   (mesh.py:1556)
    dt_kick         = ᔑdt['a**(-3*w_eff)', name]              (interactions.py:2384-2387)
    r_scale = shortrange_params['gravity']['scale'] for 'gravity long-range', else 0.
    Returns the new momenta (pos unchanged)."""
    G = int(gridsize)
    pos = np.asarray(pos, dtype=np.float64)
    mom = np.array(mom, dtype=np.float64, copy=True)
    h = boxsize/G
    # mesh.py:1556-1573: contribution = q·mass · factor·(G/L)³ with factor = G⁻³ (mesh.py:582)
    contribution = dt_rho_over_dt1
    contribution *= mass
    contribution_factor = float(G)**(-3)*(G/boxsize)**3
    contribution *= contribution_factor
    shifts = [(0.0, 0.0, 0.0)] + ([(BCC_SHIFT,)*3] if interlace else [])
    nl = len(shifts)
    deconv_global = order*(2 if deconvolve else 0)   # interactions.py:2069-2080 (both promoted)
    # Upstream: deposit each lattice, FFT, Nyquist-nullify, interlace-combine (mesh.py:598-616, 654-710)
    slab = None
    for n, s in enumerate(shifts):
        rho = deposit(pos, boxsize, G, order, contribution, s)
        if taps is not None:
            taps['rho' if n == 0 else 'rho_shifted'] = rho
        f = forward_fft(rho)
        f[~mode_mask(G)] = 0
        if nl > 1:
            f = f*(1/nl)
            if n > 0:
                f = f*interlace_phase(G, s)
        slab = f if slab is None else slab + f
    slab = slab*potential_factor(G, boxsize, G_Newton, deconv_global, r_scale)
    kick_factor = mass*(-dt_kick)
    # Downstream: per sub-lattice (interactions.py:2214-2330)
    for n, s in enumerate(shifts):
        f = slab
        if nl > 1:
            f = f*(1/nl)
            if n > 0:
                f = f*interlace_phase(G, s)
        if diff_order == 0:
            # Fourier differentiation: ×i·(2π/L)·k_dim (mesh.py:3383-3392)
            ki = signed_wavenumbers(G).astype(np.float64)
            kvec = [ki[:, None, None], ki[None, :, None], np.arange(G//2 + 1, dtype=np.float64)[None, None, :]]
            for dim in range(3):
                force = backward_fft(1j*(2*np.pi/boxsize)*kvec[dim]*f, G)
                if taps is not None:
                    taps[f'forcegrid{dim}' + ('_shifted' if n else '')] = force
                mom[:, dim] += gather(force, pos, boxsize, order, s)*kick_factor
        else:
            phi = backward_fft(f, G)
            if taps is not None:
                taps['phi' if n == 0 else 'phi_shifted'] = phi
            for dim in range(3):
                force = diff_grid(phi, dim, diff_order, h)
                if taps is not None:
                    taps[f'forcegrid{dim}' + ('_shifted' if n else '')] = force
                mom[:, dim] += gather(force, pos, boxsize, order, s)*kick_factor
    return mom


# --------------------------------------------------------------------------
# Drift (a13) and v_rms (a17)
# --------------------------------------------------------------------------
def drift(pos, mom, dt_over_mass, boxsize):
    """Component.drift (species.py:2191-2196): pos ← mod(pos + mom·Δ, L), mod → [0, L) with
    x == L ↦ 0 (commons.py:5102-5131)."""
    x = np.mod(np.asarray(pos, dtype=np.float64) + np.asarray(mom, dtype=np.float64)*dt_over_mass, boxsize)
    x[x == boxsize] = 0
    return x


def sum_mom2(mom):
    """Σ mom² as in measure(component, 'v_rms') (analysis.py:3965-3972)."""
    m = np.asarray(mom, dtype=np.float64).reshape(-1)
    return float(np.dot(m, m))


def v_rms(mom, N, a, mass, w_eff=0.0):
    return np.sqrt(sum_mom2(mom)/N)/(a**(2 - 3*w_eff)*mass)


# --------------------------------------------------------------------------
# P³M short range (a18, a19)
# --------------------------------------------------------------------------
def softened_r3inv(r2, eps):
    """get_softened_r3inv, spline kernel (interactions.py:1875-1897), h = 2.8·ε."""
    from math import sqrt
    h = 2.8*eps
    r = sqrt(r2)
    if r >= h:
        return 1/(r2*r)
    u = r/h
    if u < 0.5:
        return 32/h**3*(1./3. + u**2*(-6./5. + u))
    return 32/(3*r**3)*(u**3*(2 + u*(-9./2. + u*(18./5. - u))) - 3./480.)


def shortrange_table(scale, rng, tablesize, softening):
    """get_shortrange_table (gravity.py:373-421): tabulated in r² at bin mid-points,
    table[i] = −r⁻³(x/√π·e^{−x²/4} + erfc(x/2) − 1) − r⁻³_softened, x = r/scale; last entry NaN."""
    from math import erfc, exp, pi, sqrt
    maxr2 = (1 + 1/tablesize)*rng**2
    r_tab = np.sqrt(np.linspace(0, maxr2, tablesize))
    table = np.empty(tablesize)
    for i in range(tablesize - 1):
        r2 = 0.5*(r_tab[i]**2 + r_tab[i + 1]**2)
        r = sqrt(r2)
        x = r*(1/scale)
        r3_inv = 1/(r2*r)
        table[i] = -r3_inv*(1/sqrt(pi)*x*exp(-(0.5*x)**2) + (erfc(0.5*x) - 1)) - softened_r3inv(r2, softening)
    table[tablesize - 1] = np.nan
    return table, maxr2


def shortrange_sums(pos, boxsize, rng, table, maxr2, active=None):
    """S_i = Σ_{j≠i, r²≤R²} (x_i − x_j)·table[int(r²·(T−1)/maxr²)] with minimum-image separations
    (gravity_pairwise_shortrange, gravity.py:263-354; pair enumeration interactions.py:1353-1791).
    Brute force O(N²); `active` restricts the receivers."""
    pos = np.asarray(pos, dtype=np.float64)
    N = len(pos)
    T = len(table)
    scaling = (T - 1)/maxr2
    S = np.zeros((N, 3))
    idx = np.arange(N) if active is None else np.nonzero(active)[0]
    tab = np.nan_to_num(table, nan=0.0)
    for i in idx:
        d = pos[i] - pos
        d -= boxsize*np.round(d/boxsize)
        r2 = (d*d).sum(1)
        m = r2 <= rng**2
        m[i] = False
        f = tab[(r2[m]*scaling).astype(np.int64)]
        S[i] = (d[m]*f[:, None]).sum(0)
    return S


def get_rung(acc, current, rung_factor, n_rungs):
    """Component.get_rung (species.py:2340-2360) vectorised."""
    acc2 = (np.asarray(acc)**2).sum(1)
    with np.errstate(divide='ignore'):
        rf = rung_factor + 0.25*np.log2(np.where(acc2 > 0, acc2, 1.0))
    rung = np.where(rf < 0, 0, np.where(rf > n_rungs - 1, n_rungs - 1, 1 + np.floor(np.clip(rf, 0, n_rungs)).astype(np.int64)))
    return np.where(acc2 == 0, current, rung).astype(np.int8)


def rung_factor(dt, fac_softening, softening_length):
    """Component.get_rung_factor (species.py:2362-2370)"""
    return 0.5*np.log2(dt**2/(2*fac_softening*softening_length))


# --------------------------------------------------------------------------
# power spectrum (§8f rank 1): analysis.py:235-579
# --------------------------------------------------------------------------
def sparse_mode_mask(G, k2_max):
    """Modes visited by fourier_loop(G, sparse=True, skip_origin=True, k2_max) (mesh.py:2615-2890): no
    Nyquist planes, no origin, one point of each conjugate pair on the kk = 0 plane (ki > 0 and
    (ki = 0, kj > 0) skipped, mesh.py:2813-2826), k² ≤ k2_max.  Natural layout [i, j, kk]."""
    ki = signed_wavenumbers(G)
    m = mode_mask(G) & (k2_grid(G) <= k2_max)
    skip0 = (ki[:, None] > 0) | ((ki[:, None] == 0) & (ki[None, :] >= 0))
    m[:, :, 0] &= ~skip0
    return m


def density_fourier(pos, mass, a, boxsize, gridsize, order, deconvolve=True, interlace=True, w_eff=0.0):
    """interpolate_upstream(components, …, 'ρ', order, deconvolve, interlace, output_space='Fourier')
    (mesh.py:492-616) for one particle component: for every lattice deposit ρ with the G⁻³ FFT factor,
    forward FFT, nullify Nyquist, deconvolve^order and rotate by the lattice phase, each /n_lattices."""
    G = int(gridsize)
    contribution = a**(-3*(1 + w_eff))*mass
    contribution *= float(G)**(-3)*(G/boxsize)**3
    shifts = [(0.0, 0.0, 0.0), (-0.5, -0.5, -0.5)] if interlace else [(0.0, 0.0, 0.0)]
    slab = np.zeros((G, G, G//2 + 1), dtype=np.complex128)
    for shift in shifts:
        s = forward_fft(deposit(pos, boxsize, G, order, contribution, shift))
        s[~mode_mask(G)] = 0
        s = s*(deconv_factor(G, order*bool(deconvolve))*(1/len(shifts)))
        if any(shift):
            s = s*interlace_phase(G, shift)
        slab += s
    return slab


def power_by_k2(slab, k2_max):
    """The mode loop of compute_powerspec before binning: Σ|δ̂|² and multiplicity per integer k²."""
    G = slab.shape[0]
    m = sparse_mode_mask(G, k2_max)
    k2 = k2_grid(G)[m]
    p = (slab.real**2 + slab.imag**2)[m]
    return np.bincount(k2, weights=p, minlength=k2_max + 1), np.bincount(k2, minlength=k2_max + 1)


# --------------------------------------------------------------------------
# Component-specific upstream / downstream grid sizes (a5 / a10 / a16)
# --------------------------------------------------------------------------
def copy_modes(slab_from, G_onto, deconv_order=0, shift=None, n_lattices=1):
    """copy_modes(slab_from, slab_onto, deconv_order, lattice, '=') (mesh.py:980-1322) onto a nullified slab of
    grid size G_onto; natural layout [i, j, kk].  Deconvolution and interlacing are evaluated for the grid the
    modes come from (fourier_loop(gridsize_small, gridsize_from, …), :1245-1250).  Between different grid sizes
    only the modes strictly inside the smaller grid's Nyquist cube are copied (:1134-1135), and every mode gets
    the phase (π/G_onto − π/G_from)·(ki + kj + kk) of the half-cell offset between two cell-centred grids (:1302)."""
    G_from = slab_from.shape[0]
    ki = signed_wavenumbers(G_from).astype(np.float64)
    kk = np.arange(G_from//2 + 1, dtype=np.float64)
    value = slab_from*(deconv_factor(G_from, deconv_order)*(1/n_lattices))
    theta = None
    if G_from != G_onto:
        theta = (np.pi/G_onto - np.pi/G_from)*((ki[:, None, None] + ki[None, :, None]) + kk[None, None, :])
    if shift is not None and tuple(shift) != (0, 0, 0):
        lattice = ((ki*(-2*np.pi/G_from*shift[0]))[:, None, None] + (ki*(-2*np.pi/G_from*shift[1]))[None, :, None]) \
            + (kk*(-2*np.pi/G_from*shift[2]))[None, None, :]
        theta = lattice if theta is None else theta + lattice
    if theta is not None:
        value = value*(np.cos(theta) + 1j*np.sin(theta))
    if G_from == G_onto:
        value[~mode_mask(G_from)] = 0
        return value
    n = min(G_from, G_onto)//2
    out = np.zeros((G_onto, G_onto, G_onto//2 + 1), dtype=complex)
    idx = np.arange(-n + 1, n)
    out[np.ix_(idx % G_onto, idx % G_onto, np.arange(n))] = value[np.ix_(idx % G_from, idx % G_from, np.arange(n))]
    return out


def pm_kick_multigrid(components, *, boxsize, gridsize_global, order, G_Newton, dt1, deconvolve=True, interlace=False,
                      r_scale=0.0):
    """interactions.gravity('pm'|'p3m', components, components, ᔑdt, 'long-range') for particle components with
    their own (upstream, downstream) grid sizes (particle_mesh, interactions.py:1985-2335; interpolate_upstream,
    mesh.py:492-616; add_upstream_to_global_slabs :618-710).
    components: list of dicts with pos, mom, mass, upstream, downstream, dt_rho (= ᔑdt['a**(-3*w_eff-1)', name]),
    dt_kick (= ᔑdt['a**(-3*w_eff)', name]) and diff_order.  Returns the list of new momenta."""
    Gg = int(gridsize_global)
    shifts = [(0.0, 0.0, 0.0)] + ([(BCC_SHIFT,)*3] if interlace else [])
    nl = len(shifts)
    # which deconvolutions are promoted to the global slab (interactions.py:2069-2080)
    all_up_global = all(c['upstream'] == Gg for c in components)
    all_down_global = all(c['downstream'] == Gg for c in components)
    deconv_up = order*int(deconvolve and not all_up_global)
    deconv_down = order*int(deconvolve and not all_down_global)
    deconv_global = order*(int(deconvolve and all_up_global) + int(deconvolve and all_down_global))
    # upstream
    slab_global = np.zeros((Gg, Gg, Gg//2 + 1), dtype=complex)
    for G_up in sorted({c['upstream'] for c in components}, key=lambda g: (g != Gg, g)):
        for s in shifts:
            rho = np.zeros((G_up, G_up, G_up))
            for c in components:
                if c['upstream'] != G_up:
                    continue
                contribution = c['dt_rho']/dt1
                contribution *= c['mass']
                contribution *= float(G_up)**(-3)*(G_up/boxsize)**3
                rho += deposit(c['pos'], boxsize, G_up, order, contribution, s)
            f = forward_fft(rho)
            f[~mode_mask(G_up)] = 0
            slab_global = slab_global + copy_modes(f, Gg, deconv_up, s, nl)
    slab_global = slab_global*potential_factor(Gg, boxsize, G_Newton, deconv_global, r_scale)
    # downstream
    out = [np.array(c['mom'], dtype=np.float64, copy=True) for c in components]
    for G_down in sorted({c['downstream'] for c in components}, key=lambda g: (g != Gg, g)):
        slab_down = slab_global if G_down == Gg else copy_modes(slab_global, G_down)
        h = boxsize/G_down
        ki = signed_wavenumbers(G_down).astype(np.float64)
        kvec = [ki[:, None, None], ki[None, :, None], np.arange(G_down//2 + 1, dtype=np.float64)[None, None, :]]
        for q, c in enumerate(components):
            if c['downstream'] != G_down:
                continue
            kick_factor = c['mass']*(-c['dt_kick'])
            for s in shifts:
                f = copy_modes(slab_down, G_down, deconv_down, s, nl)     # fourier_operate (mesh.py:3327-3400)
                if c['diff_order'] == 0:
                    for dim in range(3):
                        force = backward_fft(1j*(2*np.pi/boxsize)*kvec[dim]*f, G_down)
                        out[q][:, dim] += gather(force, c['pos'], boxsize, order, s)*kick_factor
                else:
                    phi = backward_fft(f, G_down)
                    for dim in range(3):
                        out[q][:, dim] += gather(diff_grid(phi, dim, c['diff_order'], h), c['pos'], boxsize, order, s)*kick_factor
    return out


def density_fourier_group(components, a, boxsize, gridsize_global, order, deconvolve=True, interlace=True):
    """interpolate_upstream(components, gridsizes_upstream, gridsize_global, 'ρ', order, deconvolve, interlace,
    output_space='Fourier') (mesh.py:492-616) for particle components with their own upstream grid sizes:
    components is a list of dicts with pos, mass, w_eff and upstream.  Natural layout [i, j, kk]."""
    G = int(gridsize_global)
    shifts = [(0.0, 0.0, 0.0)] + ([(BCC_SHIFT,)*3] if interlace else [])
    slab = np.zeros((G, G, G//2 + 1), dtype=complex)
    for G_up in sorted({c['upstream'] for c in components}, key=lambda g: (g != G, g)):
        for s in shifts:
            rho = np.zeros((G_up, G_up, G_up))
            for c in components:
                if c['upstream'] == G_up:
                    contribution = a**(-3*(1 + c['w_eff']))*c['mass']*(float(G_up)**(-3)*(G_up/boxsize)**3)
                    rho += deposit(c['pos'], boxsize, G_up, order, contribution, s)
            f = forward_fft(rho)
            f[~mode_mask(G_up)] = 0
            slab += copy_modes(f, G, order*int(bool(deconvolve)), s, len(shifts))
    return slab

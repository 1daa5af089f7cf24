"""Per-step measurements on the hot path (reference analysis.py:3860-3972).

Only `measure(component, 'v_rms')` is called every time step (by the PM/P³M time-step limiters,
main.py:824-858); the Σmom² reduction runs on the GPU (pm_sum_mom2) and is summed over ranks.
"""
import math

from . import commons


def measure(component, quantity, communicate=True):
    a = commons.universals.a
    if quantity == 'v_rms':
        # analysis.py:3965-3972: sqrt(Σmom²/N)/(a^(2−3w)·mass)
        mom2 = component.sum_mom2() if communicate else component._pm_context().sum_mom2(component.mom_local)
        return math.sqrt(mom2/component.N)/(a**(2 - 3*component.w_eff(a=a))*component.mass)
    commons.abort(f'measure() of "{quantity}" is not on the PM hot path and is not provided by concept_b200')


# ---------------------------------------------------------------------------------------------
# power spectrum (reference analysis.py:60-579)
# ---------------------------------------------------------------------------------------------
import re as _re

import numpy as np

# powerspec_options defaults (commons.py:3354-3385)
POWERSPEC_DEFAULTS = dict(interpolation='PCS', deconvolve=True, interlace=True, k_max='nyquist',
                          bins_per_decade={'  4*k_min': 4, '100*k_min': 40})
_powerspec_bins_cache = {}


def _eval_bin_str(s, mapping):
    """eval_bin_str (analysis.py:435-455): an arithmetic expression in nyquist, gridsize, k_min, k_max,
    k_fundamental, k_f (case-insensitive, with or without the 'k_' prefix)."""
    if not isinstance(s, str):
        return float(s)
    expr = s
    names = {}
    for key, val in mapping.items():
        base = key.removeprefix('k_')
        for k in {key, base, f'k_{base}', f'k{base}'}:
            names[k.lower()] = val
    def repl(m):
        word = m.group(0)
        if word.lower() in names:
            return repr(float(names[word.lower()]))
        if word in ('min', 'max', 'sqrt', 'pi'):
            return word
        commons.abort(f'Cannot evaluate "{s}": unknown name "{word}"')
    expr = _re.sub(r'[A-Za-z_][A-Za-z_0-9]*', repl, expr)
    return float(eval(expr, {'__builtins__': {}}, {'min': min, 'max': max, 'sqrt': math.sqrt, 'pi': math.pi}))


def _controlpoint_spline_log10(d):
    """get_controlpoint_spline(d, np.log10) (commons.py:5436-5465): piecewise-linear in log10(x),
    constant beyond the end points."""
    x = np.array(sorted(d), dtype=float)
    y = np.array([d[k] for k in sorted(d)], dtype=float)
    lx = np.log10(x)
    return lambda x_new: float(np.interp(x_new, lx, y))   # np.interp clamps to y[0] / y[-1] like fill_value


def construct_powerspec_k_bin_centers(k_min, k_max, bins_per_decade, gridsize, nyquist, boxsize):
    """analysis.py:380-432"""
    k_fundamental = 2*math.pi/boxsize
    binsize_min = 0.5*(1 - 1e-2)*k_fundamental*(math.sqrt(3*nyquist**2 + 1) - math.sqrt(3*nyquist**2))
    mapping = {'nyquist': k_fundamental*nyquist, 'gridsize': gridsize, 'k_min': k_min, 'k_max': k_max,
               'k_fundamental': k_min, 'k_f': k_min}
    bpd = {_eval_bin_str(k, mapping): _eval_bin_str(v, mapping) for k, v in bins_per_decade.items()}
    if len(bpd) == 1:
        bpd.update({k + 1: v for k, v in bpd.items()})
    logk_min, logk_max = math.log10(k_min), math.log10(k_max)
    interp = _controlpoint_spline_log10(bpd)
    centers = []
    logk_bin_right = logk_min - 0.5/interp(logk_min)
    while logk_bin_right <= logk_max:
        logk_bin_left = logk_bin_right
        logk_bin_right = logk_bin_left + 1/interp(logk_bin_left)
        logk_bin_right = max(logk_bin_right, math.log10(10**logk_bin_left + binsize_min))
        centers.append(10**(0.5*(logk_bin_left + logk_bin_right)))
    if not centers:
        centers.append(math.sqrt(k_min*k_max))
    centers = np.asarray(centers, dtype=float)
    if len(centers) > 1:
        left = k_min
        right = 10**(logk_max - 0.5/interp(logk_max))
        centers = 10**(math.log10(left) + (np.log10(centers) - math.log10(centers[0]))*(
            (math.log10(right) - math.log10(left))/(math.log10(centers[-1]) - math.log10(centers[0]))))
    return centers


def get_powerspec_bins(gridsize, k_max='nyquist', bins_per_decade=None, n_modes_fine=None, boxsize=None):
    """get_powerspec_bins (analysis.py:235-377) → (k2_max, k_bin_indices, k_bin_centers, n_modes).
    n_modes_fine[k²] is the multiplicity of every k² over the sparse half-space; pass None to get
    (k2_max, provisional k_bin_indices) only — the GPU tallies the multiplicities (pm_power_k2)."""
    boxsize = float(commons.params.boxsize if boxsize is None else boxsize)
    bins_per_decade = dict(POWERSPEC_DEFAULTS['bins_per_decade'] if bins_per_decade is None else bins_per_decade)
    k_fundamental = 2*math.pi/boxsize
    k_min = k_fundamental
    nyquist = gridsize//2
    if isinstance(k_max, str):
        k_max = _eval_bin_str(k_max, {'nyquist': k_fundamental*nyquist, 'gridsize': gridsize, 'k_min': k_min,
                                      'k_fundamental': k_min, 'k_f': k_min})
    k_max = max(k_max, k_min)
    k2_max = min(int(round((k_max/k_fundamental)**2)), 3*nyquist**2)
    k_max = k_fundamental*math.sqrt(k2_max)
    centers = construct_powerspec_k_bin_centers(k_min, k_max, bins_per_decade, gridsize, nyquist, boxsize)
    logc = np.log(centers)
    k2 = np.arange(1, k2_max + 1)
    logk = np.log(k_fundamental*np.sqrt(k2))
    index = np.searchsorted(logc, logk)
    last = index == len(centers)
    index[last] -= 1
    inner = (~last) & (index != 0)
    dist_left = logk - logc[np.maximum(index - 1, 0)]
    dist_right = logc[np.minimum(index, len(centers) - 1)] - logk
    index[inner] -= (dist_left <= dist_right)[inner]
    k_bin_indices = np.zeros(k2_max + 1, dtype=np.int64)
    k_bin_indices[1:] = index
    if n_modes_fine is None:
        return k2_max, k_bin_indices
    n_modes_fine = np.asarray(n_modes_fine, dtype=np.int64)
    # geometric-mean bin centres weighted by multiplicity; drop empty bins (analysis.py:333-366)
    n_modes = np.zeros(len(centers), dtype=np.int64)
    centers_new = np.zeros(len(centers))
    nz = np.nonzero(n_modes_fine[1:])[0] + 1
    np.add.at(n_modes, k_bin_indices[nz], n_modes_fine[nz])
    np.add.at(centers_new, k_bin_indices[nz], n_modes_fine[nz]*np.log(k_fundamental*np.sqrt(nz)))
    good = n_modes > 0
    centers_new[good] = np.exp(centers_new[good]/n_modes[good])
    prev = k_bin_indices[0]
    out = k_bin_indices.copy()
    for q in range(1, len(out)):
        b = k_bin_indices[q]
        if b == prev or n_modes[b] == 0:
            out[q] = out[q - 1]
        elif b > prev:
            out[q] = out[q - 1] + 1
            prev = b
    return k2_max, out, centers_new[good], n_modes[good]


def powerspec(components, gridsize, interpolation=None, deconvolve=None, interlace=None, k_max=None,
              bins_per_decade=None, gridsizes_upstream=None):
    """compute_powerspec (analysis.py:500-579) of a group of particle components on the GPU:
    PCS (default) deposit of ρ onto one or two interlaced lattices, forward FFT, Nyquist nullification,
    deconvolution and interlacing phase (interpolate_upstream(…, output_space='Fourier'), mesh.py:492-616),
    |δ̂|² summed per integer k² (pm_power_k2) and binned like get_powerspec_bins.
    Returns (k_bin_centers, power, n_modes)."""
    import torch
    from . import communication, mesh
    opt = POWERSPEC_DEFAULTS
    interpolation = opt['interpolation'] if interpolation is None else interpolation
    order = {'NGP': 1, 'CIC': 2, 'TSC': 3, 'PCS': 4}.get(str(interpolation).upper(), interpolation)
    deconvolve = opt['deconvolve'] if deconvolve is None else deconvolve
    interlace = opt['interlace'] if interlace is None else interlace
    k_max = opt['k_max'] if k_max is None else k_max
    ctx = mesh.get_context(gridsize, 'f64')
    k2_max, _ = get_powerspec_bins(gridsize, k_max, bins_per_decade)
    # interlace2latticekind (commons.py): True → 'bcc' (two lattices), False → 'sc'
    shifts = [None, (-0.5, -0.5, -0.5)] if interlace in (True, 'bcc') else [None]
    nl = len(shifts)
    fft_factor = float(gridsize)**(-3)
    if gridsizes_upstream is not None and any(int(g) != int(gridsize) for g in gridsizes_upstream):
        _upstream_to_global_mixed(components, [int(g) for g in gridsizes_upstream], int(gridsize), ctx, order,
                                  int(bool(deconvolve))*order, shifts)
        shifts = []
    for l, shift in enumerate(shifts):
        ctx.grid_zero()
        for component in components:
            mesh.interpolate_particles(component, gridsize, ctx, 'ρ', order, None, shift, fft_factor)
        ctx.halo_add()
        ctx.fft_forward()
        ctx.fourier_operate(deconv_order=int(bool(deconvolve))*order, shift=shift, scale=1.0/nl)
        if nl > 1:
            ctx.slab_save() if l == 0 else ctx.slab_accumulate()
    if nl > 1 and shifts:
        ctx.slab_restore()
    dev = ctx.torch_device
    power_k2 = torch.zeros(k2_max + 1, dtype=torch.float64, device=dev)
    count_k2 = torch.zeros(k2_max + 1, dtype=torch.int64, device=dev)
    ctx.power_k2(k2_max, power_k2, count_k2)
    if communication.nprocs > 1:
        ctx.allreduce_sum(power_k2)
        cnt = count_k2.to(torch.float64)
        ctx.allreduce_sum(cnt)
        count_k2 = cnt.round().to(torch.int64)
    ctx.grid_zero()     # leave the context in real space for the next kick
    power_k2 = power_k2.cpu().numpy()
    _, k_bin_indices, k_bin_centers, n_modes = get_powerspec_bins(gridsize, k_max, bins_per_decade, count_k2.cpu().numpy())
    power = np.zeros(len(k_bin_centers))
    np.add.at(power, k_bin_indices, power_k2)
    a = commons.universals.a
    normalization = sum(a**(-3*(1 + c.w_eff(a=a)))*c.ϱ_bar for c in components)**(-2)*float(commons.params.boxsize)**3
    power *= normalization/n_modes
    return k_bin_centers, power, n_modes


def _upstream_to_global_mixed(components, gridsizes_upstream, gridsize, ctx_global, order, deconv_order, shifts):
    """interpolate_upstream (mesh.py:492-616) for components with their own upstream grid sizes: every group is
    deposited on its own grid, transformed, and copied — with its deconvolution, interlacing phase and the half-cell
    phase between grids — into the global slab (add_upstream_to_global_slabs :618-710, copy_modes :980-1322), which
    ends up in ctx_global's working slab.  On several ranks pm_fourier_copy_modes exchanges the mode rows between them."""
    from . import mesh
    nl = len(shifts)
    first = True
    for gridsize_upstream in sorted(set(gridsizes_upstream), key=lambda g: (g != gridsize, g)):
        ctx = ctx_global if gridsize_upstream == gridsize else mesh.get_context(gridsize_upstream, 'f64')
        group = [c for c, g in zip(components, gridsizes_upstream) if g == gridsize_upstream]
        for shift in shifts:
            ctx.grid_zero()
            for component in group:
                mesh.interpolate_particles(component, gridsize_upstream, ctx, 'ρ', order, None, shift, float(gridsize_upstream)**(-3))
            ctx.halo_add()
            ctx.fft_forward()
            if ctx is ctx_global:
                ctx.fourier_operate(deconv_order=deconv_order, shift=shift, scale=1.0/nl)
                ctx.slab_save() if first else ctx.slab_accumulate()
            else:
                ctx.fourier_copy_modes_into(ctx_global, deconv_order, shift, 1.0/nl, src_saved=False, dst_saved=True,
                                            accumulate=not first)
            first = False
    ctx_global.fourier_operate(from_saved=True)          # working slab = accumulated global slab


def get_linear_powerspec(component_or_components, k_magnitudes, a=-1):
    """linear.get_linear_powerspec (linear.py:3074-3133): P_lin(k) = (ζ(k)·T_δ(a, k))² of the (combined) species, with the
    δ transfer function from ic.compute_transfer — CLASS in the reference, concept_b200.linear (or installed tables)
    here.  For several components the transfer functions are averaged with the weights ϱ̄ (one combined species)."""
    from . import ic
    components = component_or_components if isinstance(component_or_components, (list, tuple)) else [component_or_components]
    if a == -1:
        a = commons.universals.a
    k_magnitudes = np.asarray(k_magnitudes, dtype=np.float64)
    δ = np.zeros_like(k_magnitudes)
    weight_total = 0.0
    for component in components:
        spline, _ = ic.compute_transfer(component, 0, -1, a=a)
        values = spline.eval_array(k_magnitudes) if hasattr(spline, 'eval_array') else np.array([spline.eval(k) for k in k_magnitudes])
        δ += component.ϱ_bar*values
        weight_total += component.ϱ_bar
    δ /= weight_total
    return (ic.get_primordial_curvature_perturbation(k_magnitudes)*δ)**2

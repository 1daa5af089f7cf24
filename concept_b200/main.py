"""Time loop of the host mirror: KDK leapfrog with long-range kicks half a step out of phase with
the drifts (reference main.py:102-471), its time-step controller and the kick/drift drivers.

Reference: timeloop :102, get_base_timestep_size :697-907, update_base_timestep_size :913-975,
get_time_step_integrals :998-1073, kick_long :1104-1144, kick_short :1173, driftkick_short :1347-1624,
constants :2325-2433.  Host-side scheduling stays Python; the only work done
per step outside libpmgrav.so is O(1) scalar arithmetic.

Out of scope here: fluids, autosave, component activation/termination, renders.
"""
import itertools
import math
import os

import numpy as np

from . import analysis, commons, communication, integration, interactions
from .commons import abort, machine_ϵ, masterprint, masterwarn, universals, ထ
from .integration import cosmic_time, hubble, init_time, scale_factor, scalefactor_integral

# main.py:2325-2433
Δt_initial_fac = 0.95
Δt_reduce_fac = 0.94
Δt_increase_fac = 0.96
Δt_increase_min_factor = 1.01
Δt_ratio_warn = 0.7
Δt_ratio_abort = 0.01
Δt_jump_fac = 0.95
Δt_reltol = 1e-9
Δt_period = 1*8
bottleneck_static_timestepping = 'static time-stepping'
initial_fac_times = set()
ᔑdt_scalar = {}


def _facs():
    p = commons.params
    return dict(
        dynamical=0.056*p.Δt_base_background_factor, hubble=0.031*p.Δt_base_background_factor,
        ẇ=0.0017*p.Δt_base_background_factor, Γ=0.0028*p.Δt_base_background_factor,
        pm=0.13*p.Δt_base_nonlinear_factor, p3m=0.14*p.Δt_base_nonlinear_factor,
        softening=0.025*p.Δt_rung_factor,
    )


def get_time_step_integrals(t_start, t_end, components):
    """main.py:998-1073: every ᔑdt[...] = ∫ integrand(a(t)) dt over [t_start, t_end]."""
    if not ᔑdt_scalar or ᔑdt_scalar.get('__names') != tuple(c.name for c in components):
        ᔑdt_scalar.clear()
        ᔑdt_scalar['__names'] = tuple(c.name for c in components)
        for integrand in ('1', 'a**2', 'a**(-1)', 'a**(-2)', 'ȧ/a'):
            ᔑdt_scalar[integrand] = 0
        for c in components:
            for integrand in ('a**(-3*w_eff)', 'a**(-3*(1+w_eff))', 'a**(-3*w_eff-1)', 'a**(3*w_eff-2)', 'a**(-3*w_eff)*Γ/H'):
                ᔑdt_scalar[integrand, c.name] = 0
        for c0, c1 in itertools.product(components, components):
            ᔑdt_scalar['a**(-3*w_eff₀-3*w_eff₁-1)', c0.name, c1.name] = 0
    for integrand in list(ᔑdt_scalar):
        if integrand == '__names':
            continue
        ᔑdt_scalar[integrand] = scalefactor_integral(integrand, t_start, t_end, components)
    return ᔑdt_scalar


def get_base_timestep_size(components, static_timestepping_func=None):
    """main.py:697-907 for particle components (no fluids, no decay, constant w)."""
    p = commons.params
    fac = _facs()
    t, a = universals.t, universals.a
    if static_timestepping_func is not None:
        return static_timestepping_func(a), bottleneck_static_timestepping
    H = hubble(a)
    Δt_max, bottleneck = ထ, ''
    ρ_bar = sum(a**(-3*(1 + c.w_eff(a=a)))*c.ϱ_bar for c in components)
    Δt_dynamical = fac['dynamical']/(math.sqrt(commons.G_Newton*ρ_bar) + machine_ϵ)
    if Δt_dynamical < Δt_max:
        Δt_max, bottleneck = Δt_dynamical, 'the dynamical time scale'
    if p.enable_Hubble:
        a_next = a + p.Δa_max_late
        if a_next < 1:
            Δt_Δa_late = p.Δt_base_background_factor*(cosmic_time(a_next) - t)
            if Δt_Δa_late < Δt_max:
                Δt_max, bottleneck = Δt_Δa_late, 'the maximum allowed Δa (late)'
        Δt_hubble, bottleneck_hubble = fac['hubble']/H, 'the Hubble time'
        if p.Δa_max_early > 0:
            a_next = a + p.Δa_max_early
            if a_next < 1:
                Δt_Δa_early = p.Δt_base_background_factor*(cosmic_time(a_next) - t)
                if Δt_Δa_early > Δt_hubble:
                    Δt_hubble, bottleneck_hubble = Δt_Δa_early, 'the maximum allowed Δa (early)'
        if Δt_hubble < Δt_max:
            Δt_max, bottleneck = Δt_hubble, bottleneck_hubble
    measurements = {}

    def v_rms_of(c):
        if c not in measurements:
            measurements[c] = analysis.measure(c, 'v_rms')
        return max(measurements[c], machine_ϵ)
    for c in components:     # PM limiter
        if c.forces.get('gravity') != 'pm':
            continue
        resolution = max(c.potential_gridsizes['gravity']['pm'])
        Δt_pm = fac['pm']*(p.boxsize/resolution)/v_rms_of(c)
        if Δt_pm < Δt_max:
            Δt_max, bottleneck = Δt_pm, f'the PM method of the gravity force for {c.name}'
    for c in components:     # P³M limiter
        if c.forces.get('gravity') != 'p3m':
            continue
        scale = commons.shortrange_scale(c.potential_gridsizes['gravity']['p3m'][0])
        Δt_p3m = fac['p3m']*scale/v_rms_of(c)
        if Δt_p3m < Δt_max:
            Δt_max, bottleneck = Δt_p3m, f'the P³M method of the gravity force for {c.name}'
    if t in initial_fac_times:
        Δt_max *= Δt_initial_fac
    # Record static time-stepping to disk (main.py:897-912)
    if communication.master and isinstance(p.static_timestepping, str):
        if t + Δt_max < cosmic_time(1):
            Δa_max = scale_factor(t + Δt_max) - a
            n = int(math.ceil(math.log10(1/Δt_reltol) + 0.5))
            with open(p.static_timestepping, mode='a', encoding='utf-8') as f:
                if f.tell() == 0:
                    header = ['Time-stepping recorded by concept_b200', '', '{}a{}Δa'.format(' '*((n + 3)//2), ' '*(n + 5))]
                    f.write('\n'.join(f'# {line}' for line in header) + '\n')
                f.write(f'{{:.{n}e}} {{:.{n}e}}\n'.format(a, Δa_max))
    return Δt_max, bottleneck


def _remove_doppelgängers(x, y, rel_tol):
    """integration.py:403-500: drop consecutive (nearly) equal x, scanning from the right"""
    size = len(x)
    if size < 2:
        return x.copy(), y.copy()
    if np.any(np.diff(x) < 0):
        abort('The values in the x array passed to remove_doppelgängers() are not in increasing order')
    accepted = [size - 1, size - 2]
    x_prev = x[size - 2]
    xdiff_prev = x[size - 1] - x_prev
    for i in range(size - 3, -1, -1):
        xdiff = x_prev - x[i]
        if xdiff > rel_tol*xdiff_prev:
            accepted.append(i)
            x_prev, xdiff_prev = x[i], xdiff
    accepted = accepted[::-1]
    return x[accepted].copy(), y[accepted].copy()


class StaticTimestepping:
    """`static_timestepping` of the reference (main.py:499-656) as a look-up object: a ↦ Δt.

    Built from a callable a ↦ Δa or from a recorded file of (a, Δa) rows.  A recorded scale factor is answered with
    the recorded Δa (rows of the same `a` are handed out in file order); any other `a` is interpolated
    log–log within its *stretch* — the table is cut wherever Δa drops (a synchronisation in the recorded run), so an
    interpolation never runs across such a drop.  Every rank evaluates the same function of the same numbers."""

    def __init__(self, func=None):
        self.func = func            # a -> Δa, or None when table-driven
        self.recorded = {}          # rounded a -> recorded Δa values, next one last
        self.lefts = []             # left edge of every stretch (first one 0)
        self.stretches = []         # per stretch: (log a, log Δa) arrays
        self.digits = int(math.ceil(math.log10(1/Δt_reltol) + 0.5))

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_file(cls, path):
        self = cls()
        a_rows, Δa_rows = (np.array(col, dtype=float) for col in np.loadtxt(path, unpack=True, ndmin=2))
        for a, Δa in zip(a_rows[::-1], Δa_rows[::-1]):           # reversed: list.pop() then follows the file order
            self.recorded.setdefault(float(a), []).append(float(Δa))
        a_tab, Δa_tab = _remove_doppelgängers(a_rows, Δa_rows, Δt_reltol)
        # a stretch ends where Δa drops; of two consecutive drops only the first counts, and never the last row
        keep = []           # (row index of the drop, accepted)
        for d in np.flatnonzero(np.diff(Δa_tab) < 0):
            d = int(d)
            suppressed = bool(keep) and keep[-1][1] and d == keep[-1][0] + 1      # directly after an accepted drop
            keep.append((d, not suppressed))
        starts = [d + 1 for d, accepted in keep if accepted and d != len(Δa_tab) - 2]
        bounds = [0] + starts + [len(a_tab)]
        self.lefts = [0.0] + [float(a_tab[i]) for i in starts]
        self.stretches = [(np.log(a_tab[lo:hi]), np.log(Δa_tab[lo:hi])) for lo, hi in zip(bounds[:-1], bounds[1:])]
        return self

    # -- evaluation --------------------------------------------------------------------------
    def _stretch_of(self, a):
        """Index of the stretch holding a, and a snapped onto a stretch edge it is indistinguishable from."""
        import bisect
        i = bisect.bisect_right(self.lefts, a) - 1
        if i + 1 < len(self.lefts) and math.isclose(float(a), self.lefts[i + 1]):
            i += 1                                  # on the right edge: that point opens the next stretch
        if i < 0:
            abort(f'static time-stepping: a = {a} lies before the recorded stretches')
        if math.isclose(float(a), float(self.lefts[i] + machine_ϵ)):
            a = self.lefts[i]
        return i, a

    @staticmethod
    def _loglog(xs, ys, a):
        if len(xs) == 1:
            return math.exp(float(ys[0]))
        x = math.log(a)
        hi = min(max(int(np.searchsorted(xs, x, side='left')), 1), len(xs) - 1)      # linear, extrapolating at both ends
        lo = hi - 1
        slope = (ys[hi] - ys[lo])/(xs[hi] - xs[lo])
        return math.exp(float(slope*(x - xs[lo]) + ys[lo]))

    def Δa(self, a):
        if self.func is not None:
            return self.func(a)
        pending = self.recorded.get(float(f'{{:.{self.digits}e}}'.format(a)))
        if pending:
            return pending.pop()
        i, a = self._stretch_of(a)
        return self._loglog(*self.stretches[i], a)

    def __call__(self, a=-1):
        """Δt of the base step starting at scale factor a (default: now); ∞ past a = 1."""
        if a == -1:
            a, t = universals.a, universals.t
        else:
            t = cosmic_time(a)
        a_next = a + self.Δa(a)
        return cosmic_time(a_next) - t if a_next <= 1 else ထ


def prepare_static_timestepping():
    """main.py:499-656.  `static_timestepping` is None, a callable a ↦ Δa, or a path: an existing file of (a, Δa)
    records is replayed, otherwise the run records its own time-stepping to that path (see get_base_timestep_size)."""
    spec = commons.params.static_timestepping
    if spec is None:
        return None
    if callable(spec):
        masterprint('Static time-stepping configured using supplied function')
        return StaticTimestepping(spec)
    if not isinstance(spec, str):
        abort(f'Could not interpret static_timestepping = {spec} of type {type(spec)}')
    if os.path.isdir(spec):
        abort(f'Supplied static_timestepping = "{spec}" is a directory, not a file')
    if os.path.exists(spec):
        masterprint(f'Static time-stepping information will be read from "{spec}"')
        return StaticTimestepping.from_file(spec)
    directory = os.path.dirname(spec)
    if directory and communication.master:
        os.makedirs(directory, exist_ok=True)
    masterprint(f'Static time-stepping information will be written to "{spec}"')
    return None


def update_base_timestep_size(Δt, Δt_min, Δt_max, bottleneck, time_step=-1, time_step_last_sync=-1,
                              *, allow_increase=True, tolerate_danger=False):
    """main.py:913-975"""
    p = commons.params
    if Δt > Δt_max:
        Δt_new = Δt_reduce_fac*Δt_max
        Δt_ratio = Δt_new/Δt
        if Δt_ratio < Δt_ratio_abort and not tolerate_danger:
            abort(f'Due to {bottleneck}, the time step size needs to be rescaled by a factor {Δt_ratio:.1g}.')
        elif Δt_ratio < Δt_ratio_warn:
            masterwarn(f'Rescaling time step size by a factor {Δt_ratio:.1g} due to {bottleneck}')
        if Δt_new < Δt_min:
            abort(f'Time evolution effectively halted with a time step size of {Δt_new} {commons.unit_time}')
        return Δt_new, bottleneck
    if not allow_increase:
        return Δt, bottleneck
    Δt_new = Δt_increase_fac*Δt_max
    if Δt_new < Δt:
        Δt_new = Δt
    period_frac = (time_step + 1 - time_step_last_sync)*(1/Δt_period)
    period_frac = min(1, max(0, period_frac))
    Δt_tmp = (1 + period_frac*(p.Δt_increase_max_factor - 1))*Δt
    if Δt_new > Δt_tmp:
        Δt_new = Δt_tmp
    if p.enable_Hubble and universals.t + Δt_new > cosmic_time(1):
        return Δt, 'a ≈ 1'
    return Δt_new, ''


def kick_long(components, Δt, sync_time, step_type):
    """main.py:1104-1144"""
    t_start = universals.t
    t_end = t_start + (Δt/2 if step_type == 'init' else Δt)
    if t_end + Δt_reltol*Δt + 2*machine_ϵ > sync_time:
        t_end = sync_time
    if t_start == t_end:
        return
    ᔑdt = get_time_step_integrals(t_start, t_end, components)
    printout = True
    for force, method, receivers, suppliers in interactions.find_interactions(components, 'long-range'):
        getattr(interactions, force)(method, receivers, suppliers, ᔑdt, 'long-range', printout)


def kick_short(components, Δt, fake=False):
    """main.py:1173-1345.  Pure-PM runs have no short-range interactions: nothing to do."""
    if not interactions.find_interactions(components, 'short-range'):
        return
    from . import shortrange
    shortrange.kick_short(components, Δt, fake)


def driftkick_short(components, Δt, sync_time):
    """main.py:1347-1624.  Without short-range interactions: one drift over the whole base step."""
    particle_components = [c for c in components if c.representation == 'particles']
    if not particle_components:
        return
    if not interactions.find_interactions(particle_components, 'short-range'):
        t_start = universals.t
        t_end = t_start + Δt
        if t_end + Δt_reltol*Δt + 2*machine_ϵ > sync_time:
            t_end = sync_time
        if t_start == t_end:
            return
        ᔑdt = get_time_step_integrals(t_start, t_end, particle_components)
        for component in particle_components:
            component.drift(ᔑdt)
        return
    from . import shortrange
    shortrange.driftkick_short(components, Δt, sync_time)


def _uses_rungs(component):
    return component.forces.get('gravity') == 'p3m' and commons.params.N_rungs > 1


def initialize_rung_populations(components, Δt):
    """main.py:1639-1675: all particles on rung 0, then a fake short kick assigns the rungs."""
    if not any(_uses_rungs(c) for c in components):
        return
    if Δt == 0:
        abort('Cannot initialise rung populations with Δt = 0')
    from . import shortrange
    for c in components:
        if _uses_rungs(c):
            shortrange.ensure_rung_state(c)
            c.rung_indices.zero_()
            c.rung_indices_jumped.zero_()
            shortrange._set_populations(c, [c.N_local] + [0]*(commons.params.N_rungs - 1))
    kick_short(components, Δt, fake=True)


def _assign_rungs(components, Δt):
    from . import shortrange
    for c in components:
        if _uses_rungs(c):
            shortrange.ensure_rung_state(c)
            shortrange.assign_rungs(c, Δt, _facs()['softening'])


class DumpTime:
    def __init__(self, a=None, t=None):
        if a is not None:
            self.time_param, self.a, self.t = 'a', float(a), cosmic_time(float(a))
        else:
            self.time_param, self.t, self.a = 't', float(t), scale_factor(float(t))


def _output_times_flat():
    """output_times as {time_param: {kind: [values]}} (commons.py:2800-2870): the parameter is either
    {kind: values} (scale factors) or {'a': {kind: values}, 't': {kind: values}}; None means no output."""
    ot = commons.params.output_times or {}
    nested = {'a': {}, 't': {}}
    for key, val in ot.items():
        if key in ('a', 't') and isinstance(val, dict):
            for kind, v in val.items():
                nested[key][kind] = v
        else:
            nested['a'][key] = val
    flat = {'a': {}, 't': {}}
    for time_param, kinds in nested.items():
        for kind, v in kinds.items():
            values = [float(x) for x in np.ravel(v).tolist() if x is not None] if v is not None else []
            if values:
                flat[time_param][kind] = values
    return flat


def _dump_times():
    flat = _output_times_flat()
    times = {}
    for a in {x for values in flat['a'].values() for x in values}:
        if a >= universals.a:
            d = DumpTime(a=a)
            times[d.t] = d
    for t in {x for values in flat['t'].values() for x in values}:
        if t >= universals.t:
            times.setdefault(t, DumpTime(t=t))
    return [times[t] for t in sorted(times)]


def _output_dir(kind):
    od = commons.params.output_dirs
    if isinstance(od, str):      # one directory for everything (param/example_basic)
        return od
    return od.get(kind, od.get('default')) if isinstance(od, dict) else None


def _output_filename(kind, dump_time):
    """`{output_dirs[kind]}/{output_bases[kind]}_{a|t}=…` with as few decimals as keep all output times of this kind (and
    the start of the run) apart and the first of them away from zero — the naming of prepare_for_output (main.py:2236-2278);
    cosmic-time names carry the time unit (main.py:1698-1699).  None without an output directory."""
    out_dir = _output_dir(kind)
    if not out_dir:
        return None
    bases = commons.user_params.get('output_bases', {})
    base = bases.get(kind, kind) if isinstance(bases, dict) else kind
    tp = dump_time.time_param
    begin = commons.params.a_begin if tp == 'a' else commons.params.t_begin
    times = sorted(set([float(begin)] + _output_times_flat()[tp].get(kind, [])))
    ndigits = 0
    while ndigits < 15:
        names = [f'{x:.{ndigits}f}' for x in times]
        if len(set(names)) == len(names) and (not times[0] or names[0] != f'{0:.{ndigits}f}'):
            break
        ndigits += 1
    value = dump_time.a if tp == 'a' else dump_time.t
    name = f'{base}_' if base else ''
    return os.path.join(out_dir, f'{name}{tp}={value:.{ndigits}f}' + (commons.unit_time if tp == 't' else ''))


_warned_outputs = set()


def _wanted(kind, dump_time):
    flat = _output_times_flat()
    return (any(x == dump_time.a for x in flat['a'].get(kind, ())) and dump_time.time_param == 'a') or \
           (any(x == dump_time.t for x in flat['t'].get(kind, ())) and dump_time.time_param == 't')


def dump_powerspec(components, dump_time, filename=None):
    """analysis.powerspec + save_powerspec (analysis.py:500-579, :796-836) with the defaults of powerspec_options
    (commons.py:3354-3385: PCS, deconvolution, bcc interlacing, k_max = Nyquist, grid 2·∛N).  Columns as in the
    reference's text files: k, number of modes, P(k), linear P(k) (powerspec_select's default, commons.py:2636-2643).
    Returns (k, power, n_modes)."""
    out_dir = os.path.dirname(filename) or '.' if filename else _output_dir('powerspec')
    particle_components = [c for c in components if c.representation == 'particles']
    if not particle_components:
        return None
    # powerspec_options (commons.py:3354-3385): upstream grid size per component (default '2*cbrt(N)'), global grid size
    # (default: the largest upstream size), interpolation, deconvolution, interlacing, k_max and the bins per decade — each
    # looked up by component name, species, 'particles', 'all', 'default'
    options = commons.user_params.get('powerspec_options', {})
    options = {str(k).lower().replace('_', ' '): v for k, v in options.items()} if isinstance(options, dict) else {}
    if 'gridsize' in options:       # one grid size for both (commons.py:3392-3398)
        options.setdefault('upstream gridsize', options['gridsize'])
        options.setdefault('global gridsize', options['gridsize'])

    def option(name, component, default):
        spec = options.get(name, default)
        if isinstance(spec, dict) and name != 'bins per decade' or \
                (name == 'bins per decade' and isinstance(spec, dict) and any(isinstance(v, dict) for v in spec.values())):
            lowered = {str(k).lower(): v for k, v in spec.items()}
            for key in (component.name, component.species, 'particles', 'all', 'default'):
                if str(key).lower() in lowered:
                    return lowered[str(key).lower()]
            return default
        return spec
    first = particle_components[0]
    gridsizes_upstream = [commons.gridsize_value(option('upstream gridsize', c, '2*cbrt(N)'), c.N) for c in particle_components]
    gridsizes_upstream = [g + (g & 1) for g in gridsizes_upstream]
    gridsize = option('global gridsize', first, -1)
    gridsize = max(gridsizes_upstream) if gridsize in (-1, None) else commons.gridsize_value(gridsize, first.N)
    k_max = option('k max', first, option('k_max', first, None))
    bins_per_decade = option('bins per decade', first, None)
    if bins_per_decade is not None and not isinstance(bins_per_decade, dict):
        bins_per_decade = {'k_min': float(bins_per_decade)}        # the same number of bins in every decade
    k, power, n_modes = analysis.powerspec(
        particle_components, gridsize, interpolation=option('interpolation', first, None), deconvolve=option('deconvolve', first, None),
        interlace=option('interlace', first, None), k_max=k_max.lower() if isinstance(k_max, str) else k_max,
        bins_per_decade=bins_per_decade, gridsizes_upstream=gridsizes_upstream)
    if out_dir and communication.master:
        os.makedirs(out_dir, exist_ok=True)
        filename = filename or _output_filename('powerspec', dump_time)
        names = ', '.join(c.name for c in particle_components)
        # powerspec_select (commons.py:2636-2643): the linear-theory column unless it is switched off for these components
        select = commons.user_params.get('powerspec_select', {})
        want_linear = True
        if isinstance(select, dict):
            lowered = {str(key).lower(): val for key, val in select.items()}
            for key in (first.name, first.species, 'particles', 'all', 'default'):
                if str(key).lower() in lowered:
                    val = lowered[str(key).lower()]
                    want_linear = bool(val.get('linear', False)) if isinstance(val, dict) else bool(val)
                    break
        columns, fmt = [k, n_modes, power], ['%.8e', '%d', '%.8e']
        header = (f'Power spectrum of {names} at a = {universals.a:.8g}, t = {universals.t:.8g} {commons.unit_time}, '
                  f'grid size {gridsize} (concept_b200)\n'
                  f'k [{commons.unit_length}^-1]\tmodes\tpower [{commons.unit_length}^3]')
        if want_linear:
            columns.append(analysis.get_linear_powerspec(particle_components, k))
            fmt.append('%.8e')
            header += f'\tlinear power [{commons.unit_length}^3]'
        np.savetxt(filename, np.column_stack(columns), fmt=tuple(fmt), delimiter='\t', header=header)
    return k, power, n_modes


def dump(components, dump_time, on_dump=None):
    """main.py:1676-1712: power spectra as text files; snapshots as GADGET-2 files when snapshot_type = 'gadget'
    (snapshot.save), else — the reference's own 'concept' format is HDF5, which this image cannot write — as .npz with a
    warning; bispectra and renders are out of scope and skipped with a warning.  `on_dump` is the hook the tests use to
    capture the state."""
    from . import snapshot
    if on_dump is not None:
        on_dump(components, dump_time)
    if _wanted('powerspec', dump_time):
        dump_powerspec(components, dump_time)
    for kind in ('bispec', 'render2D', 'render3D'):
        if _wanted(kind, dump_time) and kind not in _warned_outputs:
            _warned_outputs.add(kind)
            commons.masterwarn(f'Output of kind "{kind}" is not provided by concept_b200 (SURVEY.md §2): skipped')
    filename = _output_filename('snapshot', dump_time)
    if filename and _wanted('snapshot', dump_time):
        snapshot_type = str(commons.user_params.get('snapshot_type', 'concept')).lower()
        if snapshot_type != 'gadget' and 'snapshot_type' not in _warned_outputs:
            _warned_outputs.add('snapshot_type')
            commons.masterwarn(f'snapshot_type = "{snapshot_type}" needs HDF5, which is not available here: '
                               'the particle data are written as NumPy .npz files (use snapshot_type = "gadget" for GADGET-2 files)')
        for c in components:
            suffix = f'_{c.name}' if len(components) > 1 else ''
            if communication.master:
                os.makedirs(os.path.dirname(filename) or '.', exist_ok=True)
            if snapshot_type == 'gadget':
                snapshot.save(c, filename + suffix)
            else:
                pos, mom = c.gather_global()
                if communication.master:
                    np.savez(filename + suffix + '.npz', pos=pos, mom=mom, mass=c.mass, a=universals.a, t=universals.t,
                             boxsize=commons.params.boxsize)
    return False


class _Leapfrog:
    """The stepping state of one run (reference main.py:102-471 for particle components).

    The integrator is kick-drift-kick with the long-range kicks half a step out of phase with the drifts.  The system
    is *synchronised* (positions and momenta at the same time) at the start, at every dump and whenever the base step
    Δt has to change; a synchronised state is left with half kicks (`_leave_synchronised_state`), every other step is a
    drift over Δt followed by a long-range kick over Δt that is clipped at `sync_time` when a synchronisation is due
    (`_drift_and_kick`).  `_plan` decides after each step whether the next one has to end synchronised."""

    def __init__(self, components, static_timestepping_func, on_dump, on_step, max_steps):
        self.components = components
        self.static = static_timestepping_func
        self.on_dump, self.on_step, self.max_steps = on_dump, on_step, max_steps
        self.time_step = 0
        self.announced_step = -1
        self.last_sync_step = 0
        self.synchronised = True
        self.sync_time = ထ
        self.limit_is_fresh = False        # Δt_max was measured at the current time already
        self.Δt_shelved = -1               # the step size put aside while a dump forces a shorter one
        self.last_sort_step = 0
        initial_fac_times.add(universals.t)
        self.Δt_max, self.bottleneck = get_base_timestep_size(components, self.static)
        self.Δt = self.Δt_max

    # -- pieces of a step --------------------------------------------------------------------
    def _announce_step(self):
        if self.time_step > self.announced_step:
            self.announced_step = self.time_step
            if self.synchronised:           # all rungs are synchronised: re-assign them (main.py:225-227)
                _assign_rungs(self.components, self.Δt)
                self._restore_memory_order()
            universals.time_step = self.time_step

    def _restore_memory_order(self):
        """Deposit and gather rely on consecutive particles touching neighbouring grid rows (a randomly ordered array costs
        3.6× the cycle time, bench.py `particle_order`).  Lattice-ordered initial conditions have that order and the hole-filling
        migration keeps it, but it decays as structure forms; like the reference's tile sort at synchronised steps
        (main.py:270-305) the particles are re-ordered by grid cell there, every `cell_sort_period` base steps (a sort costs
        about 0.6 PM cycles)."""
        period = commons.params.cell_sort_period
        if period <= 0 or self.time_step - self.last_sort_step < period:
            return
        self.last_sort_step = self.time_step
        for c in self.components:
            if c.representation == 'particles' and c.N_local > 1:
                c.cell_sort()

    def _measure_limit(self):
        self.Δt_max, self.bottleneck = get_base_timestep_size(self.components, self.static)

    def _leave_synchronised_state(self):
        self.synchronised = False
        kick_long(self.components, self.Δt, self.sync_time, 'init')
        kick_short(self.components, self.Δt)

    def _advance_clock(self, Δt_half):
        universals.t += Δt_half
        if universals.t + Δt_reltol*self.Δt + 2*machine_ϵ > self.sync_time:
            universals.t = self.sync_time
        universals.a = scale_factor(universals.t)

    def _drift_and_kick(self):
        driftkick_short(self.components, self.Δt, self.sync_time)
        self._advance_clock(0.5*self.Δt)
        kick_long(self.components, self.Δt, self.sync_time, 'full')
        self._advance_clock(0.5*self.Δt)
        if self.on_step is not None:
            self.on_step(self.time_step, universals.t, universals.a, self.Δt)

    def _plan(self, dump_time, after_half_kicks):
        """Schedule a synchronisation if the dump is near, if Δt exceeds its limit, or (between synchronisations only)
        if Δt may grow again after a full period."""
        Δt = self.Δt
        if dump_time.t - universals.t <= 1.5*Δt:
            self.sync_time = dump_time.t
            return
        self._measure_limit()
        too_large = Δt > self.Δt_max
        may_grow = (not after_half_kicks and self.Δt_max > Δt_increase_min_factor*Δt
                    and (self.time_step + 1 - self.last_sync_step) >= Δt_period)
        if too_large or may_grow:
            self.sync_time = universals.t + (0.5*Δt if after_half_kicks else Δt)
            self.limit_is_fresh = True

    def _cap_for_dump(self, t_dump):
        room = t_dump - universals.t
        if self.Δt > room:
            self.Δt_shelved = self.Δt
            self.Δt = room

    def _resynchronised(self, dump_time, next_dump_time):
        """The step just taken ended at sync_time.  Returns True when that was the dump."""
        self.synchronised = True
        self.sync_time = ထ
        if self.Δt_shelved != -1:
            self.Δt = max(self.Δt, self.Δt_shelved)
            self.Δt_shelved = -1
        if not self.limit_is_fresh:
            self._measure_limit()
        self.limit_is_fresh = False
        self.Δt, self.bottleneck = update_base_timestep_size(
            self.Δt, self.Δt_min, self.Δt_max, self.bottleneck, self.time_step, self.last_sync_step,
            tolerate_danger=(self.bottleneck == bottleneck_static_timestepping))
        self.time_step += 1
        self.last_sync_step = self.time_step
        if universals.t == dump_time.t:
            dump(self.components, dump_time, self.on_dump)
            if next_dump_time is not None:
                self.Δt_max = next_dump_time.t - universals.t
                self._cap_for_dump(next_dump_time.t)
            return True
        self.Δt_max = dump_time.t - universals.t
        self._cap_for_dump(dump_time.t)
        return False

    # -- driver ------------------------------------------------------------------------------
    def run_to(self, dump_time, next_dump_time):
        """Step until `dump_time` has been dumped; False if max_steps ended the run first."""
        while True:
            if self.max_steps is not None and self.time_step >= self.max_steps:
                return False
            self._announce_step()
            if self.synchronised:
                self._leave_synchronised_state()
                self._plan(dump_time, after_half_kicks=True)
                continue
            self._drift_and_kick()
            if universals.t == self.sync_time:
                if self._resynchronised(dump_time, next_dump_time):
                    return True
                continue
            self.time_step += 1
            self._plan(dump_time, after_half_kicks=False)


def timeloop(components, on_dump=None, on_step=None, max_steps=None):
    """main.py:102-471 for particle components.  `components` replaces get_initial_conditions()
    (snapshot loading / IC generation are out of scope).  Returns the number of base steps taken."""
    init_time()
    dump_times = _dump_times()
    if not dump_times or not components:
        return 0
    ᔑdt_scalar.clear()
    initial_fac_times.clear()
    if dump_times[0].t == universals.t or dump_times[0].a == universals.a:
        dump(components, dump_times[0], on_dump)
        dump_times.pop(0)
        if not dump_times:
            return 0
    stepper = _Leapfrog(components, prepare_static_timestepping(), on_dump, on_step, max_steps)
    stepper.Δt = min(stepper.Δt, dump_times[0].t - universals.t)
    stepper.Δt_min = 1e-4*stepper.Δt
    get_time_step_integrals(0, 0, components)
    initialize_rung_populations(components, stepper.Δt)
    for dump_time, next_dump_time in zip(dump_times, dump_times[1:] + [None]):
        if not stepper.run_to(dump_time, next_dump_time):
            break
    return stepper.time_step


def get_initial_conditions(initial_conditions_touse=None, do_realization=True):
    """snapshot.get_initial_conditions (snapshot.py:3425-3474): each entry of the `initial_conditions` parameter
    is a path to a snapshot (GADGET-2 here) or a dict describing a component to be realised (ic.realize_particles)."""
    from . import ic, snapshot
    from .species import Component
    spec = commons.params.initial_conditions if initial_conditions_touse is None else initial_conditions_touse
    if not spec:
        return []
    entries = [spec] if isinstance(spec, (str, dict)) else list(spec)
    components, to_realize = [], []
    for entry in entries:
        if isinstance(entry, str):
            components.append(snapshot.load(entry))
        elif isinstance(entry, dict):
            entry = dict(entry)
            species = str(entry.pop('species', None))
            name = str(entry.pop('name', species))
            to_realize.append(Component(name, species, **{key.replace(' ', '_'): value for key, value in entry.items()}))
        else:
            abort(f'Error parsing initial_conditions of type {type(entry)}')
    if do_realization and to_realize:
        ic.n_particles_realized.update(components_tally=0, components_total=0, particles_tally=0)
        for component in to_realize:
            ic.realize_particles(component, universals.a, components_all=to_realize)
    return components + to_realize


def _load_snapshot_for_utility(path, param, extra):
    """Parameters for a utility run on a snapshot: the parameter file if one is given, else what the snapshot's header says
    about the box and the cosmology (the reference compares the two and warns, utilities.py:470-477)."""
    from . import snapshot
    if param:
        commons.load_params(param, extra)
    else:
        head = snapshot.read_gadget(path)
        commons.load_params(f"boxsize = {head['boxsize']!r}\nH0 = {head['H0']!r}\nΩb = 0.049\nΩcdm = {head['Ωm'] - 0.049!r}\n"
                            f"a_begin = {head['a']!r}\n", extra)
    init_time()
    component = snapshot.load(path)
    universals.a = float(universals.a)
    universals.t = cosmic_time(universals.a) if commons.params.enable_Hubble else universals.t
    return component


def utility_info(path):
    """`concept -u info` (utilities.py, info): what a snapshot holds, one line per fact"""
    from . import snapshot
    d = snapshot.read_gadget(path)
    u = commons.units
    lines = [f'GADGET snapshot "{path}"',
             f'  particles        {len(d["pos"])} (type {d["type"]})',
             f'  particle mass    {d["mass"]:.8g} {commons.unit_mass}',
             f'  a                {d["a"]:.8g}    (z = {1/d["a"] - 1:.6g})',
             f'  boxsize          {d["boxsize"]:.8g} {commons.unit_length}',
             f'  H0               {d["H0"]/(u.km/(u.s*u.Mpc)):.8g} km s⁻¹ Mpc⁻¹',
             f'  Ωm               {d["Ωm"]:.8g}',
             f'  files            {d["header"]["NumFiles"]}']
    if commons.communication_master():
        print('\n'.join(lines))
    return d


def utility_powerspec(path, param='', extra=''):
    """`concept -u powerspec snapshot` (utilities.py:465-497): the power spectrum of a snapshot, written next to it as
    `powerspec_<snapshot name>` (output_bases['powerspec'] + '_' + name)."""
    component = _load_snapshot_for_utility(path, param, extra)
    out_dir, basename = os.path.split(os.path.abspath(path.rstrip('/')))
    bases = commons.user_params.get('output_bases', {})
    base = bases.get('powerspec', 'powerspec') if isinstance(bases, dict) else 'powerspec'
    filename = os.path.join(out_dir, f'{base}_{basename}' if base else basename)
    if filename == os.path.abspath(path):
        filename = os.path.join(out_dir, f'powerspec_{basename}')
    dump_powerspec([component], DumpTime(a=universals.a), filename=filename)
    return filename


def run(param, extra='', on_dump=None, on_step=None, max_steps=None):
    """What `concept -p param` does for a particle-only run (main.py:2437-2470): load the parameter file, set up
    the background, obtain the initial conditions and run the time loop.  Returns the components."""
    commons.load_params(param, extra)
    init_time()
    components = get_initial_conditions()
    timeloop(components, on_dump=on_dump, on_step=on_step, max_steps=max_steps)
    return components

"""Global state shared by the host-side mirror of the reference's operator surface:
unit system, physical constants, the parameter-file loader and `universals`.

Reference: commons.py — units :1826-1905, :2044-2135; parameter file execution :1921-2042
(`exec_params`: statements are executed repeatedly until no further one succeeds, so a line may use
a name that is assigned further down); hot-path parameters :2956-3269, :3664-3702, :3862-3926.

Only the parameters that the PM hot path reads are normalised here.  Unlike the reference, the
parameters are not frozen into module globals at import: `load_params()` may be called again
(one process can run several configurations, which the tests rely on).
"""
import ast
import math
import os
import sys
import types

import numpy as np

π = math.pi
τ = 2*math.pi
machine_ϵ = float(np.finfo(np.float64).eps)
ထ = float('inf')

# ---------------------------------------------------------------------------
# Unit system: everything is expressed in (unit_length, unit_time, unit_mass) =
# (Mpc, Gyr, 10¹⁰ m☉) by default, like the reference.
# ---------------------------------------------------------------------------


def _unit_relations():
    r = {'yr': 1.0, 'pc': 1.0, 'm_sun': 1.0}
    r['kyr'], r['Myr'], r['Gyr'] = 1e3, 1e6, 1e9
    r['day'] = 1/365.25
    r['hr'] = r['day']/24
    r['minutes'] = r['hr']/60
    r['s'] = r['minutes']/60
    r['kpc'], r['Mpc'], r['Gpc'] = 1e3, 1e6, 1e9
    r['AU'] = τ/(60*60*360)
    r['m'] = r['AU']/149597870700
    r['mm'], r['cm'], r['km'] = 1e-3*r['m'], 1e-2*r['m'], 1e3*r['m']
    r['ly'] = (299792458*r['m']/r['s'])*r['yr']
    r['km_sun'], r['Mm_sun'], r['Gm_sun'] = 1e3, 1e6, 1e9
    r['kg'] = 1/1.98841e30
    r['g'] = 1e-3*r['kg']
    r['G_Newton'] = 6.67430e-11*r['m']**3/(r['kg']*r['s']**2)
    return r


class Units(types.SimpleNamespace):
    pass


def build_units(unit_length='Mpc', unit_time='Gyr', unit_mass='10¹⁰ m☉'):
    r = _unit_relations()
    mass_alias = {'10¹⁰ m☉': 1e10, '1e+10*m_sun': 1e10, '10**10*m_sun': 1e10, 'm_sun': 1.0, 'm☉': 1.0}
    yr = 1/r[unit_time]
    pc = 1/r[unit_length]
    m_sun = 1/(mass_alias[unit_mass] if unit_mass in mass_alias else r[unit_mass])
    u = Units()
    for name in ('yr', 'kyr', 'Myr', 'Gyr', 'day', 'hr', 'minutes', 's'):
        setattr(u, name, r[name]*yr)
    for name in ('pc', 'kpc', 'Mpc', 'Gpc', 'AU', 'm', 'mm', 'cm', 'km', 'ly'):
        setattr(u, name, r[name]*pc)
    for name in ('m_sun', 'km_sun', 'Mm_sun', 'Gm_sun', 'kg', 'g'):
        setattr(u, name, r[name]*m_sun)
    return u, r


def set_unit_system(length='Mpc', time='Gyr', mass='10¹⁰ m☉'):
    """The unit system every number of a run is expressed in (unit_length, unit_time, unit_mass of the parameter file,
    commons.py:1935-1999); rebinds `units`, `G_Newton`, `light_speed` and the unit names of this module, which the other
    modules always reach through `commons.`."""
    global units, unit_relations, unit_length, unit_time, unit_mass, light_speed, G_Newton
    r = _unit_relations()
    mass = str(mass).replace(' ', '')
    mass_names = {'10¹⁰m☉': '10¹⁰ m☉', '1e+10*m_sun': '10¹⁰ m☉', '1e10*m_sun': '10¹⁰ m☉', '10**10*m_sun': '10¹⁰ m☉', 'm☉': 'm_sun'}
    mass = mass_names.get(mass, mass)
    if str(length) not in r or str(time) not in r or (mass not in r and mass != '10¹⁰ m☉'):
        abort(f'Unit system ({length}, {time}, {mass}) not understood')
    units, unit_relations = build_units(str(length), str(time), mass)
    unit_length, unit_time, unit_mass = str(length), str(time), ('m☉' if mass == 'm_sun' else mass)
    light_speed = units.ly/units.yr
    G_Newton = (unit_relations['G_Newton']/(unit_relations['m']**3/(unit_relations['kg']*unit_relations['s']**2))
                * units.m**3/(units.kg*units.s**2))


units = unit_relations = unit_length = unit_time = unit_mass = light_speed = G_Newton = None


# ---------------------------------------------------------------------------
# universals (commons.py `universals` struct)
# ---------------------------------------------------------------------------
universals = types.SimpleNamespace(t=0.0, a=1.0, t_begin=0.0, a_begin=1.0, z_begin=0.0, time_step=0)


# ---------------------------------------------------------------------------
# Printing / aborting
# ---------------------------------------------------------------------------
class ConceptAbort(SystemExit):
    """Raised by abort(); the reference prints and terminates all ranks (commons.py:1002-1030)."""


verbose = bool(int(os.environ.get('CONCEPT_B200_VERBOSE', '0')))


def communication_master():
    return int(os.environ.get('RANK', '0')) == 0


def masterprint(*args, **kwargs):
    if verbose and int(os.environ.get('RANK', '0')) == 0:
        print(*args, **kwargs)
        sys.stdout.flush()


def masterwarn(*args, **kwargs):
    if int(os.environ.get('RANK', '0')) == 0:
        print('Warning:', *args, file=sys.stderr, **kwargs)


def abort(*args, exit_code=1):
    msg = ' '.join(str(a) for a in args)
    print('Aborting:', msg, file=sys.stderr)
    raise ConceptAbort(msg or exit_code)


set_unit_system()


# ---------------------------------------------------------------------------
# Parameter file
# ---------------------------------------------------------------------------
class Param:
    """`param` object available inside parameter files (commons.py:1909): .path, .name, .dir"""

    def __init__(self, path=''):
        self.path = os.path.abspath(path) if path else ''
        self.name = os.path.basename(self.path)
        self.dir = os.path.dirname(self.path)

    def __str__(self):
        return self.path

    __fspath__ = __str__


def _param_namespace(param_path):
    ns = {name: getattr(units, name) for name in vars(units)}
    ns.update(
        units=units, light_speed=light_speed, c=light_speed, G_Newton=G_Newton, G=G_Newton,
        π=π, pi=π, τ=τ, tau=τ, ထ=ထ, inf=ထ, machine_ϵ=machine_ϵ, eps=machine_ϵ,
        np=np, numpy=np, asarray=np.asarray, arange=np.arange, linspace=np.linspace, logspace=np.logspace,
        zeros=np.zeros, ones=np.ones, empty=np.empty, array=np.array,
        sqrt=np.sqrt, cbrt=np.cbrt, exp=np.exp, log=np.log, log2=np.log2, log10=np.log10,
        sin=np.sin, cos=np.cos, tan=np.tan, min=min, max=max, sum=np.sum, prod=np.prod, mean=np.mean,
        isint=lambda x: float(x).is_integer(), param=Param(param_path), os=os,
    )
    return ns


def exec_params(content, namespace):
    """Execute parameter-file text statement by statement, retrying failed statements until no
    further one succeeds (commons.py:2001-2036) — a statement may use names assigned later."""
    tree = ast.parse(content)
    pending = [ast.Module(body=[node], type_ignores=[]) for node in tree.body]
    progress = True
    while pending and progress:
        progress = False
        still = []
        for mod in pending:
            try:
                exec(compile(mod, '<param>', 'exec'), namespace)
                progress = True
            except Exception:
                still.append(mod)
        pending = still
    return namespace


_ORDER_NAMES = {'ngp': 1, 'cic': 2, 'tsc': 3, 'pcs': 4}


def _lower_keys(d):
    return {str(k).lower(): v for k, v in d.items()}


def _truthy(val):
    try:
        return any(bool(v) for v in val) if isinstance(val, (list, tuple, set, dict)) else bool(val)
    except (TypeError, ValueError):
        return True      # (arrays and the like)


def fill_ellipses(d):
    """`...` as a dict value means "as the entry before" (the reference's example files use it for output_dirs and the
    *_select dicts; replace_ellipsis, commons.py:2142-2161): an ellipsis takes the nearest earlier value that is set to
    something (non-empty, non-zero) — the first such value of the dict for leading ellipses — and, if there is none at all,
    simply the value before it.  Applied in place to nested dicts as well."""
    if not isinstance(d, dict):
        return d
    for val in d.values():
        fill_ellipses(val)
    keys = list(d)
    if not any(d[k] is ... for k in keys):
        return d
    first_set = next((d[k] for k in keys if d[k] is not ... and _truthy(d[k])), None)
    last_set, last_any = first_set, first_set
    for k in keys:
        if d[k] is ...:
            d[k] = last_set if last_set is not None else last_any
        else:
            last_any = d[k]
            if _truthy(d[k]):
                last_set = d[k]
    return d


class Params(types.SimpleNamespace):
    """Normalised hot-path parameters (subset of commons.py:2470-4432)."""


params = Params()
user_params = {}


def load_params(path_or_text='', extra='', **overrides):
    """Load a CO*N*CEPT parameter file (path or literal text).  `extra` mimics `-c "name = value"`
    (appended lines, doc/command_line_options.rst).  Keyword overrides are applied last."""
    global params, user_params
    if path_or_text and '\n' not in path_or_text and os.path.isfile(path_or_text):
        with open(path_or_text, encoding='utf-8') as f:
            content = f.read()
        path = path_or_text
    else:
        content, path = path_or_text, ''
    content = content + '\n' + extra + '\n'
    # h is defined from H0 after the file (commons.py:1785-1798) — but files use it (`256*Mpc/h`), so
    # it must be resolvable during the retry loop
    content += '\ntry:\n    h = H0/(100*km/(s*Mpc))\nexcept NameError:\n    h = 1\nh = float("{:.15f}".format(h))\n'
    # the unit system first (its names are plain strings in the file); everything else is read in it
    set_unit_system()
    probe = _param_namespace(path)
    exec_params(content, probe)
    set_unit_system(probe.get('unit_length', 'Mpc'), probe.get('unit_time', 'Gyr'), probe.get('unit_mass', '10¹⁰ m☉'))
    ns = _param_namespace(path)
    base_keys = set(ns)
    exec_params(content, ns)
    exec_params(content, ns)   # the reference executes the file twice (commons.py:2352, 2419)
    up = {k: v for k, v in ns.items() if k not in base_keys and not k.startswith('__')}
    up.update(overrides)
    for val in up.values():
        fill_ellipses(val)
    user_params = up
    p = Params()
    p.user = up
    p.boxsize = float(up.get('boxsize', 512*units.Mpc))
    p.H0 = float(up.get('H0', 67*units.km/(units.s*units.Mpc)))
    p.Ωb = float(up.get('Ωb', 0.049))
    p.Ωcdm = float(up.get('Ωcdm', 0.27))
    p.Ωm = p.Ωb + p.Ωcdm
    p.a_begin = float(up.get('a_begin', 1.0))
    p.t_begin = float(up.get('t_begin', 0.0))
    p.enable_Hubble = bool(up.get('enable_Hubble', True))
    p.ρ_crit = 3*p.H0**2/(8*π*G_Newton)            # commons.py:4435
    if str(up.get('softening_kernel', 'spline')).lower() != 'spline':
        abort(f'softening_kernel = "{up["softening_kernel"]}": only the (default) spline kernel is implemented')
    p.initial_conditions = up.get('initial_conditions', None)
    p.output_times = up.get('output_times', {})
    p.output_dirs = up.get('output_dirs', {})
    # initial conditions (commons.py:3640-3658, :3742-3830, :3895-3925)
    p.random_seeds = {'general': 0, 'primordial amplitudes': 1_000, 'primordial phases': 2_000}
    for key, val in dict(up.get('random_seeds', {})).items():
        if key not in p.random_seeds:
            abort(f'Key {key} in random_seeds not understood')
        p.random_seeds[key] = int(val)
    p.random_generator = str(up.get('random_generator', 'PCG64DXSM'))
    p.primordial_noise_imprinting = str(up.get('primordial_noise_imprinting', 'distributed')).lower()
    if p.primordial_noise_imprinting not in {'simple', 'distributed'}:
        abort(f'primordial_noise_imprinting = "{p.primordial_noise_imprinting}" ∉ {{"simple", "distributed"}}')
    p.primordial_amplitude_fixed = bool(up.get('primordial_amplitude_fixed', False))
    p.primordial_phase_shift = float(np.mod(float(up.get('primordial_phase_shift', 0)), τ))
    ps = {}
    for key, val in dict(up.get('primordial_spectrum', {})).items():
        k = str(key).lower()
        for ch in '_sk':
            k = k.replace(ch, '')
        k = k.replace('alpha', 'α')
        full = {'a': 'A_s', 'n': 'n_s', 'α': 'α_s', 'pivot': 'pivot'}.get(k)
        if full is None:
            masterwarn(f'Could not understand primordial spectrum parameter "{key}"')
        else:
            ps[full] = val
    p.primordial_spectrum = {key: float(ps.get(key, val))
                             for key, val in {'A_s': 2.1e-9, 'n_s': 0.96, 'α_s': 0, 'pivot': 0.05/units.Mpc}.items()}
    ro = {}
    for key, val in dict(up.get('realization_options', {})).items():
        k = str(key).lower().replace('-', '').replace('_', '').replace(' ', '')
        if isinstance(val, dict):     # per-component dicts: only the 'default'/'all'/'particles' entry is used here
            val = next((val[q] for q in ('default', 'all', 'particles', 'matter') if q in val), None)
        ro[k] = val
    p.realization_options = {
        'backscale': bool(ro.get('backscale', False) or False),
        'lpt': int(round(ro.get('lpt', 1) or 1)),
        'dealias': bool(ro.get('dealias', False) or False),
        'nongaussianity': float(next((ro[q] for q in ('nongauss', 'nongaussian', 'nongaussianity') if ro.get(q)), 0.0)),
        'structure': 'primordial',
    }
    if p.realization_options['lpt'] not in {1, 2, 3}:
        abort(f"{p.realization_options['lpt']}LPT not implemented")
    p.N_rungs = int(up.get('N_rungs', 8))
    p.Δt_base_background_factor = float(up.get('Δt_base_background_factor', 1))
    p.Δt_base_nonlinear_factor = float(up.get('Δt_base_nonlinear_factor', 1))
    p.Δt_increase_max_factor = float(up.get('Δt_increase_max_factor', ထ))
    p.Δt_rung_factor = float(up.get('Δt_rung_factor', 1))
    p.Δa_max_early = float(up.get('Δa_max_early', 0.00153))
    p.Δa_max_late = float(up.get('Δa_max_late', 0.022))
    p.static_timestepping = up.get('static_timestepping', None)
    # concept_b200 only: base steps between two re-orderings of the particles by grid cell (pm_sort_particles, the analogue of
    # the tile sort the reference performs at its synchronised steps, main.py:270-305); 0 disables it
    # particle_reordering (commons.py, default True): the periodic re-ordering of the particles in memory; here by grid cell
    # every cell_sort_period base steps (0: never)
    p.cell_sort_period = int(up.get('cell_sort_period', 64)) if up.get('particle_reordering', True) else 0
    p.cell_centered = bool(up.get('cell_centered', True))
    if not p.cell_centered:
        abort('cell_centered = False (grid values at the cell vertices) is not implemented: the mesh kernels, the lattice of the '
              'initial conditions and the phases between grids of different size assume cell-centred values (the default)')
    p.grid_dtype = str(up.get('grid_dtype', 'f64'))     # extension: 'f32' selects the mixed-precision grid
    # select_forces (commons.py:3664-3702): default for particles is gravity via P³M
    sf = up.get('select_forces', {})
    p.select_forces = {str(k): _lower_keys(v) if isinstance(v, dict) else {'gravity': str(v).lower()} for k, v in sf.items()}
    # potential_options (commons.py:2958-3237)
    po = up.get('potential_options', {})
    if not isinstance(po, dict):
        po = {'gridsize': po}
    p.potential_options_raw = po
    return p_finish(p)


def _method_dict(spec, default_pm, default_p3m):
    """Accepts a scalar, {'gravity': x}, {'gravity': {'pm': x, 'p3m': y}} → {'pm': x, 'p3m': y}"""
    out = {'pm': default_pm, 'p3m': default_p3m}
    if spec is None:
        return out
    if isinstance(spec, dict):
        spec = _lower_keys(spec)
        if 'default' in spec and isinstance(spec['default'], dict):
            spec = _lower_keys(spec['default'])
        if 'global' in spec or 'particles' in spec:
            merged = {}
            for key in ('global', 'particles'):
                if isinstance(spec.get(key), dict):
                    merged.update(_lower_keys(spec[key]))
            spec = merged or spec
        g = spec.get('gravity', spec if ('pm' in spec or 'p3m' in spec) else None)
        if isinstance(g, dict):
            g = _lower_keys(g)
            for m in ('pm', 'p3m'):
                if m in g:
                    out[m] = g[m]
        elif g is not None:
            out = {'pm': g, 'p3m': g}
    else:
        out = {'pm': spec, 'p3m': spec}
    return out


def p_finish(p):
    global params
    po = _lower_keys(p.potential_options_raw)
    p.gridsize_spec = _method_dict(po.get('gridsize'), None, None)
    interp = _method_dict(po.get('interpolation'), 'CIC', 'CIC')
    p.interpolation_order = {m: (_ORDER_NAMES[str(v).lower()] if isinstance(v, str) else int(v)) for m, v in interp.items()}
    p.deconvolve = {m: (tuple(bool(x) for x in v) if isinstance(v, (tuple, list)) else (bool(v), bool(v)))
                    for m, v in _method_dict(po.get('deconvolve'), (True, True), (True, True)).items()}
    p.interlace = {m: (tuple(bool(x) for x in v) if isinstance(v, (tuple, list)) else (bool(v), bool(v)))
                   for m, v in _method_dict(po.get('interlace'), (False, False), (False, False)).items()}
    diff = _method_dict(po.get('differentiation'), 2, 4)
    p.differentiation = {m: (0 if str(v).lower() == 'fourier' else int(v)) for m, v in diff.items()}
    params = p
    return p


def _lattice_cells(N):
    """Ñ = n³ of species.py:1137-1143: N, N/2 or N/4 for a simple-cubic, body-centred or face-centred pre-initial lattice"""
    N = int(N)
    for per_cell in (1, 2, 4):
        if N % per_cell == 0:
            n = round((N//per_cell)**(1/3))
            if n**3 == N//per_cell:
                return N//per_cell
    return N


def gridsize_value(g, N):
    """A grid size given as a number or as an expression in 'N', 'Ñ', 'gridsize' (= ∛Ñ for particles) and 'nprocs',
    e.g. '2*cbrt(N)' — Component.__init__'s to_float, species.py:1134-1157."""
    if isinstance(g, str):
        from . import communication
        Ñ = _lattice_cells(N)
        expr = g.replace('gridsize', repr(float(round(Ñ**(1/3))))).replace('nprocs', str(communication.nprocs))
        names = {k: getattr(math, k) for k in ('sqrt', 'log', 'log2', 'log10', 'exp', 'floor', 'ceil', 'pi')}
        names.update(cbrt=lambda x: x**(1/3), N=int(N), Ñ=Ñ, min=min, max=max, round=round, π=math.pi)
        try:
            g = eval(expr, {'__builtins__': {}}, names)
        except Exception as exc:
            abort(f'Could not understand the grid size "{g}": {exc}')
    return int(round(float(g)))


def component_differentiation(name, species, method):
    """potential_options['differentiation'] is selected per component (Component.__init__, species.py:1217-1236): an entry
    keyed by the component's name, its species, 'particles' or 'all' wins over 'default' (2 for PM, 4 for P³M);
    'fourier' → 0."""
    spec = params.potential_options_raw.get('differentiation') if isinstance(params.potential_options_raw, dict) else None
    if isinstance(spec, dict):
        lowered = _lower_keys(spec)
        for key in (name, species, 'particles', 'all'):
            if str(key).lower() in lowered:
                v = _method_value(lowered[str(key).lower()], method)
                if v is not None:
                    return 0 if str(v).lower() == 'fourier' else int(v)
    return params.differentiation[method]


def component_softening_length(name, species, N):
    """select_softening_length (commons.py:3862-3873, doc/parameters/physics.rst): a length or an expression in boxsize, N
    and the units, looked up by component name, species, 'particles', 'all', 'default'; default 0.025·boxsize/∛N."""
    spec = user_params.get('select_softening_length', {})
    value = '0.025*boxsize/cbrt(N)'
    if isinstance(spec, dict):
        lowered = _lower_keys(spec)
        for key in (name, species, 'particles', 'all', 'default'):
            if str(key).lower() in lowered:
                value = lowered[str(key).lower()]
                break
    elif spec:
        value = spec
    if isinstance(value, str):
        names = {k: getattr(units, k) for k in vars(units) if not k.startswith('_')}
        names.update(cbrt=lambda x: x**(1/3), sqrt=math.sqrt, N=max(int(N), 1), boxsize=params.boxsize, π=math.pi, pi=math.pi)
        try:
            value = eval(value, {'__builtins__': {}}, names)
        except Exception as exc:
            abort(f'Could not understand the softening length "{value}": {exc}')
    return float(value)


def gridsize_for(method, N):
    """Default PM grid sizes: cbrt(Ñ) for pm, 2·cbrt(Ñ) for p3m, Ñ the cells of the pre-initial lattice
    (doc/parameters/numerics.rst:72-100, species.py:1192-1207)."""
    g = params.gridsize_spec.get(method)
    if isinstance(g, (tuple, list)):
        g = g[0]
    if g is None or g == -1:
        n = round(_lattice_cells(N)**(1/3))
        g = n if method == 'pm' else 2*n
    g = gridsize_value(g, N)
    return g + (g & 1)


def _method_value(spec, method):
    """{'gravity': {'pm': x}} / {'gravity': x} / {'pm': x} / x  →  x for this method (None if absent)"""
    if isinstance(spec, dict):
        spec = _lower_keys(spec)
        if 'gravity' in spec:
            return _method_value(spec['gravity'], method)
        if 'pm' in spec or 'p3m' in spec:
            return spec.get(method)
        return None
    return spec


def _even_gridsize(g):
    """An odd grid size reaches get_fftw_slab() in the reference and aborts there (mesh.py:3784-3789)."""
    g = int(g)
    if g & 1:
        abort(f'An odd grid size ({g}) was given for a potential grid. Some operations may not function correctly.')
    return g


def component_gridsizes(name, species, method, N):
    """(upstream, downstream) potential grid sizes of one component (commons.py:2958-3237,
    doc/parameters/numerics.rst `potential_options`): an entry of potential_options['gridsize'] keyed by the
    component's name or species — an int or an (upstream, downstream) pair — else the default for all components."""
    spec = params.potential_options_raw.get('gridsize') if isinstance(params.potential_options_raw, dict) else None
    if isinstance(spec, dict):
        for key in (name, species, 'particles', 'all', 'default'):
            for k, v in spec.items():
                if str(k).lower() == str(key).lower() and str(k).lower() not in ('global', 'gravity', 'pm', 'p3m'):
                    g = _method_value(v, method)
                    if g is not None:
                        g = tuple(g) if isinstance(g, (tuple, list)) else (g, g)
                        if len(g) == 1:
                            g = g*2
                        if -1 in g:
                            break       # -1: the default for this component (species.py:1169-1171)
                        return tuple(_even_gridsize(gridsize_value(x, N)) for x in g)
    g = gridsize_for(method, N)
    return (g, g)


def global_gridsize(method, components):
    """Global grid size of a potential: potential_options['gridsize']['global'] if given, else the largest
    upstream/downstream grid size in use (doc/parameters/numerics.rst:300-310)."""
    spec = params.potential_options_raw.get('gridsize') if isinstance(params.potential_options_raw, dict) else None
    if isinstance(spec, dict):
        for k, v in spec.items():
            if str(k).lower() == 'global':
                g = _method_value(v, method)
                if g is not None and g != -1:
                    return _even_gridsize(gridsize_value(g, max((c.N for c in components), default=0)))
    return max(max(c.potential_gridsizes['gravity'][method]) for c in components)


def _shortrange_expression(value, names):
    if not isinstance(value, str):
        return float(value)
    env = {k: getattr(units, k) for k in vars(units) if not k.startswith('_')}
    env.update(cbrt=lambda x: x**(1/3), sqrt=math.sqrt, π=math.pi, pi=math.pi, boxsize=params.boxsize, **names)
    try:
        return float(eval(value, {'__builtins__': {}}, env))
    except Exception as exc:
        abort(f'Could not understand the short-range parameter "{value}": {exc}')


def _shortrange_gravity():
    sp = user_params.get('shortrange_params', {})
    if sp and not isinstance(list(sp.values())[0], dict):
        sp = {'gravity': sp}
    return _lower_keys(sp.get('gravity', {})) if isinstance(sp, dict) else {}


def shortrange_scale(gridsize):
    """shortrange_params['gravity']['scale'], a length or an expression in boxsize and gridsize; default
    '1.25*boxsize/gridsize' (commons.py:3254-3269)"""
    scale = _shortrange_gravity().get('scale', None)
    if scale is None:
        return 1.25*params.boxsize/gridsize
    return _shortrange_expression(scale, {'gridsize': gridsize})


def shortrange_range(gridsize):
    """shortrange_params['gravity']['range'], a length or an expression in scale, boxsize and gridsize; default '4.5*scale'"""
    scale = shortrange_scale(gridsize)
    rng = _shortrange_gravity().get('range', None)
    if rng is None:
        return 4.5*scale
    return _shortrange_expression(rng, {'scale': scale, 'gridsize': gridsize})

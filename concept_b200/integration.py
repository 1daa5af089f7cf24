"""Background cosmology and time-step integrals (host side, O(1) scalar work per step).

Mirrors the parts of the reference's integration.py that feed the hot path:
  Spline                   integration.py:40-260  (natural cubic spline, optional log axes)
  hubble                   :570   H = H0·sqrt(Ωm·a⁻³ + 1 − Ωm)  (flat matter + Λ)
  scale_factor / cosmic_time   :602, :621  (splines a(t), t(a) in log–log)
  scalefactor_integral     :712-827  ∫ integrand(a(t)) dt by integrating a spline of the integrand
  init_time                :864-1000
  solve_matterΛ_background :1043-1188  (DOP853, rtol 1e-12, n = int(ln(1/1e-14)/7e-3) points)
CLASS backgrounds (`enable_class_background`) are out of scope: CLASS is not available, and the
reference falls back to exactly this matter + Λ background when it is disabled.
"""
import math

import numpy as np
import scipy.integrate
import scipy.interpolate

from . import commons
from .commons import abort, machine_ϵ, universals


class Spline:
    size_min = 3

    def __init__(self, x, y, name='', *, logx=False, logy=False):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if x.shape[0] != y.shape[0] or x.shape[0] < self.size_min:
            abort(f'Spline "{name}": bad tabulation ({x.shape[0]}, {y.shape[0]})')
        self.name, self.logx, self.logy = name, bool(logx), bool(logy)
        if self.logx and np.any(x <= 0):
            self.logx = False
        self.negativey = False
        if self.logy:
            if np.any(y == 0) or (np.any(y < 0) and np.any(y > 0)):
                self.logy = False
            elif np.any(y < 0):
                self.negativey = True
                y = -y
        xs = np.log(x) if self.logx else x.copy()
        ys = np.log(y) if self.logy else y.copy()
        keep = np.concatenate(([True], np.diff(xs) != 0))   # remove_doppelgängers
        xs, ys = xs[keep], ys[keep]
        self.x = np.exp(xs) if self.logx else xs.copy()
        self.y = (1 - 2*self.negativey)*np.exp(ys) if self.logy else ys.copy()
        self.xmin, self.xmax = xs[0], xs[-1]
        abs_tol = 1e-9*(self.xmax - self.xmin) + machine_ϵ
        self.abs_tol_min = abs_tol + 0.5*(xs[1] - self.xmin)
        self.abs_tol_max = abs_tol + 0.5*(self.xmax - xs[-2])
        self.spline = scipy.interpolate.CubicSpline(xs, ys, bc_type='natural')

    def in_interval(self, x, action='interpolate to'):
        if x < self.xmin:
            if x > self.xmin - self.abs_tol_min:
                return self.xmin
            abort(f'Spline "{self.name}": could not {action} {x}: outside [{self.xmin}, {self.xmax}]')
        elif x > self.xmax:
            if x < self.xmax + self.abs_tol_max:
                return self.xmax
            abort(f'Spline "{self.name}": could not {action} {x}: outside [{self.xmin}, {self.xmax}]')
        return x

    def eval(self, x_in):
        x = math.log(x_in) if self.logx else x_in
        x = self.in_interval(x)
        y = float(self.spline(x))
        if self.logy:
            y = math.exp(y)
            if self.negativey:
                y *= -1
        return y

    def integrate(self, a, b):
        if self.logx or self.logy:
            abort(f'Spline "{self.name}": integration not possible for logged data')
        a = self.in_interval(a, 'integrate from')
        b = self.in_interval(b, 'integrate to')
        sign = 1
        if a > b:
            a, b, sign = b, a, -1
        return sign*float(self.spline.integrate(a, b))


class _TemporalSplines:
    def __init__(self):
        self.reset()

    def reset(self):
        self.a_t = self.t_a = self.a_H = None
        self.initialized = False
        self.key = None


temporal_splines = _TemporalSplines()
spline_t_integrands = {}


def hubble(a=-1):
    p = commons.params
    if not p.enable_Hubble:
        return 0.0
    if a == -1:
        a = universals.a
    ΩΛ = 1 - p.Ωm
    return p.H0*math.sqrt(p.Ωm*np.power(a, -3.0) + ΩΛ)


def _dloga_dlogt(logt, loga):
    t, a = math.exp(logt), math.exp(loga[0])
    return [t*hubble(a)]


def solve_matterΛ_background(a_today=1.0):
    a_begin_bg = 1e-14
    kw = dict(method='DOP853', rtol=1e-12, atol=0)
    t_begin_bg = 2/(3*hubble(a_begin_bg))

    def event(logt, loga):
        return loga[0] - math.log(a_today)
    event.terminal = True
    sol = scipy.integrate.solve_ivp(_dloga_dlogt, (math.log(t_begin_bg), math.inf), np.asarray([math.log(a_begin_bg)]),
                                    events=event, **kw)
    t_today = math.exp(sol.t_events[0][0])
    n_bg = int(math.log(a_today/a_begin_bg)/7e-3)
    logt_values = np.linspace(math.log(t_begin_bg), math.log(t_today), n_bg)
    t_values = np.exp(logt_values)
    a_values = np.exp(scipy.integrate.solve_ivp(_dloga_dlogt, (math.log(t_begin_bg), math.log(t_today)),
                                                [math.log(a_begin_bg)], t_eval=logt_values, **kw).y[0])
    t_values[0], t_values[-1] = t_begin_bg, t_today
    a_values[0], a_values[-1] = a_begin_bg, a_today
    H_values = np.asarray([hubble(a) for a in a_values])
    return a_values, t_values, H_values


def init_time(reinitialize=False):
    p = commons.params
    key = (p.H0, p.Ωm, p.enable_Hubble, p.a_begin, p.t_begin)
    if temporal_splines.initialized and temporal_splines.key == key and not reinitialize:
        universals.t, universals.a = universals.t_begin, universals.a_begin
        return
    spline_t_integrands.clear()
    if p.enable_Hubble:
        a_values, t_values, H_values = solve_matterΛ_background(1.0)
        a_values[-1] = 1.0
        H_values[-1] = p.H0
        temporal_splines.a_t = Spline(a_values, t_values, 't(a)', logx=True, logy=True)
        temporal_splines.t_a = Spline(t_values, a_values, 'a(t)', logx=True, logy=True)
        temporal_splines.a_H = Spline(a_values, H_values, 'H(a)', logx=True, logy=True)
        if 'a_begin' in p.user:
            a_begin = p.a_begin
            t_begin = temporal_splines.a_t.eval(a_begin)
        elif 't_begin' in p.user:
            t_begin = p.t_begin
            a_begin = temporal_splines.t_a.eval(t_begin)
        else:
            abort('No initial scale factor (a_begin) or initial cosmic time (t_begin) specified.')
    else:
        t_begin, a_begin = p.t_begin, 1.0
    temporal_splines.initialized = True
    temporal_splines.key = key
    universals.t_begin, universals.a_begin = t_begin, a_begin
    universals.z_begin = 1/a_begin - 1
    universals.t, universals.a = t_begin, a_begin


def scale_factor(t=-1):
    if not commons.params.enable_Hubble:
        return 1.0
    if t == -1:
        t = universals.t
    if temporal_splines.t_a is None:
        abort('The function a(t) has not been tabulated. Have you called init_time?')
    return temporal_splines.t_a.eval(t)


def cosmic_time(a=-1):
    if not commons.params.enable_Hubble:
        abort('cosmic_time() is only meaningful when Hubble expansion is enabled')
    if a == -1:
        a = universals.a
    if temporal_splines.a_t is None:
        abort('The function t(a) has not been tabulated. Have you called init_time?')
    return temporal_splines.a_t.eval(a)


def _integrand_values(integrand, a, w_effs):
    if integrand in ('1', ''):
        return np.ones_like(a)
    if integrand == 'a**2':
        return a**2
    if integrand == 'a**(-1)':
        return 1/a
    if integrand == 'a**(-2)':
        return 1/a**2
    if integrand == 'ȧ/a':
        return np.asarray([hubble(x) for x in a])
    w = w_effs[0] if w_effs else 0.0
    if integrand == 'a**(-3*w_eff)':
        return a**(-3*w)
    if integrand == 'a**(-3*(1+w_eff))':
        return a**(-3*(1 + w))
    if integrand == 'a**(-3*w_eff-1)':
        return a**(-3*w - 1)
    if integrand == 'a**(3*w_eff-2)':
        return a**(3*w - 2)
    if integrand == 'a**(-3*w_eff)*Γ/H':
        return np.zeros_like(a)      # stable matter: Γ = 0
    if integrand == 'a**(-3*w_eff₀-3*w_eff₁-1)':
        return a**(-3*w_effs[0] - 3*w_effs[1] - 1)
    abort(f'The scale factor integral with "{integrand}" as the integrand is not implemented')


def scalefactor_integral(key, t_start, t_end, all_components=()):
    """∫_{t_start}^{t_end} integrand(a(t)) dt (integration.py:712-827): a natural cubic spline of the
    integrand tabulated on the background's own t grid is integrated exactly."""
    if t_start == t_end:
        return 0.0
    spline = spline_t_integrands.get(key)
    if spline is None:
        if isinstance(key, str):
            integrand, w_effs = key, ()
        else:
            integrand, *names = key
            by_name = {c.name: c for c in all_components}
            w_effs = tuple(by_name[n].w_eff() for n in names)
        if commons.params.enable_Hubble:
            a_tab = temporal_splines.a_t.x
            t_tab = temporal_splines.a_t.y
        else:
            t_tab = np.linspace(t_start, t_end, Spline.size_min)
            a_tab = np.ones_like(t_tab)
        spline = Spline(t_tab, _integrand_values(integrand, np.asarray(a_tab), w_effs), str(integrand))
        if commons.params.enable_Hubble:
            spline_t_integrands[key] = spline
    return spline.integrate(t_start, t_end)

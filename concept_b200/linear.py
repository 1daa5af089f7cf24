"""Linear-theory input of the initial-condition generator.

In the reference these come from CLASS (linear.py: compute_cosmo :2587, compute_transfer :2730; growth
factors growth_fac_D1 … of CosmoResults).  CLASS is not part of this framework (off the hot
path; no network, no CLASS build), so the two look-ups that ic.realize_particles makes are served here
by an analytic stand-in for a flat matter + Λ universe — the same background integration.py uses:

  * δ transfer function  T_δ(k, a) = −(2/5)·(k·c)²/(Ωm·H0²)·T_EH(k)·D1(a)
    with the Eisenstein & Hu (1998) no-wiggle fit T_EH and D1 → a in the matter era; the sign follows
    CLASS (δ < 0 for positive curvature perturbation);
  * θ transfer function  T_θ = −a·H(a)·f1(a)·T_δ        (continuity equation, N-body gauge);
  * growth factors and rates D1, f1, D2, f2, D3a, f3a, D3b, f3b, D3c, f3c from the growth ODEs the reference
    itself integrates when CLASS backgrounds are off (integration.py:1104-1290): D1(a = 1) = 1, all taken positive.

Tabulated CLASS output can be installed instead with `install_transfer(k, delta, theta)` (or by
assigning `concept_b200.ic.compute_transfer` / `.compute_cosmo`, which is what the parity tests do).
"""
import math

import numpy as np
import scipy.integrate

from . import commons
from .integration import hubble


class TransferFunction:
    """The object compute_transfer returns in place of the reference's spline: .eval(k) like
    integration.Spline, plus a vectorised .eval_array(k)."""

    def __init__(self, func):
        self.func = func

    def eval(self, k):
        return float(self.func(np.asarray(k, dtype=np.float64)))

    def eval_array(self, k):
        return np.asarray(self.func(np.asarray(k, dtype=np.float64)), dtype=np.float64)


_installed = {}


def install_transfer(k, delta, theta=None):
    """Use tabulated transfer functions (k in 1/unit_length, δ and optionally θ per unit ζ at the realisation
    time) instead of the analytic stand-in; log–log interpolation.  install_transfer(None, None) removes them."""
    _installed.clear()
    if k is None:
        return
    k = np.asarray(k, dtype=np.float64)
    for variable, table in ((0, delta), (1, theta)):
        if table is None:
            continue
        table = np.asarray(table, dtype=np.float64)
        sign = -1.0 if table[len(table)//2] < 0 else 1.0
        logk, logt = np.log(k), np.log(sign*table)
        _installed[variable] = TransferFunction(lambda q, logk=logk, logt=logt, sign=sign: sign*np.exp(np.interp(np.log(q), logk, logt)))


def eisenstein_hu_nowiggle(k):
    """Eisenstein & Hu 1998 (ApJ 496, 605) eqs. 26, 28-31; k in 1/unit_length."""
    p = commons.params
    h = p.H0/(100*commons.units.km/(commons.units.s*commons.units.Mpc))
    ωm, ωb = p.Ωm*h**2, p.Ωb*h**2
    fb = p.Ωb/p.Ωm
    Θ = 2.7255/2.7
    k_Mpc = np.asarray(k, dtype=np.float64)*commons.units.Mpc            # 1/Mpc
    s = 44.5*math.log(9.83/ωm)/math.sqrt(1 + 10*ωb**0.75)                # Mpc
    α = 1 - 0.328*math.log(431*ωm)*fb + 0.38*math.log(22.3*ωm)*fb**2
    Γ = p.Ωm*h*(α + (1 - α)/(1 + (0.43*k_Mpc*s)**4))
    q = k_Mpc/h*Θ**2/Γ
    L0 = np.log(2*math.e + 1.8*q)
    C0 = 14.2 + 731/(1 + 62.5*q)
    return L0/(L0 + C0*q**2)


class CosmoResults:
    """Growth factors and rates of flat matter + Λ: growth_fac_D1 / f1 / D2 / f2 / D3a / f3a / D3b / f3b / D3c / f3c.
    The same ODE system, matter-era initial conditions and normalisation D1(a = 1) = 1 (D2 ∝ D1², D3 ∝ D1³; all
    taken positive) that the reference integrates itself when CLASS backgrounds are disabled
    (integration.py:1104-1148, dgrowth_da :1190-1290)."""
    keys = ('D1', 'D2', 'D3a', 'D3b', 'D3c')

    def __init__(self):
        p = commons.params
        self.key = (p.H0, p.Ωm)
        Ωm = p.Ωm

        def rhs(a, y):
            E2 = Ωm*a**-3 + 1 - Ωm
            dH_da_over_H = -1.5*Ωm/E2/a**4
            D, dD, D2, dD2, D3a, dD3a, D3b, dD3b, D3c, dD3c = y
            damping = 3/a + dH_da_over_H
            source = -dH_da_over_H/a
            return [dD, -damping*dD + source*D,
                    dD2, -damping*dD2 + source*(D2 + D**2),
                    dD3a, -damping*dD3a + source*(D3a + 2*D**3),
                    dD3b, -damping*dD3b + source*(D3b + 2*D*D2 + 2*D**3),
                    dD3c, -damping*dD3c + source*D**3]
        a0 = 1e-5
        y0 = [a0, 1, 3/7*a0**2, 6/7*a0, 1/3*a0**3, a0**2, 10/21*a0**3, 10/7*a0**2, 1/7*a0**3, 3/7*a0**2]
        self.a_min = a0
        self.sol = scipy.integrate.solve_ivp(rhs, (a0, 1.0 + 1e-9), y0, method='DOP853', rtol=1e-11, atol=0, dense_output=True)
        self.normalization = 1/float(self.sol.sol(1.0)[0])

    def _y(self, a):
        if not self.a_min <= a <= 1.0 + 1e-9:
            commons.abort(f'growth factors requested at a = {a} outside [{self.a_min}, 1]')
        return self.sol.sol(a)

    def growth_unnormalised(self, a):
        """D1 normalised to D1 → a in the matter era (what relates ζ to δ)"""
        return float(self._y(a)[0])

    def _D(self, a, index, power):
        return float(self._y(a)[2*index])*self.normalization**power

    def _f(self, a, index):
        y = self._y(a)
        return float(a*y[2*index + 1]/y[2*index])

    def growth_fac_D1(self, a):
        return self._D(a, 0, 1)

    def growth_fac_f1(self, a):
        return self._f(a, 0)

    def growth_fac_D2(self, a):
        return self._D(a, 1, 2)

    def growth_fac_f2(self, a):
        return self._f(a, 1)

    def growth_fac_D3a(self, a):
        return self._D(a, 2, 3)

    def growth_fac_f3a(self, a):
        return self._f(a, 2)

    def growth_fac_D3b(self, a):
        return self._D(a, 3, 3)

    def growth_fac_f3b(self, a):
        return self._f(a, 3)

    def growth_fac_D3c(self, a):
        return self._D(a, 4, 3)

    def growth_fac_f3c(self, a):
        return self._f(a, 4)


_cosmoresults = None


def compute_cosmo(gridsize_or_k_magnitudes=-1, gauge='synchronous', filename='', class_call_reason=''):
    global _cosmoresults
    p = commons.params
    if _cosmoresults is None or _cosmoresults.key != (p.H0, p.Ωm):
        _cosmoresults = CosmoResults()
    return _cosmoresults


def compute_transfer(component, variable, gridsize_or_k_magnitudes, specific_multi_index=None, a=-1, a_next=-1,
                     gauge='N-body', get='spline', weight=None, backscale=False):
    """(transfer function of δ (variable 0) or θ (variable 1) at scale factor a, cosmoresults)"""
    if variable not in (0, 1):
        commons.abort(f'compute_transfer(): variable {variable} is not available (particle realisations use δ and θ)')
    cosmo = compute_cosmo()
    if variable in _installed:
        return _installed[variable], cosmo
    if a == -1:
        a = commons.universals.a
    p = commons.params
    D1, f1 = cosmo.growth_unnormalised(a), cosmo.growth_fac_f1(a)
    norm = -(2/5)*commons.light_speed**2/(p.Ωm*p.H0**2)*D1
    if variable == 1:
        norm *= -a*hubble(a)*f1
    return TransferFunction(lambda k: norm*k**2*eisenstein_hu_nowiggle(k)), cosmo

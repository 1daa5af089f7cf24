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

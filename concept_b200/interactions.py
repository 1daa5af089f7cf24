"""Interaction registry and the PM / P³M-long-range driver of the host mirror.

Reference: interactions.py — register :2646-2689, find_interactions :2456-2636, gravity :2854-2961,
particle_mesh :1985-2335, apply_particle_mesh_force :2359-2402, get_potential_specs :2786-2821.
The call that main.kick_long makes is unchanged:

    getattr(interactions, force)(method, receivers, suppliers, ᔑdt, interaction_type, printout)

All grid and particle arithmetic happens in libpmgrav.so; this module only sequences the C-ABI calls
and forms the host-side scalars exactly as the reference does.
"""
import collections
import math

import torch

from . import commons, mesh
from .commons import abort, masterprint
from .pmsolver import BCC_SHIFT, make_kick_params

InteractionInfo = collections.namedtuple(
    'InteractionInfo',
    ('force', 'methods', 'conjugated_name', 'dependent', 'affected', 'deterministic', 'instantaneous'),
)
interactions_registered = {}


def register(force, methods, conjugated_name=None, *, dependent=('pos',), affected=('mom',),
             deterministic=True, instantaneous=False):
    """interactions.py:2646-2689"""
    if isinstance(methods, str):
        methods = [methods]
    interactions_registered[force] = InteractionInfo(
        force, list(methods), conjugated_name or force, list(dependent), list(affected), deterministic, instantaneous)


# Methods ordered from the most to the least expensive, as in the reference (interactions.py:2837);
# 'pp' / 'ppnonperiodic' are test-only direct sums in the reference and are not provided here.
register('gravity', ['p3m', 'pm'], 'gravitational')

PotentialSpecs = collections.namedtuple('PotentialSpecs', ('gridsize', 'interpolation_order', 'deconvolve', 'interlace'))
UpDown = collections.namedtuple('UpDown', ('upstream', 'downstream'))


def find_interactions(components, interaction_type='any', instantaneous='both'):
    """interactions.py:2456-2636 for particle components: a list of
    (force, method, receivers, suppliers).  'long-range' selects pm and p3m (its mesh part),
    'short-range' selects p3m (its pair part)."""
    if interaction_type not in ('any', 'long-range', 'short-range'):
        abort(f'find_interactions() called with interaction_type = "{interaction_type}"')
    out = []
    for force, info in interactions_registered.items():
        by_method = collections.OrderedDict()
        for component in components:
            method = component.forces.get(force)
            if not method:
                continue
            if method not in info.methods:
                abort(f'Method "{method}" of force "{force}" (component {component.name}) is not available in '
                      f'concept_b200; available: {info.methods}')
            by_method.setdefault(method, []).append(component)
        for method in info.methods:
            group = by_method.get(method)
            if not group:
                continue
            if interaction_type == 'short-range' and method != 'p3m':
                continue
            # all components participating in this force supply; the group receives
            suppliers = [c for c in components if c.forces.get(force)]
            out.append((force, method, group, suppliers))
    return out


def get_potential_specs(force, method, receivers, suppliers):
    """interactions.py:2786-2821"""
    p = commons.params
    gridsizes = {c.potential_gridsizes[force][method][0] for c in list(receivers) + list(suppliers)}
    if len(gridsizes) != 1:
        abort('concept_b200 requires all components of one PM interaction to share a grid size '
              f'(got {sorted(gridsizes)}); up/down-scaling between grids is out of scope')
    return PotentialSpecs(gridsizes.pop(), p.interpolation_order[method], UpDown(*p.deconvolve[method]),
                          UpDown(*p.interlace[method]))


def gravity(method, receivers, suppliers, ᔑdt, interaction_type, printout=True):
    """interactions.py:2854-2961"""
    force = 'gravity'
    if method not in ('pm', 'p3m'):
        abort(f'gravity() was called with the "{method}" method')
    specs = get_potential_specs(force, method, receivers, suppliers)
    quantity = 'a²ρ'
    ᔑdt_key = ('a**(-3*w_eff)', 'component')
    if method == 'pm':
        if printout:
            masterprint(f'Executing gravitational interaction for {[c.name for c in receivers]} via the PM method ...')
        particle_mesh(receivers, suppliers, specs.gridsize, quantity, force, method, 'gravity',
                      specs.interpolation_order, specs.deconvolve.upstream, specs.deconvolve.downstream,
                      specs.interlace.upstream, specs.interlace.downstream, ᔑdt, ᔑdt_key)
    else:
        if 'any' in interaction_type or 'long' in interaction_type:
            particle_mesh(receivers, suppliers, specs.gridsize, quantity, force, method, 'gravity long-range',
                          specs.interpolation_order, specs.deconvolve.upstream, specs.deconvolve.downstream,
                          specs.interlace.upstream, specs.interlace.downstream, ᔑdt, ᔑdt_key)
        if 'any' in interaction_type or 'short' in interaction_type:
            from . import shortrange
            shortrange.component_component(force, receivers, suppliers, ᔑdt)
    if printout:
        masterprint('done')


def particle_mesh(receivers, suppliers, gridsize_global, quantity, force, method, potential, interpolation_order,
                  deconvolve_upstream, deconvolve_downstream, interlace_upstream, interlace_downstream, ᔑdt, ᔑdt_key):
    """interactions.py:1985-2335 for particle suppliers/receivers sharing one grid size."""
    if not receivers or not suppliers:
        return
    if potential not in ('gravity', 'gravity long-range'):
        abort(f'particle_mesh() got potential "{potential}" ∉ {{"gravity", "gravity long-range"}}')
    if bool(interlace_upstream) != bool(interlace_downstream):
        abort('concept_b200 supports interlacing only when enabled both upstream and downstream')
    p = commons.params
    L, G = p.boxsize, int(gridsize_global)
    ctx = mesh.get_context(G)
    order = int(interpolation_order)
    # both deconvolutions are promoted to the global slab (interactions.py:2069-2080)
    deconv_order_global = order*(int(bool(deconvolve_upstream)) + int(bool(deconvolve_downstream)))
    prefactor = -L**2*commons.G_Newton/math.pi
    gauss = (2*math.pi/L*commons.shortrange_scale(G))**2 if potential == 'gravity long-range' else 0.0
    diff_orders = {c.potential_differentiations[force][method] for c in receivers}
    # Fast path: one component kicked by its own potential — the whole kick is one C call
    if len(receivers) == 1 and len(suppliers) == 1 and receivers[0] is suppliers[0]:
        c = receivers[0]
        kp = make_kick_params(
            mass=c.mass, boxsize=L, gridsize=G, order=order, G_Newton=commons.G_Newton,
            dt_rho_over_dt1=ᔑdt['a**(-3*w_eff-1)', c.name]/ᔑdt['1'], dt_kick=ᔑdt[ᔑdt_key[0], c.name],
            diff_order=diff_orders.pop(), deconvolve=False, interlace=bool(interlace_upstream))
        kp.deconv_order = deconv_order_global
        kp.gauss = gauss
        ctx.kick_long(c.pos_local, c.mom_local, kp)
        return
    # General path: several suppliers and/or receivers
    shifts = [None, BCC_SHIFT] if interlace_upstream else [None]
    nl = len(shifts)
    for l, shift in enumerate(shifts):
        ctx.grid_zero()
        for c in suppliers:
            mesh.interpolate_particles(c, G, ctx, quantity, order, ᔑdt, shift, float(G)**(-3))
        ctx.halo_add()
        ctx.fft_forward()
        if nl > 1:
            ctx.fourier_operate(0, shift, 1.0/nl, -1, False)
            ctx.slab_save() if l == 0 else ctx.slab_accumulate()
    if nl > 1:
        ctx.slab_restore()
    ctx.kspace_potential(prefactor, deconv_order_global, gauss, 1.0)
    need_copy = nl > 1 or 0 in diff_orders or len(diff_orders) > 1
    if need_copy:
        ctx.slab_save()
    first = True
    for diff_order in sorted(diff_orders, reverse=True):
        group = [c for c in receivers if c.potential_differentiations[force][method] == diff_order]
        for l, shift in enumerate(shifts):
            if diff_order == 0:
                for dim in range(3):
                    ctx.fourier_operate(0, shift, 1.0/nl, dim, True)
                    ctx.fft_backward()
                    ctx.halo_fill()
                    for c in group:
                        ctx.gather(0, c.pos_local, c.mom_local, order, dim, c.mass*(-ᔑdt[ᔑdt_key[0], c.name]), shift)
            else:
                if need_copy and not (first and nl == 1):
                    ctx.fourier_operate(0, shift, 1.0/nl, -1, True)
                ctx.fft_backward()
                ctx.halo_fill()
                for c in group:
                    ctx.gather_kick(c.pos_local, c.mom_local, order, diff_order, c.mass*(-ᔑdt[ᔑdt_key[0], c.name]), shift)
            first = False

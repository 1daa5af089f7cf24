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
    components = list(dict.fromkeys(list(receivers) + list(suppliers)))
    return PotentialSpecs(commons.global_gridsize(method, components), p.interpolation_order[method],
                          UpDown(*p.deconvolve[method]), UpDown(*p.interlace[method]))


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
    order = int(interpolation_order)
    gridsizes_upstream = [c.potential_gridsizes[force][method][0] for c in suppliers]
    gridsizes_downstream = [c.potential_gridsizes[force][method][1] for c in receivers]
    if any(g != G for g in gridsizes_upstream + gridsizes_downstream):
        return _particle_mesh_mixed_gridsizes(receivers, suppliers, gridsizes_upstream, gridsizes_downstream, G, quantity,
                                              force, method, potential, order, deconvolve_upstream, deconvolve_downstream,
                                              interlace_upstream, ᔑdt, ᔑdt_key)
    ctx = mesh.get_context(G)
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


def _particle_mesh_mixed_gridsizes(receivers, suppliers, gridsizes_upstream, gridsizes_downstream, G, quantity, force, method,
                                   potential, order, deconvolve_upstream, deconvolve_downstream, interlace, ᔑdt, ᔑdt_key):
    """particle_mesh (interactions.py:1985-2335) when components bring their own upstream / downstream grid sizes:
    suppliers are deposited group by group onto upstream grids, transformed, Nyquist-nullified and *copied* — with the
    upstream deconvolution, the interlacing phase and the half-cell phase between grids — into the global slab
    (interpolate_upstream, mesh.py:492-616; add_upstream_to_global_slabs :618-710; copy_modes :980-1322); the global
    potential is copied to every downstream grid size in use, where the downstream deconvolution, differentiation and
    interpolation happen.  A deconvolution is promoted to the global slab only if all grids on its side have the global
    size (interactions.py:2069-2080).  The accumulating global slab and the finished potential live in the saved
    slabs of the contexts (pm_slab_save), the working slabs being transformed in place.  On several ranks every grid is
    slab-decomposed like the particles and pm_fourier_copy_modes exchanges the rows of the shared mode cube between the
    ranks (the reference's subslab exchange, mesh.py:1105-1230)."""
    from . import communication
    p = commons.params
    for gridsize in set(gridsizes_upstream) | set(gridsizes_downstream) | {G}:
        # every grid is cut into the same x-slabs as the particles (fft.c:105-212 asks for the same divisibility)
        if gridsize % communication.nprocs or (communication.nprocs > 1 and gridsize//communication.nprocs < 8):
            abort(f'Grid size {gridsize} cannot be cut into {communication.nprocs} slabs of at least 8 planes')
    if str(p.grid_dtype) not in ('f64', 'float64'):
        abort('Component-specific upstream/downstream grid sizes are available for fp64 grids only')
    L = p.boxsize
    all_upstream_global = all(g == G for g in gridsizes_upstream)
    all_downstream_global = all(g == G for g in gridsizes_downstream)
    deconv_order_upstream = order*int(bool(deconvolve_upstream) and not all_upstream_global)
    deconv_order_downstream = order*int(bool(deconvolve_downstream) and not all_downstream_global)
    deconv_order_global = order*(int(bool(deconvolve_upstream) and all_upstream_global)
                                 + int(bool(deconvolve_downstream) and all_downstream_global))
    prefactor = -L**2*commons.G_Newton/math.pi
    gauss = (2*math.pi/L*commons.shortrange_scale(G))**2 if potential == 'gravity long-range' else 0.0
    shifts = [None, BCC_SHIFT] if interlace else [None]
    nl = len(shifts)
    ctx_global = mesh.get_context(G, 'f64')

    def global_first(gridsize):
        return (gridsize != G, gridsize)
    # upstream: Σ over groups and lattices into the saved slab of the global context
    first = True
    for gridsize_upstream in sorted(set(gridsizes_upstream), key=global_first):
        ctx = ctx_global if gridsize_upstream == G else mesh.get_context(gridsize_upstream, 'f64')
        group = [c for c, g in zip(suppliers, gridsizes_upstream) if g == gridsize_upstream]
        for shift in shifts:
            ctx.grid_zero()
            for c in group:
                mesh.interpolate_particles(c, gridsize_upstream, ctx, quantity, order, ᔑdt, shift, float(gridsize_upstream)**(-3))
            ctx.halo_add()
            ctx.fft_forward()
            if ctx is ctx_global:
                ctx.fourier_operate(deconv_order_upstream, shift, 1.0/nl, -1, False)      # also nullifies the Nyquist planes
                ctx.slab_save() if first else ctx.slab_accumulate()
            else:
                ctx.fourier_copy_modes_into(ctx_global, deconv_order_upstream, shift, 1.0/nl, src_saved=False, dst_saved=True,
                                            accumulate=not first)
            first = False
    ctx_global.fourier_operate(0, None, 1.0, -1, True)          # working slab = accumulated global slab
    ctx_global.kspace_potential(prefactor, deconv_order_global, gauss, 1.0)
    ctx_global.slab_save()
    # downstream
    for gridsize_downstream in sorted(set(gridsizes_downstream), key=global_first):
        if gridsize_downstream == G:
            ctx = ctx_global
        else:
            ctx = mesh.get_context(gridsize_downstream, 'f64')
            ctx_global.fourier_copy_modes_into(ctx, 0, None, 1.0, src_saved=True, dst_saved=True, accumulate=False)
        group_downstream = [c for c, g in zip(receivers, gridsizes_downstream) if g == gridsize_downstream]
        diff_orders = {c.potential_differentiations[force][method] for c in group_downstream}
        for diff_order in sorted(diff_orders, reverse=True):
            group = [c for c in group_downstream if c.potential_differentiations[force][method] == diff_order]
            for shift in shifts:
                if diff_order == 0:
                    for dim in range(3):
                        ctx.fourier_operate(deconv_order_downstream, shift, 1.0/nl, dim, True)
                        ctx.fft_backward()
                        ctx.halo_fill()
                        for c in group:
                            ctx.gather(0, c.pos_local, c.mom_local, order, dim, c.mass*(-ᔑdt[ᔑdt_key[0], c.name]), shift)
                else:
                    ctx.fourier_operate(deconv_order_downstream, shift, 1.0/nl, -1, True)
                    ctx.fft_backward()
                    ctx.halo_fill()
                    for c in group:
                        ctx.gather_kick(c.pos_local, c.mom_local, order, diff_order, c.mass*(-ᔑdt[ᔑdt_key[0], c.name]), shift)

"""P³M short-range force and rung scheduling — host side.

Reference: main.kick_short :1173-1238, main.driftkick_short :1347-1624, main.initialize_rung_populations
:1639-1675; interactions.component_component :122-329 with gravity_pairwise_shortrange (gravity.py:263-354),
get_shortrange_table (gravity.py:373-421), compute_factors (gravity.py:51-67), get_softened_r3inv
(interactions.py:1847-1897); Component rung methods (species.py:2253-2598).

The host keeps the reference's schedule (which rungs are kicked when, over which time interval, with
which ᔑdt_rungs index); pairs, Δmom, accelerations and rung indices live on the GPU (pm_shortrange,
pm_apply_dmom, pm_assign_rungs, pm_flag_rung_jumps, pm_apply_rung_jumps).  On several GPUs the particles within `range` of a slab
face reach the neighbour as read-only ghosts inside pm_shortrange (the analogue of domain_domain's sendrecv_component,
interactions.py:398-590), rung populations are summed over the ranks, and Δmom / rung indices migrate with their particles.
"""
import ctypes
import math

import numpy as np
import torch

from . import commons, communication, interactions, mesh
from ._lib import check
from .commons import abort, machine_ϵ, universals

ᔑdt_rungs = {}
_tables = {}


def N_rungs():
    return commons.params.N_rungs


def shortrange_params(gridsize):
    """shortrange_params['gravity'] with its defaults (commons.py:3254-3269)"""
    scale = commons.shortrange_scale(gridsize)
    rng = commons.shortrange_range(gridsize)
    tablesize = int(commons._shortrange_gravity().get('tablesize', 2**12))
    return scale, rng, tablesize


def get_softened_r3inv(r2, ϵ):
    """interactions.py:1847-1897 — spline softening kernel (the default, commons.py:3862)"""
    h = 2.8*ϵ
    r = math.sqrt(r2)
    if r >= h:
        return 1/(r2*r)
    u = r/h
    if u < 0.5:
        return 32/h**3*(1./3. + u**2*(-6./5. + u))
    return 32/(3*r**3)*(u**3*(2 + u*(-9./2. + u*(18./5. - u))) - 3./480.)


def get_shortrange_table(gridsize, softening, device):
    """gravity.py:373-421, cached per (grid size, softening); returns (device tensor, maxr2, range, size)"""
    scale, rng, size = shortrange_params(gridsize)
    return build_shortrange_table(scale, rng, size, softening, device)


def build_shortrange_table(scale, rng, size, softening, device):
    """The tabulation itself (gravity.py:395-421) for an explicit scale r_s, range and table size"""
    key = (scale, rng, size, softening, str(device))
    if key in _tables:
        return _tables[key]
    maxr2 = (1 + 1/size)*rng**2
    r_tab = np.sqrt(np.linspace(0, maxr2, size))
    table = np.empty(size, dtype=np.float64)
    for i in range(size - 1):
        r2 = 0.5*(r_tab[i]**2 + r_tab[i + 1]**2)
        r = math.sqrt(r2)
        x = r*(1/scale)
        r3_inv = 1/(r2*r)
        table[i] = (-r3_inv*(1/math.sqrt(math.pi)*x*math.exp(-(0.5*x)**2) + (math.erfc(0.5*x) - 1))
                    - get_softened_r3inv(r2, softening))
    table[size - 1] = np.nan
    _tables[key] = (torch.as_tensor(table, device=device), maxr2, rng, size)
    return _tables[key]


# ---------------------------------------------------------------------------------------------
# Component-side rung state (species.py:956-996, 2290-2598)
# ---------------------------------------------------------------------------------------------
def ensure_rung_state(c):
    n = c.N_allocated
    if getattr(c, 'Δmom', None) is None or c.Δmom.shape[0] < n:
        c.Δmom = torch.zeros((n, 3), dtype=torch.float64, device=c.device)
        c.rung_indices = torch.zeros(n, dtype=torch.int8, device=c.device)
        c.rung_indices_jumped = torch.zeros(n, dtype=torch.int8, device=c.device)
        c.rungs_N = [0]*N_rungs()
        c.rungs_N[0] = c.N_local
        c.lowest_populated_rung = c.highest_populated_rung = c.lowest_active_rung = 0
        c.use_rungs = N_rungs() > 1


def _set_populations(c, counts):
    # the reference keeps per-process populations and reduces the lowest/highest populated rung over the processes
    # (species.py:2547-2598); every rank must follow the same kick schedule, so the populations are summed right away
    c.rungs_N = communication.allreduce_ints([int(v) for v in counts])
    populated = [r for r, v in enumerate(c.rungs_N) if v > 0]
    c.lowest_populated_rung = populated[0] if populated else N_rungs() - 1
    c.highest_populated_rung = populated[-1] if populated else 0


def get_rung_factor(c, Δt, fac_softening):
    return 0.5*math.log2(Δt**2/(2*fac_softening*c.softening_length))


def assign_rungs(c, Δt, fac_softening):
    ctx = c._pm_context()
    counts = (ctypes.c_int64*N_rungs())()
    check(ctx.lib.pm_assign_rungs(ctx._h, c.Δmom.data_ptr(), c.N_local, get_rung_factor(c, Δt, fac_softening), N_rungs(),
                                  c.rung_indices.data_ptr(), c.rung_indices_jumped.data_ptr(), counts))
    _set_populations(c, counts)


def flag_rung_jumps(c, Δt, Δt_jump_fac, fac_softening):
    if not c.use_rungs:
        return False
    ctx = c._pm_context()
    dt1 = np.ascontiguousarray(ᔑdt_rungs['1'], dtype=np.float64)
    any_jump = ctypes.c_int(0)
    check(ctx.lib.pm_flag_rung_jumps(ctx._h, c.Δmom.data_ptr(), c.N_local, c.rung_indices.data_ptr(),
                                     c.rung_indices_jumped.data_ptr(), int(c.lowest_active_rung),
                                     get_rung_factor(c, Δt*Δt_jump_fac, fac_softening), get_rung_factor(c, Δt/Δt_jump_fac, fac_softening),
                                     dt1.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), N_rungs(), ctypes.byref(any_jump)))
    return bool(communication.allreduce_ints([int(any_jump.value)])[0])     # any process (species.py:2506-2510)


def apply_rung_jumps(c):
    ctx = c._pm_context()
    counts = (ctypes.c_int64*N_rungs())()
    check(ctx.lib.pm_apply_rung_jumps(ctx._h, c.N_local, c.rung_indices.data_ptr(), c.rung_indices_jumped.data_ptr(), N_rungs(), counts))
    _set_populations(c, counts)


def apply_and_convert_Δmom(c, apply):
    """apply_Δmom (species.py:2253-2266) + convert_Δmom_to_acc (species.py:2290-2325)"""
    ctx = c._pm_context()
    a = universals.a
    conv = a**(3*c.w_eff(a=a))/(c.mass*(machine_ϵ + np.asarray(ᔑdt_rungs['a**2'], dtype=np.float64)))
    conv = np.ascontiguousarray(conv)
    check(ctx.lib.pm_apply_dmom(ctx._h, c.mom.data_ptr(), c.Δmom.data_ptr(), c.N_local, c.rung_indices.data_ptr(),
                                c.rung_indices_jumped.data_ptr(), int(c.lowest_active_rung),
                                conv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(conv), int(bool(apply))))


# ---------------------------------------------------------------------------------------------
# The pair interaction (interactions.component_component with gravity_pairwise_shortrange)
# ---------------------------------------------------------------------------------------------
def component_component(force, receivers, suppliers, ᔑdt_rungs_arg):
    if len(receivers) != 1 or len(suppliers) != 1 or receivers[0] is not suppliers[0]:
        abort('concept_b200: the short-range force is implemented for one particle component')
    c = receivers[0]
    ensure_rung_state(c)
    ctx = c._pm_context()
    G = c.potential_gridsizes['gravity']['p3m'][0]
    table, maxr2, rng, size = get_shortrange_table(G, c.softening_length, c.device)
    # compute_factors (gravity.py:51-67)
    factors = commons.G_Newton*c.mass*c.mass*np.asarray(ᔑdt_rungs_arg['a**(-3*w_eff₀-3*w_eff₁-1)', c.name, c.name], dtype=np.float64)
    factors = np.ascontiguousarray(factors)
    check(ctx.lib.pm_shortrange(ctx._h, c.pos.data_ptr(), c.N_local, c.rung_indices.data_ptr(), c.rung_indices_jumped.data_ptr(),
                                int(c.lowest_active_rung), factors.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(factors),
                                rng, table.data_ptr(), size, maxr2, c.Δmom.data_ptr()))


# ---------------------------------------------------------------------------------------------
# Scheduling (main.kick_short, main.driftkick_short)
# ---------------------------------------------------------------------------------------------
def _reset_ᔑdt_rungs(components):
    from . import main
    scalar = main.get_time_step_integrals(0, 0, components)
    for key in scalar:
        if key != '__names' and key not in ᔑdt_rungs:
            ᔑdt_rungs[key] = np.zeros(3*N_rungs() - 1, dtype=np.float64)
    for key in list(ᔑdt_rungs):
        if key not in scalar:
            del ᔑdt_rungs[key]


def _store(ᔑdt_rung, index):
    for integrand, integral in ᔑdt_rung.items():
        if integrand != '__names':
            ᔑdt_rungs[integrand][index] = integral


def kick_short(components, Δt, fake=False):
    """main.py:1173-1238: half a sub-step kick for every rung; fake=True assigns rungs instead
    (main.initialize_rung_populations, main.py:1639-1675)."""
    from . import main
    particle_components = [c for c in components if c.representation == 'particles']
    interactions_list = interactions.find_interactions(particle_components, 'short-range')
    if not interactions_list:
        return
    _reset_ᔑdt_rungs(particle_components)
    for c in particle_components:
        ensure_rung_state(c)
        c.lowest_active_rung = c.lowest_populated_rung
    highest_populated_rung = max(c.highest_populated_rung for c in particle_components)
    t_start = universals.t
    for rung_index in range(highest_populated_rung + 1):
        t_end = t_start + Δt/2**(rung_index + 1)
        _store(main.get_time_step_integrals(t_start, t_end, particle_components), rung_index)
    receivers_all = []
    # nullify_Δ('mom') (species.py:3717-3742) only clears ACTIVE particles; pm_shortrange overwrites exactly
    # those, and inactive particles keep their stored acceleration (needed by later rung decisions)
    for force, method, receivers, suppliers in interactions_list:
        getattr(interactions, force)(method, receivers, suppliers, ᔑdt_rungs, 'short-range', False)
        receivers_all += [r for r in receivers if r not in receivers_all]
    fac_softening = main._facs()['softening']
    if fake:
        for c in receivers_all:
            apply_and_convert_Δmom(c, apply=False)
        for c in particle_components:
            assign_rungs(c, Δt, fac_softening)
    else:
        for c in receivers_all:
            apply_and_convert_Δmom(c, apply=True)


def driftkick_short(components, Δt, sync_time):
    """main.py:1347-1624: 2^(N_rungs−1) synchronous sub-drifts interleaved with rung-selective kicks."""
    from . import main
    Δt_reltol, Δt_jump_fac = main.Δt_reltol, main.Δt_jump_fac
    fac_softening = main._facs()['softening']
    particle_components = [c for c in components if c.representation == 'particles']
    interactions_list = interactions.find_interactions(particle_components, 'short-range')
    nr = N_rungs()
    _reset_ᔑdt_rungs(particle_components)
    for c in particle_components:
        ensure_rung_state(c)
    any_kicks = True
    index_start = 0
    clip = Δt_reltol*Δt + 2*machine_ϵ

    def t_at(index):
        t = universals.t + Δt*(float(index)/2**nr)
        return sync_time if t + clip > sync_time else t
    for driftkick_index in range(2**(nr - 1)):
        if any_kicks:
            index_start = 2*driftkick_index
        lowest_active_rung = nr - 1
        for rung_index in range(nr):
            if (driftkick_index + 1) % 2**(nr - 1 - rung_index) == 0:
                lowest_active_rung = rung_index
                break
        any_kicks = False
        for c in particle_components:
            c.lowest_active_rung = max(lowest_active_rung, c.lowest_populated_rung)
            if c.highest_populated_rung >= c.lowest_active_rung:
                any_kicks = True
        if not any_kicks:
            continue
        index_end = 2*driftkick_index + 2
        t_start, t_end = t_at(index_start), t_at(index_end)
        if t_end > t_start:
            ᔑdt = main.get_time_step_integrals(t_start, t_end, particle_components)
            for c in particle_components:
                c.drift(ᔑdt)
                c.lowest_active_rung = max(lowest_active_rung, c.lowest_populated_rung)
        highest_populated_rung = max(c.highest_populated_rung for c in particle_components)
        for rung_index in range(lowest_active_rung, highest_populated_rung + 1):
            i0 = 2**(nr - 1 - rung_index) + (driftkick_index//2**(nr - 1 - rung_index))*2**(nr - rung_index)
            t0 = t_at(i0)
            _store(main.get_time_step_integrals(t0, t_at(i0 + 2**(nr - rung_index)), particle_components), rung_index)
            if rung_index > 0 and ((driftkick_index + 1) - 2**(nr - 1 - rung_index)) % 2**(nr - rung_index) == 0:
                _store(main.get_time_step_integrals(t0, t_at(i0 + 2**(nr - 1 - rung_index)), particle_components), rung_index + nr)
            else:
                for integrals in ᔑdt_rungs.values():
                    integrals[rung_index + nr] = -1
            if rung_index < nr - 1:
                _store(main.get_time_step_integrals(t0, t_at(i0 + 3*2**(nr - 2 - rung_index)), particle_components), rung_index + 2*nr)
        if sum(ᔑdt_rungs['1'][lowest_active_rung:highest_populated_rung + 1]) == 0:
            continue
        any_rung_jumps = [flag_rung_jumps(c, Δt, Δt_jump_fac, fac_softening) for c in particle_components]
        receivers_all = []
        for force, method, receivers, suppliers in interactions_list:
            getattr(interactions, force)(method, receivers, suppliers, ᔑdt_rungs, 'short-range', False)
            receivers_all += [r for r in receivers if r not in receivers_all]
        for c in receivers_all:
            apply_and_convert_Δmom(c, apply=True)
        for c, jumped in zip(particle_components, any_rung_jumps):
            if jumped:
                apply_rung_jumps(c)


# ---------------------------------------------------------------------------------------------
# The pair kick on bare device arrays (bench.py's configs[4] leg, tools): one rung, all particles active
# ---------------------------------------------------------------------------------------------
class PairKick:
    """mom += G·m²·Δt·Σ_j (x⃗_i − x⃗_j)·table[r²] over all pairs within the range (gravity.py:263-354 with one rung), on
    capacity-sized device arrays as bench.py keeps them.  scale r_s = 1.25 cells, range 4.5·r_s, 4096 table entries and
    spline softening 0.025·L/∛N are the reference's defaults (commons.py:3254-3269, :3862-3873)."""

    def __init__(self, ctx, boxsize, gridsize, n_total, capacity, G_Newton, mass=1.0, dt=1e-3):
        self.ctx = ctx
        dev = ctx.torch_device
        self.scale = 1.25*boxsize/gridsize
        self.range = 4.5*self.scale
        softening = 0.025*boxsize/round(n_total**(1/3))
        self.table, self.maxr2, _, self.size = build_shortrange_table(self.scale, self.range, 2**12, softening, dev)
        self.dmom = torch.zeros((capacity, 3), dtype=torch.float64, device=dev)
        self.rung = torch.zeros(capacity, dtype=torch.int8, device=dev)
        self.factors = (ctypes.c_double*1)(G_Newton*mass*mass*dt)
        self.conv = (ctypes.c_double*1)(1.0)

    def kick(self, pos, mom, n):
        ctx = self.ctx
        check(ctx.lib.pm_shortrange(ctx._h, pos.data_ptr(), int(n), self.rung.data_ptr(), self.rung.data_ptr(), 0, self.factors, 1,
                                    self.range, self.table.data_ptr(), self.size, self.maxr2, self.dmom.data_ptr()))
        check(ctx.lib.pm_apply_dmom(ctx._h, mom.data_ptr(), self.dmom.data_ptr(), int(n), self.rung.data_ptr(), self.rung.data_ptr(),
                                    0, self.conv, 1, 1))

    def exchange(self, pos, mom, n):
        return self.ctx.exchange(pos, mom, None, n, Δmom=self.dmom, rung_indices=self.rung, rung_indices_jumped=self.rung)

    def pair_stats(self, pos, n):
        """(pairs within the range, candidates examined) of one pair kernel over the current particles"""
        ctx = self.ctx
        check(ctx.lib.pm_shortrange_stats(ctx._h, 1, None, None))
        check(ctx.lib.pm_shortrange(ctx._h, pos.data_ptr(), int(n), self.rung.data_ptr(), self.rung.data_ptr(), 0, self.factors, 1,
                                    self.range, self.table.data_ptr(), self.size, self.maxr2, self.dmom.data_ptr()))
        pairs, cands = ctypes.c_int64(0), ctypes.c_int64(0)
        check(ctx.lib.pm_shortrange_stats(ctx._h, 0, ctypes.byref(pairs), ctypes.byref(cands)))
        return pairs.value//2, cands.value

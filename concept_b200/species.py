"""Component — the particle data model of the host mirror (reference: species.py:852-2199).

Particle state lives on the GPU for the whole run: `pos`, `mom` are torch CUDA tensors of shape
(N_allocated, 3), float64, AoS exactly like the reference's `double* pos/mom` of length 3·N_allocated
(species.py:1411-1431); `ids` (int64) travel with the particles through slab migration.  On several
GPUs a Component holds the particles of this rank's x-slab only (N_local of N).
"""
import numpy as np
import torch

from . import commons, communication, mesh
from .commons import abort


class Component:
    representation = 'particles'

    def __init__(self, name, species, *, N=-1, mass=-1, gridsize=-1, boltzmann_order=1, boltzmann_closure=None):
        if boltzmann_order not in (1, -2) or gridsize not in (-1, None):
            abort('concept_b200 implements particle components only (fluids are out of scope, SURVEY.md §2)')
        self.name, self.species = str(name), str(species)
        self.N = int(N)
        self.mass = float(mass)
        self.N_local = 0
        self.N_allocated = 0
        self.pos = self.mom = self.ids = None
        p = commons.params
        # forces (commons.py:3664-3702): particles default to gravity via p3m
        forces = None
        for key in (self.name, self.species, 'all', 'particles', 'default'):
            if key in p.select_forces:
                forces = dict(p.select_forces[key])
                break
        self.forces = forces if forces is not None else {'gravity': 'p3m'}
        self.potential_gridsizes = {'gravity': {}}
        self.potential_differentiations = {'gravity': {}}
        for method in ('pm', 'p3m'):
            self.potential_gridsizes['gravity'][method] = (
                commons.component_gridsizes(self.name, self.species, method, self.N) if self.N > 0 else (None, None))
            self.potential_differentiations['gravity'][method] = commons.component_differentiation(self.name, self.species, method)
        # select_softening_length, default 0.025·L/∛N (commons.py:3862-3873)
        self.softening_length = commons.component_softening_length(self.name, self.species, self.N)
        self._ϱ_bar = -1

    # -- storage ------------------------------------------------------------------------------
    @property
    def device(self):
        return torch.device('cuda', communication.local_rank)

    def resize(self, size):
        """Grow the particle buffers (species.py:2002-2065); contents up to N_local are kept."""
        size = int(size)
        if size <= self.N_allocated:
            return
        def grow(old, shape, dtype):
            new = torch.zeros(shape, dtype=dtype, device=self.device)
            if old is not None and self.N_local:
                new[:self.N_local] = old[:self.N_local]
            return new
        self.pos = grow(self.pos, (size, 3), torch.float64)
        self.mom = grow(self.mom, (size, 3), torch.float64)
        self.ids = grow(self.ids, (size,), torch.int64)
        if getattr(self, 'Δmom', None) is not None:      # P³M state follows (species.py:2040-2060)
            self.Δmom = grow(self.Δmom, (size, 3), torch.float64)
            self.rung_indices = grow(self.rung_indices, (size,), torch.int8)
            self.rung_indices_jumped = grow(self.rung_indices_jumped, (size,), torch.int8)
        self.N_allocated = size

    def populate(self, data, var):
        """species.py:1911-1926: `var` ∈ posx,posy,posz,momx,momy,momz (one column) or 'pos'/'mom'/'ids'
        (whole (n,3) / (n,) array).  Sets N_local to the length of the data."""
        t = torch.as_tensor(np.ascontiguousarray(data) if isinstance(data, np.ndarray) else data)
        n = t.shape[0]
        if n > self.N_allocated:
            keep = self.N_local
            self.N_local = min(keep, self.N_allocated)
            self.resize(n)
        self.N_local = n
        if var in ('pos', 'mom'):
            getattr(self, var)[:n] = t.to(self.device, torch.float64)
        elif var == 'ids':
            self.ids[:n] = t.to(self.device, torch.int64)
        else:
            prefix, suffix = var[:-1], var[-1]
            if prefix not in ('pos', 'mom') or suffix not in 'xyz':
                abort(f'populate() called with var = "{var}"')
            getattr(self, prefix)[:n, 'xyz'.index(suffix)] = t.to(self.device, torch.float64)
        if var != 'ids' and self.ids is not None and not getattr(self, '_ids_set', False):
            self.ids[:n] = torch.arange(n, dtype=torch.int64, device=self.device)

    def set_particles(self, pos, mom, ids=None, distribute=True):
        """Load a full particle set given on every rank (numpy or torch, (N,3)); with several ranks
        each keeps the particles of its own x-slab (the initial domain decomposition)."""
        pos = torch.as_tensor(pos, dtype=torch.float64)
        mom = torch.as_tensor(mom, dtype=torch.float64)
        n = pos.shape[0]
        if self.N <= 0:
            self.N = n
        ids = torch.arange(n, dtype=torch.int64) if ids is None else torch.as_tensor(ids, dtype=torch.int64)
        if distribute and communication.nprocs > 1:
            G = self.potential_gridsizes['gravity'][self.forces.get('gravity', 'pm')][0]
            keep = communication.slab_owner(pos[:, 0], commons.params.boxsize, G) == communication.rank
            pos, mom, ids = pos[keep], mom[keep], ids[keep]
        n_local = pos.shape[0]
        cap = n_local if communication.nprocs == 1 else int(1.5*n/communication.nprocs) + 1024
        self.N_local = 0
        self.resize(max(cap, n_local))
        self.pos[:n_local] = pos.to(self.device)
        self.mom[:n_local] = mom.to(self.device)
        self.ids[:n_local] = ids.to(self.device)
        self.N_local = n_local

    @property
    def pos_local(self):
        return self.pos[:self.N_local]

    @property
    def mom_local(self):
        return self.mom[:self.N_local]

    # host views in the reference's naming (copies: the live data is on the device)
    @property
    def pos_mv3(self):
        return self.pos_local.cpu().numpy()

    @property
    def mom_mv3(self):
        return self.mom_local.cpu().numpy()

    def gather_global(self):
        """(pos, mom) of all N particles ordered by id, on the host (tests / snapshots)."""
        pos, mom, ids = self.pos_local.cpu().numpy(), self.mom_local.cpu().numpy(), self.ids[:self.N_local].cpu().numpy()
        if communication.nprocs > 1:
            parts = communication.allgather((pos, mom, ids))
            pos = np.concatenate([p[0] for p in parts])
            mom = np.concatenate([p[1] for p in parts])
            ids = np.concatenate([p[2] for p in parts])
        n = len(ids)
        if n == 0 or (ids[0] == 0 and ids[-1] == n - 1 and bool(np.all(ids[1:] > ids[:-1]))):
            return pos.copy(), mom.copy()    # already in id order (one rank, never re-ordered); copies: never views of live data
        if int(ids.min()) == 0 and int(ids.max()) == n - 1:
            # ids are a permutation of 0 … n−1: one scatter instead of a sort (1.5 s for 256³ particles)
            pos_out, mom_out = np.empty_like(pos), np.empty_like(mom)
            pos_out[ids] = pos
            mom_out[ids] = mom
            return pos_out, mom_out
        order = np.argsort(ids, kind='stable')
        return pos[order], mom[order]

    # -- physics -------------------------------------------------------------------------------
    def w_eff(self, a=-1, t=-1):
        return 0.0      # matter; species.py:3016 in general

    def ẇ(self, a=-1):
        return 0.0

    def Γ(self, a=-1):
        return 0.0

    def is_active(self, a=-1):
        return True

    @property
    def ϱ_bar(self):
        """species.py:1792-1819: (Ωb + Ωcdm)·ρ_crit for matter, else N·mass/boxsize³."""
        if self._ϱ_bar != -1:
            return self._ϱ_bar
        p = commons.params
        sp = self.species.lower()
        if sp in ('matter', 'baryons + cold dark matter', 'cold dark matter + baryons'):
            self._ϱ_bar = (p.Ωb + p.Ωcdm)*p.ρ_crit
        elif sp in ('cold dark matter', 'cdm', 'dark matter'):
            self._ϱ_bar = p.Ωcdm*p.ρ_crit
        elif sp in ('baryons', 'baryon', 'b'):
            self._ϱ_bar = p.Ωb*p.ρ_crit
        else:
            self._ϱ_bar = self.N*self.mass/p.boxsize**3
        return self._ϱ_bar

    def realize(self, a=-1, a_next=-1, variables=None, multi_indices=None, use_gridˣ=False):
        """species.py:2094-2098 → ic.realize (ic.py:297-398): for a particle component the full realisation"""
        if not self.is_active(a):
            return
        from . import ic
        ic.realize_particles(self, commons.universals.a if a == -1 else a)

    def cell_sort(self, gridsize=None):
        """Reorder the local particles by grid cell (the tile_sort analogue, species.py:2657-2780): keeps the
        deposit/gather locality that lattice-ordered particles lose over many steps."""
        ctx = self._pm_context() if gridsize is None else mesh.get_context(gridsize)
        n = self.N_local
        if getattr(self, 'Δmom', None) is None:
            ctx.sort_particles(self.pos, self.mom, self.ids, n)
            return
        # P³M state (Δmom, rung_indices, rung_indices_jumped; species.py:956-996) follows the particles: sort a
        # position index along with them and apply the same permutation to the per-particle rung arrays
        index = torch.arange(n, dtype=torch.int64, device=self.device)
        ids = self.ids[:n].clone()
        ctx.sort_particles(self.pos, self.mom, index, n)
        self.ids[:n] = ids[index]
        self.Δmom[:n] = self.Δmom[:n][index]
        self.rung_indices[:n] = self.rung_indices[:n][index]
        self.rung_indices_jumped[:n] = self.rung_indices_jumped[:n][index]

    def _pm_context(self):
        method = self.forces.get('gravity', 'pm')
        method = method if method in ('pm', 'p3m') else 'pm'
        return mesh.get_context(self.potential_gridsizes['gravity'][method][0])

    def drift(self, ᔑdt, a_next=-1):
        """species.py:2179-2199: pos = mod(pos + mom·ᔑdt['a**(-2)']·a^{3w}/mass, boxsize); exchange."""
        a = commons.universals.a
        Δt_over_mass = ᔑdt['a**(-2)']*a**(3*self.w_eff(a=a))/self.mass
        ctx = self._pm_context()
        ctx.drift(self.pos_local, self.mom_local, Δt_over_mass)
        self.exchange()

    def exchange(self):
        """communication.exchange (communication.py:135-517) for x-slabs: pm_exchange."""
        if communication.nprocs == 1:
            return
        ctx = self._pm_context()
        from ._lib import PMError
        for attempt in range(8):
            try:
                self.N_local = ctx.exchange(self.pos, self.mom, self.ids, self.N_local, Δmom=getattr(self, 'Δmom', None),
                                            rung_indices=getattr(self, 'rung_indices', None),
                                            rung_indices_jumped=getattr(self, 'rung_indices_jumped', None))
                return
            except PMError as err:
                if err.status != -6:     # PM_ERR_OVERFLOW: this rank's buffers are too small for its arrivals
                    raise
                # the arrivals wait in the mailbox: grow (species.py:2002-2065 grows by realloc the same way) and unpack again
                self.resize(int(1.3*self.N_allocated) + 4096)
        abort(f'Component "{self.name}": particle exchange keeps overflowing')

    def sum_mom2(self):
        return communication.allreduce_sum(self._pm_context().sum_mom2(self.mom_local)) if self.N_local or communication.nprocs > 1 else 0.0

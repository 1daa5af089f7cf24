"""Mesh layer of the host mirror: the cache of libpmgrav contexts and the one mesh function that forms scalars on the
host.

Reference counterparts (mesh.py): get_fftw_slab :3769-3866 (cached slabs + plans, never freed) → get_context;
free_fftw_slab :3897 → free_contexts; interpolate_particles :1512-1636 (its contribution scalar, :1542-1573) →
interpolate_particles.  The other mesh operators of the reference (fft :4012, fourier_operate :3327, nullify_modes :3545,
diff_domaingrid :4874, interpolate_domaingrid_to_particles :376, copy_modes :980) have no host-side body here: they are
entry points of the C ABI (include/pmgrav.h), called as methods of the context (pmsolver.PMContext) from interactions.py and
analysis.py.  Nothing here touches grid values.
"""
import torch

from . import commons, communication
from .pmsolver import PMContext

_contexts = {}


def get_context(gridsize, dtype=None):
    """One context per (gridsize, dtype), created on first use and kept (mesh.py:3861-3864)."""
    dtype = dtype or commons.params.grid_dtype
    key = (int(gridsize), str(dtype), float(commons.params.boxsize), communication.rank, communication.nprocs)
    ctx = _contexts.get(key)
    if ctx is None:
        if not torch.cuda.is_available():
            raise RuntimeError('concept_b200 needs a CUDA device: there is no CPU fallback')
        ctx = PMContext(gridsize, commons.params.boxsize, dtype=dtype, rank=communication.rank,
                        nranks=communication.nprocs, device=communication.local_rank)
        if communication.nprocs > 1:
            ctx.connect(communication.bcast, communication.allgather, communication.master)
        _contexts[key] = ctx
    return ctx


def free_contexts():
    """free_fftw_slab analogue (mesh.py:3897)"""
    for ctx in _contexts.values():
        ctx.close()
    _contexts.clear()


def interpolate_particles(component, gridsize, ctx, quantity, order, ᔑdt=None, shift=None, factor=1.0):
    """mesh.py:1512-1636: add the component's `quantity` to the context's real grid.
    contribution per particle: 'ρ' a^(−3(1+w))·m, 'a²ρ' a^(−3w−1)·m, 'ϱ' m — or their time-step
    averages when ᔑdt is given (mesh.py:1542-1573) — times factor·(G/L)³."""
    a = commons.universals.a
    w_eff = component.w_eff(a=a)
    if quantity == 'ρ':
        contribution = ᔑdt['a**(-3*(1+w_eff))', component.name]/ᔑdt['1'] if ᔑdt else a**(-3*(1 + w_eff))
    elif quantity == 'a²ρ':
        contribution = ᔑdt['a**(-3*w_eff-1)', component.name]/ᔑdt['1'] if ᔑdt else a**(-3*w_eff - 1)
    elif quantity == 'ϱ':
        contribution = 1
    else:
        commons.abort(f'interpolate_particles() called with quantity = "{quantity}" ∉ {{"ρ", "a²ρ", "ϱ"}}')
    contribution *= component.mass
    contribution_factor = factor*(gridsize/commons.params.boxsize)**3
    contribution *= contribution_factor
    ctx.deposit(component.pos_local, order, contribution, shift)

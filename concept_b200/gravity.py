"""The reference's gravity module surface for the short-range force (gravity.py:40-67, :263-354, :373-421) —
the pairwise-kernel plugin of interactions.component_component (`func_interaction`, interactions.py:61-75).

In the reference the caller enumerates tile and subtile pairings and hands them to the plugin, which loops over
the particle pairs of those tiles.  Here the pair enumeration lives inside the kernel (cell lists built on the
device, csrc/pm_shortrange.cu), so the plugin is called once per (receiver, supplier) with no tile selection; a
call that passes tile index arrays cannot be honoured and aborts.
"""
import numpy as np

from . import commons, communication, shortrange
from .commons import abort


def compute_factors(receiver, supplier, ᔑdt_rungs):
    """gravity.py:51-67: G·m_r·m_s·ᔑdt_rungs['a**(-3*w_eff₀-3*w_eff₁-1)', receiver, supplier] per rung index"""
    ᔑdt_arr = np.asarray(ᔑdt_rungs['a**(-3*w_eff₀-3*w_eff₁-1)', receiver.name, supplier.name], dtype=np.float64)
    return commons.G_Newton*receiver.mass*supplier.mass*ᔑdt_arr


def combine_softening_lengths(ϵᵢ, ϵⱼ):
    """interactions.py:1820-1830: the arithmetic mean (Hernquist & Barnes 1990)"""
    return 0.5*(ϵᵢ + ϵⱼ)


def get_shortrange_table(softening, gridsize, device=None):
    """gravity.py:373-421: the tabulated short-range factor −r⁻³(x/√π·e^{−x²/4} + erfc(x/2) − 1) − r⁻³_softened over r²,
    as (device tensor, maxr2, range, size); `gridsize` fixes the scale r_s = 1.25·L/gridsize (commons.py:3254-3269)."""
    import torch
    device = device if device is not None else torch.device('cuda', 0)
    return shortrange.get_shortrange_table(gridsize, softening, device)


def gravity_pairwise_shortrange(interaction_name, receiver, supplier, ᔑdt_rungs, rank_supplier=0, only_supply=False,
                                pairing_level='domain', tile_indices_receiver=None, tile_indices_supplier_paired=None,
                                tile_indices_supplier_paired_N=None, extra_args=None):
    """gravity.py:263-354 with the plugin signature of interactions.py:61-75.  Δmom of the receiver's active particles
    is overwritten with the short-range momentum updates (see pm_shortrange in include/pmgrav.h)."""
    if tile_indices_receiver is not None or tile_indices_supplier_paired is not None:
        abort('gravity_pairwise_shortrange(): tile selections cannot be passed in — concept_b200 enumerates the pairs '
              'inside the kernel (csrc/pm_shortrange.cu)')
    if rank_supplier != communication.rank:
        abort('gravity_pairwise_shortrange(): non-local suppliers are not available (single-GPU short-range force)')
    shortrange.component_component('gravity', [receiver], [supplier], ᔑdt_rungs)

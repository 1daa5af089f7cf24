"""concept_b200 — B200-native particle-mesh gravity behind CO*N*CEPT's operator surface.

Host mirror of the reference's modules for the PM hot path (same names, argument meaning and error
behaviour): commons (parameters, units, universals), species.Component, interactions.gravity /
particle_mesh / find_interactions, main.kick_long / driftkick_short / timeloop, integration
(background, time-step integrals), analysis.measure('v_rms').  All arithmetic on particles and grids
runs in csrc/ (libpmgrav.so, hand-written sm_100a kernels + cuFFT + NCCL) through include/pmgrav.h.
"""
__version__ = '0.1'

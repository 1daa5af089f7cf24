/* pmgrav.h — C ABI of libpmgrav.so: the B200 (sm_100a) particle-mesh gravity hot path.
 *
 * Drop-in boundary for CO*N*CEPT's PM long-range kick + drift.  Every entry point
 * names the reference interface it replaces (file:line under the reference's src/).
 * The reference's only native FFI is fft.c (fftw_setup / fftw_execute / fftw_clean,
 * mesh.py:43-71, fft.c:75-83,105-112,281-285); pm_create / pm_fft / pm_destroy are the
 * equivalents.  The remaining entry points replace functions that the reference
 * compiles from Python to C through Cython (`cdef` functions inside mesh.so,
 * interactions.so, species.so) and that a maintainer would re-bind with ctypes/cffi
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only; all particle/grid pointers are DEVICE pointers unless the
 *     name says `_host`;
 *   - particles are AoS double[n][3] (x,y,z interleaved) exactly like Component.pos /
 *     .mom (species.py:1411-1431);
 *   - every function returns 0 on success or a negative pm_status; pm_last_error()
 *     gives the message.  Nothing falls back to a CPU path;
 *   - work is enqueued on the context's CUDA stream; no hidden host synchronisation
 *     except in the `_host` functions and pm_get_grid;
 *   - one context per (rank, gridsize, grid dtype), reused forever — like the cached
 *     slabs/plans of get_fftw_slab (mesh.py:3769-3866).
 */
#ifndef PMGRAV_H
#define PMGRAV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pm_ctx pm_ctx;

typedef enum {
    PM_OK = 0,
    PM_ERR_ARG = -1,       /* invalid argument (the reference abort()s, commons.py:1002) */
    PM_ERR_CUDA = -2,      /* CUDA runtime / cuFFT failure */
    PM_ERR_ALLOC = -3,
    PM_ERR_STATE = -4,     /* call order violated (e.g. gather before solve) */
    PM_ERR_COMM = -5,      /* NCCL failure / communicator missing */
    PM_ERR_OVERFLOW = -6   /* a particle buffer is too small for migrating particles */
} pm_status;

enum { PM_GRID_F64 = 0, PM_GRID_F32 = 1 };

/* Which internal grid pm_get_grid() copies out */
enum {
    PM_TAP_REAL = 0,     /* real-space grid [nx_local][G][G] (padding stripped): density after
                            pm_deposit(+halo), potential after pm_fft_backward */
    PM_TAP_FOURIER = 1,  /* Fourier slab as doubles (re,im interleaved):
                            nranks==1: natural [i][j][kk];  nranks>1: [j_local][i][kk]
                            (the FFTW-MPI transposed layout of fft.c:55-72) */
    PM_TAP_FORCE = 2     /* scratch force grid written by pm_diff() */
};

/* ---- library ---------------------------------------------------------- */
const char* pm_version(void);
const char* pm_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t pm_launch_count(void);
int pm_device_count(void);

/* ---- context: replaces fftw_setup / get_fftw_slab (fft.c:105-212, mesh.py:3769-3866) */
/* gridsize G: cells per side; must be even and G % nranks == 0 (mesh.py:3779-3783).
 * boxsize  L: in the caller's length unit.
 * grid_dtype: PM_GRID_F64 (reference-exact) or PM_GRID_F32 (mixed precision).
 * rank/nranks: x-slab decomposition — rank r owns planes [r·G/P, (r+1)·G/P).
 * device: CUDA device ordinal; stream: cudaStream_t (NULL = the legacy default stream). */
int pm_create(pm_ctx** out, int gridsize, double boxsize, int grid_dtype,
              int rank, int nranks, int device, void* stream);
int pm_destroy(pm_ctx* ctx);                                   /* fftw_clean, fft.c:281 */
int pm_set_stream(pm_ctx* ctx, void* stream);
int pm_sync(pm_ctx* ctx);
/* local slab geometry, like fftw_return_struct (fft.c:75-83) */
int pm_local_shape(const pm_ctx* ctx, int64_t* nx_local, int64_t* x_start,
                   int64_t* nj_local, int64_t* j_start);
int64_t pm_device_bytes(const pm_ctx* ctx);

/* ---- multi-GPU: NCCL communicator owned by the library ------------------ */
/* id_out: 128-byte ncclUniqueId created on rank 0 and broadcast by the host code */
int pm_comm_unique_id(void* id_out_128);
int pm_comm_init(pm_ctx* ctx, const void* id_128);
/* sum a device array of n doubles over all ranks in place (allreduce, analysis.py:3971) */
int pm_allreduce_sum(pm_ctx* ctx, double* dev_values, int n);
/* CUDA-IPC mapping of every rank's slab into every rank (same node, NVLink/NVSwitch peer access):
 * each rank exports a 64-byte handle; the host code all-gathers them (rank order) and hands the
 * nranks×64 bytes to pm_ipc_open_peers. */
int pm_ipc_get_handle(pm_ctx* ctx, void* handle_out_64);
int pm_ipc_open_peers(pm_ctx* ctx, const void* handles_nranks_x_64);

/* ---- mesh operators ------------------------------------------------------ */
/* get_buffer(..., nullify=True) for 'grid_updownstream' (mesh.py:600) */
int pm_grid_zero(pm_ctx* ctx);
/* interpolate_particles (mesh.py:1512-1636): grid[cell] += (wx·contribution)·wy·wz.
 * order 1..4 = NGP, CIC, TSC, PCS (set_weights_*, mesh.py:5305-5379).
 * contribution = q·mass·(G/L)³·G⁻³ (mesh.py:1550-1573, 582).
 * shift[3]: lattice shift in grid units (mesh.py:85-100), NULL = (0,0,0). */
int pm_deposit(pm_ctx* ctx, const double* pos, int64_t n, int order,
               double contribution, const double* shift);
/* communicate_ghosts(grid,'+=') (communication.py:563-660, mesh.py:609): fold the x-halo
 * planes into the neighbour slabs.  No-op for nranks == 1 (periodic wrap is done in-kernel). */
int pm_halo_add(pm_ctx* ctx);
/* communicate_ghosts(grid,'=') after the inverse transform (mesh.py:2243): fill x-halo planes */
int pm_halo_fill(pm_ctx* ctx);
/* the same, only as many planes as a gather of interpolation order `order` (1..4) with finite differences of order
 * `diff_order` (0 = none) reaches beyond the slab (+1 with the interlacing shift): what pm_kick_long itself fills */
int pm_halo_fill_for(pm_ctx* ctx, int order, int diff_order, int interlace);
/* fft(slab,'forward'|'backward') (mesh.py:4012-4157 → fftw_execute): unnormalised, in place.
 * forward also performs nullify_modes(slab,'nyquist') only when asked via pm_kspace_*. */
int pm_fft_forward(pm_ctx* ctx);
int pm_fft_backward(pm_ctx* ctx);
/* The potential loop of particle_mesh (interactions.py:2092-2118) fused with
 * nullify_modes 'nyquist' (mesh.py:3591-3622) and 'origin' (:3585-3590):
 *   slab[k] *= [Π x_l/sin x_l]^deconv_order · scale · prefactor/k² · exp(−k²·gauss)
 * prefactor = −L²·G_N/π ; gauss = (2π·r_s/L)² or 0 (plain PM) ; scale = 1/n_lattices.
 * If prefactor == 0 the 1/k² and exp terms are skipped (pure deconvolution, used by
 * the power-spectrum path, fourier_operate mesh.py:3327-3400). */
int pm_kspace_potential(pm_ctx* ctx, double prefactor, int deconv_order, double gauss,
                        double scale);
/* fourier_operate (mesh.py:3327-3400): in-place deconvolution^deconv_order · scale,
 * phase rotation by θ = −2π/G·k·shift, and (diff_dim ∈ {0,1,2}) ×i·(2π/L)·k_dim.
 * If `from_saved` the operation reads the saved copy made by pm_slab_save and writes the
 * working slab (the reference copies slab_downstream → slab_updownstream_subgroup). */
int pm_fourier_operate(pm_ctx* ctx, int deconv_order, const double* shift, double scale,
                       int diff_dim, int from_saved);
/* pm_fft_forward + pm_kspace_potential + pm_fft_backward in one call: 2-D (y,z) transforms per x plane,
 * then ONE kernel for the x direction — forward transform, Green's function, inverse transform — that
 * addresses all ranks' slabs through peer pointers (the FFTW-MPI transpose of fft.c:34-73 never
 * materialises), then the inverse 2-D transforms.  The slab ends in real space (potential).
 * Implementations (pm_set_fused_solve):
 *   PM_SOLVE_FFT2_L2        hand-written transforms, G ∈ {128, 256, 512}; the two passes of each 2-D
 *                           transform run dependency-ordered in one launch so the intermediate plane
 *                           stays in L2 (fp64 grids; fp32 grids use PM_SOLVE_FFT2_SPLIT)
 *   PM_SOLVE_FFT2_SPLIT     same kernels, one launch per pass
 *   PM_SOLVE_CUFFT2D_XSOLVE batched cuFFT 2-D plans + the x kernel, G ∈ {64, 512}
 *   PM_SOLVE_UNFUSED        pm_kick_long uses the three-call path (cuFFT 3-D + k-space kernel)
 *   PM_SOLVE_AUTO           the first available of the above (default)
 * With several ranks pm_ipc_open_peers must have been called. */
enum { PM_SOLVE_UNFUSED = 0, PM_SOLVE_AUTO = 1, PM_SOLVE_CUFFT2D_XSOLVE = 2, PM_SOLVE_FFT2_SPLIT = 3, PM_SOLVE_FFT2_L2 = 4 };
int pm_solve_fused(pm_ctx* ctx, double prefactor, int deconv_order, double gauss);
/* One kernel of pm_solve_fused at a time, for per-kernel timing: stage 1 = forward 2-D transforms,
 * 2 = x solve, 3 = inverse 2-D transforms (call 1, 2, 3 in order; 0 = all).  Hand-written transforms only. */
int pm_solve_fused_stage(pm_ctx* ctx, double prefactor, int deconv_order, double gauss, int stage);
int pm_fused_solve_available(const pm_ctx* ctx);
int pm_set_fused_solve(pm_ctx* ctx, int mode);
/* Host-synchronising check of the sticky flags that kernels raise instead of hanging or failing silently: a tile
 * dependency of the dependency-ordered transforms that was not satisfied within seconds, a rank that did not arrive at a
 * device barrier (PM_ERR_COMM), and — several ranks — a deposit contribution that fell outside this rank's x-slab plus halo
 * (the particle is not distributed by slab; the reference's domain decomposition makes that impossible by construction,
 * communication.py:135-517).  Returns the error code with a message if any is set. */
int pm_check_async_error(pm_ctx* ctx);
/* The mode loop of compute_powerspec (analysis.py:500-547) on the Fourier slab: for every mode of
 * fourier_loop(gridsize, sparse=True, skip_origin=True, k2_max) (mesh.py:2615-2890)
 *   power[k²] += re² + im²   and, if count != NULL,   count[k²] += 1
 * (device arrays of k2_max + 1 entries, not zeroed here; k² in grid units).  With several ranks every
 * rank adds its part of the slab; the host code sums over ranks (Reduce, analysis.py:549-553). */
int pm_power_k2(pm_ctx* ctx, int k2_max, double* power, unsigned long long* count);
int pm_slab_save(pm_ctx* ctx);      /* slab_updownstream_subgroup[...] = slab (interactions.py:2256) */
int pm_slab_accumulate(pm_ctx* ctx);/* saved += working slab (copy_modes '+=' for interlacing) */
int pm_slab_restore(pm_ctx* ctx);   /* working slab = saved */
/* diff_domaingrid (mesh.py:4874-5030) into the scratch force grid; order ∈ {1,2,4,6,8} */
int pm_diff(pm_ctx* ctx, int dim, int order);
/* interpolate_domaingrid_to_particles (mesh.py:376-459) + apply_particle_mesh_force
 * (interactions.py:2359-2402) for one dimension from an explicit force grid:
 *   mom[dim] += factor · Σ W·grid[cell];  which = PM_TAP_REAL or PM_TAP_FORCE */
int pm_gather(pm_ctx* ctx, int which, const double* pos, double* mom, int64_t n,
              int order, int dim, double factor, const double* shift);
/* Fused diff_domaingrid ×3 + gather ×3 + kick from the real-space potential:
 *   mom[d] += factor · Σ W·(∂_d φ)[cell],  factor = −mass·ᔑdt['a**(-3*w_eff)'].
 * diff_order ∈ {1,2,4,6,8}.  If sum_mom2 != NULL, Σ mom² of the updated momenta is
 * ADDED to *sum_mom2 (device double; measure('v_rms'), analysis.py:3965-3972). */
int pm_gather_kick(pm_ctx* ctx, const double* pos, double* mom, int64_t n, int order,
                   int diff_order, double factor, const double* shift, double* sum_mom2);

/* pm_gather_kick with Component.drift (species.py:2179-2199) of the same particles fused in:
 * after the kick, pos = mod(pos + mom·dt_over_mass, L) with the kicked momenta (dt_over_mass == 0: none).
 * pos and mom cross HBM once for both operators. */
int pm_gather_kick_drift(pm_ctx* ctx, double* pos, double* mom, int64_t n, int order, int diff_order,
                         double factor, const double* shift, double* sum_mom2, double dt_over_mass);

/* ---- particle operators ------------------------------------------------ */
/* Component.drift (species.py:2179-2199): pos = mod(pos + mom·dt_over_mass, L) */
int pm_drift(pm_ctx* ctx, double* pos, const double* mom, int64_t n, double dt_over_mass);
/* measure(component,'v_rms') reduction: *out += Σ mom² (device double) */
int pm_sum_mom2(pm_ctx* ctx, const double* mom, int64_t n, double* out);
/* Reorder the local particles by grid cell (x plane, y row, z; stable) — the analogue of
 * Component.tile_sort (species.py:2657-2780): deposit and gather rely on consecutive particles touching
 * neighbouring grid rows.  Lattice-ordered initial conditions have that order; call this every few
 * hundred steps (or after loading unordered particles).  ids may be NULL. */
int pm_sort_particles(pm_ctx* ctx, double* pos, double* mom, int64_t* ids, int64_t n);
/* exchange(component) (communication.py:135-517) for the x-slab decomposition: particles
 * whose owner floor(x/L·P) differs from this rank are sent to their owner.  pos/mom/ids are
 * device arrays with room for `capacity` particles; *n_inout is updated (host int64).
 * ids may be NULL. */
int pm_exchange(pm_ctx* ctx, double* pos, double* mom, int64_t* ids, int64_t* n_inout,
                int64_t capacity);
/* The same for a P³M component: the short-range state of a particle — Δmom (the stored acceleration) and its
 * rung indices (species.py:956-996; the reference ships them in exchange() too, communication.py:300-330) — migrates
 * with it.  dmom, rung and rung_jumped may be NULL.
 * Both run over the CUDA-IPC peer mappings (pm_ipc_open_peers): movers are written straight into the owner's mailbox,
 * arrivals fill the holes the movers leave (communication.py:430-517).  PM_ERR_OVERFLOW on a rank whose buffers are too
 * small for its arrivals: grow them (contents preserved) and call again — only that rank repeats its unpacking. */
int pm_exchange_rungs(pm_ctx* ctx, double* pos, double* mom, int64_t* ids, double* dmom, signed char* rung,
                      signed char* rung_jumped, int64_t* n_inout, int64_t capacity);

/* ---- P3M short range --------------------------------------------------------------
 * On several ranks every rank calls pm_shortrange together: particles within `range` of a slab face are handed to the
 * neighbour as read-only ghosts over the CUDA-IPC peer mappings (pm_ipc_open_peers) before the pair kernel runs. */
/* gravity_pairwise_shortrange over all pairs within `range` (gravity.py:263-354; pair enumeration
 * interactions.py:1353-1791), gather form:
 *   dmom[i] = factors[rung_jumped[i]] · Σ_{j≠i, r²≤range²} (x_i − x_j)·table[int(r²·(tablesize−1)/maxr2)]
 * for receivers with rung[i] ≥ lowest_active_rung (others untouched); periodic minimum image.
 * factors (host, nfactors = 3·N_rungs−1) = G·m²·ᔑdt_rungs['a**(-3*w_eff₀-3*w_eff₁-1)'] (gravity.py:51-67);
 * table_dev: device copy of get_shortrange_table() (gravity.py:373-421). */
int pm_shortrange(pm_ctx* ctx, const double* pos, int64_t n, const signed char* rung,
                  const signed char* rung_jumped, int lowest_active_rung, const double* factors_host,
                  int nfactors, double range, const double* table_dev, int tablesize, double maxr2,
                  double* dmom);
/* Counters of the pair kernel (measurement aid): with enable != 0 every following pm_shortrange counts the pairs within
 * the range (each counted from both sides) and the candidates it looked at; the counts of the last such call are returned
 * (host-synchronising).  Either output may be NULL. */
int pm_shortrange_stats(pm_ctx* ctx, int enable, int64_t* pairs_out, int64_t* candidates_out);
/* apply_Δmom (species.py:2253-2266; skipped when apply = 0, the "fake" kick) then
 * convert_Δmom_to_acc (species.py:2290-2325): dmom *= conv[rung_jumped], active rungs only */
int pm_apply_dmom(pm_ctx* ctx, double* mom, double* dmom, int64_t n, const signed char* rung,
                  const signed char* rung_jumped, int lowest_active_rung, const double* conv_host,
                  int nconv, int apply);
/* assign_rungs (species.py:2415-2435); rungs_N_host[n_rungs] receives the populations */
int pm_assign_rungs(pm_ctx* ctx, const double* acc, int64_t n, double rung_factor, int n_rungs,
                    signed char* rung, signed char* rung_jumped, int64_t* rungs_N_host);
/* flag_rung_jumps (species.py:2461-2510); dt1_host = ᔑdt_rungs['1'] (3·n_rungs−1 doubles) */
int pm_flag_rung_jumps(pm_ctx* ctx, const double* acc, int64_t n, const signed char* rung,
                       signed char* rung_jumped, int lowest_active_rung, double rung_factor_up,
                       double rung_factor_down, const double* dt1_host, int n_rungs, int* any_host);
/* apply_rung_jumps (species.py:2523-2545) */
int pm_apply_rung_jumps(pm_ctx* ctx, int64_t n, signed char* rung, signed char* rung_jumped, int n_rungs,
                        int64_t* rungs_N_host);

/* ---- initial conditions (PM_GRID_F64 contexts, lattice size == gridsize) ------- */
/* preinitialize_particles (ic.py:2138-2247): the rank's x-slab of the n³ lattice (n = gridsize), cell-centred,
 * pos = ((x_start + ½ + shift) + i)·L/n …, mom = 0, ids (may be NULL) = id_bgn + ((i·n + j)·n + k).
 * Particles index_bgn … index_bgn + n_local − 1 of the device arrays are written; shift[3] is the
 * lattice shift in grid units (mesh.py:85-100), NULL = (0,0,0). */
int pm_ic_lattice(pm_ctx* ctx, double* pos, double* mom, int64_t* ids, const double* shift, int64_t index_bgn,
                  int64_t id_bgn, int64_t* n_local_out);
/* realize_grid (ic.py:670-782; scalar, output_space='Fourier', then nullify_modes origin + nyquist) fused with
 * laplacian_inverse (mesh.py:3422-3437) into the working slab:
 *   slab[k] = amplitudes[k²]·noise[k]·e^{iθ(k, −shift)} · (−lap_factor/k_f²)/k²
 * noise: device doubles (re, im) in the slab layout of PM_TAP_FOURIER (the primordial noise of
 * generate_primordial_noise, ic.py:928-1163, drawn by the host code); amplitudes: device table over integer k²
 * with k2_max + 1 entries (get_amplitudes, ic.py:542-627).  lap_factor == 0 leaves the inverse Laplacian out
 * (plain realize_grid). */
int pm_ic_potential(pm_ctx* ctx, const double* noise, const double* amplitudes, int k2_max, const double* shift,
                    double lap_factor);
/* Local non-Gaussianity of realize_grid (ic.py:766-771) on the real-space grid: x += f_nl·x².  The caller
 * realises with lap_factor = 0, transforms back, calls this, transforms forward and applies the inverse Laplacian
 * and the forward normalisation G⁻³ with pm_kspace_potential. */
int pm_ic_nongaussianity(pm_ctx* ctx, double f_nl);
/* displace_particles (ic.py:2249-2283) from the real-space grid: lattice particle (i, j, k) gets
 *   pos[dim] += pos_factor·ψ[i][j][k],  mom[dim] += mom_factor·ψ[i][j][k]      (either array may be NULL) */
int pm_ic_displace(pm_ctx* ctx, double* pos, double* mom, int64_t index_bgn, int dim, double pos_factor,
                   double mom_factor);
/* pos = mod(pos, L) (ic.py:1396-1398) */
int pm_ic_wrap(pm_ctx* ctx, double* pos, int64_t n);
/* copy the real-space grid, padding stripped, into a device buffer of nx_local·G·G doubles */
int pm_real_export(pm_ctx* ctx, double* dev_out);
/* source of the 2LPT potential (carryout_2lpt, ic.py:1553-1575) into the real-space grid:
 *   −Φ,₀₀Φ,₁₁ − Φ,₁₁Φ,₂₂ − Φ,₂₂Φ,₀₀ + Φ,₀₁² + Φ,₁₂² + Φ,₂₀²   from six exported second-derivative grids */
int pm_ic_2lpt_source(pm_ctx* ctx, const double* d00, const double* d11, const double* d22, const double* d01,
                      const double* d12, const double* d02);
/* One term of an LPT source (handle_lpt_term, ic.py:1895-2057; the 3LPT potentials of carryout_3lpt_a/b/c,
 * :1619-1893) on compact device grids of n doubles: acc (= | +=) factor·a·b[·third]   (third may be NULL).
 * `ctx` only supplies the stream. */
int pm_lpt_accumulate(pm_ctx* ctx, double* acc, int64_t n, double factor, const double* a, const double* b,
                      const double* third, int assign);
/* the inverse of pm_real_export: a compact device grid of nx_local·G·G doubles becomes the real-space grid */
int pm_real_import(pm_ctx* ctx, const double* dev_in);
/* resize_grid(…, 'fourier') as the dealiased LPT terms use it (ic.py:2093-2108, :2027-2034): the working
 * Fourier slab of `src` is copied into the one of `dst` (another grid size, same device, one rank each) for the
 * modes |k| < min(G_src, G_dst)/2; every other mode of `dst` is nullified. */
int pm_fourier_resize(pm_ctx* src, pm_ctx* dst);

/* copy_modes(slab_from, slab_onto, deconv_order, lattice, '=' | '+=') between grids of DIFFERENT size
 * (mesh.py:980-1322), as used for component-specific upstream/downstream grids (add_upstream_to_global_slabs,
 * mesh.py:618-710; interactions.py:2120-2140).  For every mode inside the cube |k| < min(G_src, G_dst)/2:
 *   dst (+)= scale·[Π x_l/sin x_l]^deconv_order·e^{i(θ_cell + θ_shift)}·src
 * deconvolution and θ_shift = −2π/G_src·k·shift evaluated for the SOURCE grid, θ_cell = (π/G_dst − π/G_src)·(ki+kj+kk)
 * the half-cell offset between cell-centred grids.  '=' (accumulate == 0) nullifies every other mode of dst.
 * src_saved / dst_saved select the saved copy (pm_slab_save) instead of the working slab.  Several ranks (both contexts
 * distributed over the same ranks, a collective call): the rows of the shared cube are exchanged between the ranks that
 * hold them in the two slab decompositions (the subslab exchange of mesh.py:1105-1230) over dst's communicator. */
int pm_fourier_copy_modes(pm_ctx* src, pm_ctx* dst, int deconv_order, const double* shift, double scale,
                          int src_saved, int dst_saved, int accumulate);

/* ---- whole-path entry points --------------------------------------------- */
typedef struct {
    int order;            /* interpolation order 1..4 */
    int diff_order;       /* 0 = Fourier differentiation, else 1,2,4,6,8 */
    int deconv_order;     /* total deconvolution power (order·(up+down)), interactions.py:2069-2080 */
    int interlace;        /* 0 / 1 (bcc, two lattices) */
    double contribution;  /* deposit scalar, see pm_deposit */
    double prefactor;     /* −L²·G_N/π */
    double gauss;         /* (2π·r_s/L)² for 'gravity long-range', else 0 */
    double kick_factor;   /* −mass·ᔑdt['a**(-3*w_eff)'] */
} pm_kick_params;
/* gravity('pm'|'p3m', [c], [c], ᔑdt, 'long-range') for one particle component
 * (interactions.py:2854-2961 → particle_mesh :1985-2335).  sum_mom2 as in pm_gather_kick. */
int pm_kick_long(pm_ctx* ctx, const double* pos, double* mom, int64_t n,
                 const pm_kick_params* p, double* sum_mom2);
/* kick_long immediately followed by Component.drift (species.py:2179-2199) of the same particles with
 * the kicked momenta, pos = mod(pos + mom·dt_over_mass, L) — what main.timeloop does across a step
 * boundary (kick_long at the end of one step, driftkick_short's drift at the start of the next,
 * main.py:255-361) whenever the next step's Δt is already fixed.  On the default path the drift is
 * fused into the gather/kick kernel (pos and mom cross HBM once).  dt_over_mass == 0: no drift. */
int pm_kick_drift(pm_ctx* ctx, double* pos, double* mom, int64_t n, const pm_kick_params* p,
                  double dt_over_mass, double* sum_mom2);
/* Same, with HOST particle buffers: H2D, kick, optional drift (dt_over_mass != 0), D2H.
 * pos_host is updated only when drifting. Synchronous. */
int pm_kick_long_host(pm_ctx* ctx, double* pos_host, double* mom_host, int64_t n,
                      const pm_kick_params* p, double dt_over_mass, double* sum_mom2_host);

/* ---- parity taps ------------------------------------------------------------ */
/* copies the chosen grid to host as float64; host_out must hold pm_tap_size() doubles */
int64_t pm_tap_size(const pm_ctx* ctx, int which);
int pm_get_grid(pm_ctx* ctx, int which, double* host_out);
/* overwrite the real-space grid from host float64 [nx_local][G][G] (tests) */
int pm_set_grid(pm_ctx* ctx, const double* host_in);

#ifdef __cplusplus
}
#endif
#endif /* PMGRAV_H */

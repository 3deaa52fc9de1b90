/*
 * gravhopper_b200.h -- C ABI of libgravhopper_b200.so, the B200 (sm_100a) gravity engine that
 * replaces GravHopper's C backend and step loop.
 *
 * What each entry point replaces in the reference (/root/reference, v1.2.0):
 *
 *   gh_direct_summation           gravhopper/_jbgrav.c:68-135   (wrapper) + :140-193 (workhorse)
 *   gh_direct_summation_position  gravhopper/_jbgrav.c:200-294  (wrapper) + :299-353 (workhorse)
 *   gh_tree_force                 gravhopper/_jbgrav.c:564-630  (wrapper) + :737-806, :360-558
 *   gh_tree_force_position        gravhopper/_jbgrav.c:634-727  (wrapper) + :737-806, :360-558
 *   gh_engine_*                   gravhopper/gravhopper.py:293-342 (Simulation.run / init_run),
 *                                 :405-416 (perform_timestep, DKD leapfrog), :419-459
 *                                 (calculate_acceleration incl. the external-force sum)
 *
 * Conventions (identical to the reference's _jbgrav level unless noted):
 *   - positions are (N,3) row-major float64, masses (N,) float64, raw units kpc / Msun, G = 1;
 *     accelerations come back as (N,3) row-major float64 in Msun/kpc^2;
 *   - the engine works in GravHopper's internal units kpc, km/s, Msun, Myr (gravhopper.py:150-154)
 *     and applies the two unit factors of jbgrav.py:48 and gravhopper.py:409 itself;
 *   - `prec` selects the pair arithmetic: GH_PREC_F64 (reference-exact contract: <=1e-12) or
 *     GH_PREC_F32 (fp32 pair maths, two-level fp32->fp64 accumulation: <=1e-5).  State, moments,
 *     kick and drift are always float64;
 *   - `mem` says where the caller's buffers live: GH_MEM_HOST (the library copies in and out) or
 *     GH_MEM_DEVICE (pointers are CUDA device pointers on the current device, e.g. a torch
 *     tensor's data_ptr(); nothing leaves the device);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the library's own stream / legacy
 *     default).  With GH_MEM_DEVICE the call is asynchronous on that stream;
 *   - every function returns 0 on success, a GH_E* code otherwise, and never calls exit();
 *     gh_last_error() returns the calling thread's last message;
 *   - inputs are never modified; the library never frees caller memory;
 *   - one engine handle must be driven by one thread at a time; different handles are
 *     independent.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * GH_ECUDA.
 */
#ifndef GRAVHOPPER_B200_H
#define GRAVHOPPER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GH_OK 0
#define GH_EINVAL 1  /* bad argument (shape, precision, null pointer) */
#define GH_ECUDA 2   /* CUDA runtime error or no device */
#define GH_ENOMEM 3  /* allocation failed (the reference calls exit(209|435) here) */
#define GH_ESTATE 4  /* engine used before upload, etc. */

#define GH_PREC_F32 32
#define GH_PREC_F64 64

#define GH_MEM_HOST 0
#define GH_MEM_DEVICE 1

#define GH_ALG_DIRECT 0
#define GH_ALG_TREE 1

/* unit factors the engine applies (astropy CODATA-2018 definitions; jbgrav.py:40,48) */
#define GH_C_ACC 4.398600412921223e-09           /* km/s/Myr per (Msun/kpc^2)  */
#define GH_KPC_PER_KMS_MYR 1.022712165045695e-3  /* kpc per (km/s * Myr)       */
#define GH_G 4.30091727003628e-06                /* kpc (km/s)^2 / Msun        */

const char *gh_last_error(void);
int gh_version(void);
/* number of CUDA devices visible; *n = 0 and GH_ECUDA when there is none */
int gh_device_count(int *n);
/* Measurement aid (bench.py's roofline denominator, not on the hot path): FP32 FMA throughput of
 * the current device measured with an FFMA-only kernel (8 independent chains per thread, every SM
 * full, no memory traffic), best of `repeats` launches, in TFLOP/s (2 flop per FFMA). */
int gh_fp32_fma_probe(int repeats, double *tflops);

/* ---- stateless force evaluation: the four _jbgrav entry points ---------------------------- */

/* a_i = sum_{j != i} m_j (x_j - x_i) / (|x_j - x_i|^2 + eps^2)^{3/2}      (_jbgrav.c:140-193) */
int gh_direct_summation(int prec, const double *pos, const double *mass, int64_t np, double eps,
                        double *acc_out, int mem, void *stream);

/* same at nf arbitrary target positions; a source exactly at a target with eps = 0 contributes
 * zero (_jbgrav.c:299-353, zero guard :327-330) */
int gh_direct_summation_position(int prec, const double *pos, const double *mass, int64_t np,
                                 const double *force_pos, int64_t nf, double eps,
                                 double *acc_out, int mem, void *stream);

/* Barnes-Hut octree, one particle per leaf, monopole, opening test size/|x - cell centre| < theta
 * (_jbgrav.c:487-541) on the octree the reference would build (_jbgrav.c:737-786). */
int gh_tree_force(int prec, const double *pos, const double *mass, int64_t np, double eps,
                  double theta, double *acc_out, int mem, void *stream);

int gh_tree_force_position(int prec, const double *pos, const double *mass, int64_t np,
                           const double *force_pos, int64_t nf, double eps, double theta,
                           double *acc_out, int mem, void *stream);

/* The stateless entry points keep grow-only device scratch per calling thread (so repeated calls
 * do not cudaMalloc); this frees the calling thread's scratch. */
int gh_release_thread_scratch(void);

/* Page-locked host memory for results and inputs of the GH_MEM_HOST entry points: copies to and
 * from such buffers are direct DMA transfers (a pageable destination costs a staged copy plus the
 * first-touch page faults of a fresh allocation -- 28 ms instead of 8.5 ms per tree evaluation
 * at N = 4M).  The Python binding hands out its result arrays from a recycling pool of these
 * (the reference returns a NEW array per call, _jbgrav.c:111,267,607,700; so does the binding). */
int gh_host_alloc(void **ptr, int64_t bytes);
int gh_host_free(void *ptr);

/* Statistics of the most recent tree evaluation made by the calling thread:
 * out[0] = tree entries (cells + leaves), out[1] = cells, out[2] = deepest cell level,
 * out[3] = accepted entries summed over targets, out[4] = visited entries summed over targets
 * out[5] = entries stepped through summed over warps (each warp scans the union of what its 32
 * targets need), out[6] = the largest such count of any warp, out[7] = number of warps
 * (out[3..6] only when the evaluation was made with gh_set_tree_stats(1)). */
int gh_tree_last_stats(int64_t out[8]);
int gh_set_tree_stats(int enable);

/* How GH_PREC_F32 tree evaluations walk the tree (process-wide; GH_PREC_F64 always walks per
 * target, which reproduces _jbgrav.c:502 node for node):
 *   GH_WALK_TARGET  every target applies the reference's opening test itself: the accepted node set
 *                   is exactly the reference's.
 *   GH_WALK_GROUP   (default) the 32 Morton-consecutive targets of a warp share one traversal; a
 *                   cell is accepted only if the reference's test passes for every point of the
 *                   targets' bounding box, so each target's accepted set is a refinement of the
 *                   reference's (never a coarser one).  Statistics in this mode: out[3] = list
 *                   entries evaluated summed over targets, out[4] = entries tested summed over
 *                   targets, out[5] = traversal iterations summed over warps, out[6] = warps that
 *                   fell back to the per-target walk.
 * The environment variable GH_TREE_WALK=target|group sets the initial mode. */
#define GH_WALK_TARGET 0
#define GH_WALK_GROUP 1
int gh_set_tree_walk(int mode);
int gh_get_tree_walk(void);

/* Opt-in accuracy upgrade BEYOND the reference (the reference is monopole only,
 * _jbgrav.c:522-524): every accepted cell also contributes its traceless quadrupole about its
 * centre of mass.  Off by default (GH_TREE_QUADRUPOLES=1 sets the initial value); with it on, tree
 * evaluations use the per-target walk -- the reference's octree and accepted node set -- in
 * either precision, and multi-GPU steps build the tree redundantly.  Mean force error at
 * theta = 0.7: ~3.7x below the monopole tree's (oracle.tree_force_quad is the CPU model). */
int gh_set_tree_quadrupoles(int enable);
int gh_get_tree_quadrupoles(void);
/* Hybrid rule of the group walk (off by default = 0; GH_WALK_HYBRID=<kappa> sets the initial
 * value): a target whose net acceleration is smaller than kappa times the summed magnitude of its
 * list's contributions (estimated from a 1/16 sample) is re-evaluated with the reference's own
 * per-target criterion.  The shared list makes the truncation errors of a group's 32 targets one
 * coherent vector, which matters only where the net force nearly cancels (the softened core of a
 * cusp); kappa = 0.1 re-evaluates ~0.1 % of the targets and brings the extreme tail of the
 * relative-error distribution back to the reference tree's (CPU model: oracle.tree_force_group).
 * In this mode out[6] of gh_tree_last_stats carries the number of re-evaluated targets in its
 * high 32 bits. */
int gh_set_tree_walk_hybrid(double kappa);
double gh_get_tree_walk_hybrid(void);

/* ---- device-side initial conditions (inputs of the path; gravhopper.py:1327-1734) ----------- */

/* Sample n particles of an equilibrium model on the GPU with a counter-based generator
 * (Philox4x32-10, one stream per particle; NOT the numpy stream of the host generators) and
 * centre positions and velocities on their means (force_centers, gravhopper.py:1768-1785).
 *   GH_IC_PLUMMER   params = {b [kpc], M [Msun]};          table: cumulative q-distribution -> q
 *   GH_IC_HERNQUIST params = {a [kpc], M [Msun], cutoff};  table: E -> cumulative f(E) (both ways)
 *   GH_IC_TSIS      params = {maxrad [kpc], M [Msun]};     no table
 * Tables are host pointers (a few hundred doubles, built by the caller exactly as the reference
 * builds them).  Outputs: pos (n,3) kpc, vel (n,3) km/s, mass (n) Msun, host or device per mem. */
#define GH_IC_PLUMMER 1
#define GH_IC_HERNQUIST 2
#define GH_IC_TSIS 3
int gh_ic_sample(int kind, int64_t n, const double *params, int nparams, const double *table_x,
                 const double *table_y, int ntable, uint64_t seed, double *pos, double *vel,
                 double *mass, int mem, void *stream);

/* Exponential disk (IC.expdisk, gravhopper.py:1611-1734) sampled on the GPU the same way.
 *   params4 = {sigma0 [Msun/kpc^2], Rd [kpc], z0 [kpc], sigma_R(Rd) [km/s]}
 * Four host tables of ntable entries on one radial grid table_R [kpc] (first entry 0): the cumulative
 * mass fraction table_cum (gravhopper.py:1675-1677), the mean rotation table_vphi = R Omega(R) [km/s]
 * and table_ratio = 4 Omega^2 / kappa^2, built by the caller from the disk's own Bessel-function
 * rotation curve plus any external one (:1700-1714); gravhopper_b200/ic_gpu.py builds them. */
int gh_ic_sample_expdisk(int64_t n, const double *params4, const double *table_R,
                         const double *table_cum, const double *table_vphi,
                         const double *table_ratio, int ntable, uint64_t seed, double *pos,
                         double *vel, double *mass, int mem, void *stream);

/* ---- device-resident leapfrog engine: Simulation.run ------------------------------------- */

typedef struct gh_engine gh_engine;

/* Create an engine on CUDA device `device` for a system of n_total particles of which this
 * engine integrates the targets [i_begin, i_begin + i_count) (single GPU: 0, n_total).
 * All ranks of a multi-GPU run hold the full source set; see gh_engine_bind_sources. */
int gh_engine_create(gh_engine **out, int device, int64_t n_total, int64_t i_begin,
                     int64_t i_count, int prec);
int gh_engine_destroy(gh_engine *e);

/* Upload the state at a snapshot: pos/vel of the OWNED targets (i_count,3) in kpc and km/s, and
 * the masses of ALL n_total particles in Msun (gravhopper.py:338-340).  Host pointers. */
int gh_engine_upload(gh_engine *e, const double *pos, const double *vel, const double *mass_all);

/* Same from DEVICE pointers (e.g. the outputs of gh_ic_sample with GH_MEM_DEVICE). */
int gh_engine_upload_device(gh_engine *e, const double *pos, const double *vel,
                            const double *mass_all);

/* Multi-GPU only: use two caller-owned device buffers (e.g. torch tensors) as the double-buffered
 * source array that an NCCL all-gather fills each step.  Layout per particle: 3 float64
 * (x_half) for GH_PREC_F64, or one float4 (x_half - origin, mass) for GH_PREC_F32.  After
 * upload, and after every gh_engine_step, the engine has written its own slice
 * [i_begin, i_begin+i_count) of buffer gh_engine_source_index(); the caller all-gathers that
 * buffer in place and then calls gh_engine_step. */
int gh_engine_bind_sources(gh_engine *e, void *buf0, void *buf1);
int gh_engine_source_index(gh_engine *e, int *idx);
int gh_engine_source_stride_bytes(gh_engine *e, int64_t *bytes);

/* Origin subtracted from positions before they are rounded to fp32 (GH_PREC_F32 engines only;
 * default 0,0,0).  Every rank of a sharded run must set the same value. */
int gh_engine_set_origin(gh_engine *e, const double origin[3]);
/* Velocity (km/s) the fp32 origin moves with: the system's mean velocity at upload.  Every step
 * advances the origin by vel * dt, so bulk motion costs no fp32 digits in long runs. */
int gh_engine_set_origin_velocity(gh_engine *e, const double vel[3]);

/* Build x_half = x + ((0.5 v) dt) K for the next step from the current state
 * (gravhopper.py:409) into the owned slice of source buffer gh_engine_source_index().  Single-GPU
 * engines do this implicitly; sharded engines call it once after upload (and whenever dt
 * changes), all-gather, then step. */
int gh_engine_prepare(gh_engine *e, double dt);

/* One DKD step (gravhopper.py:405-416): the sources hold x_half = x + 0.5 v dt; this computes
 * a(x_half) [+ ext_acc], v += a dt, x = x_half + 0.5 v dt, and x_half of the NEXT step into the
 * owned slice of the other source buffer (which becomes gh_engine_source_index()).
 * ext_acc (nullable): (i_count,3) float64, km/s/Myr, the summed external accelerations the
 * caller evaluated at x_half (gravhopper.py:455-457); ext_mem says where it lives.
 * Asynchronous on the engine's stream. */
int gh_engine_step(gh_engine *e, double dt, double eps, double theta, int algorithm,
                   const double *ext_acc, int ext_mem);

/* Device-native analytic external potentials, evaluated at x_half inside the step kernels
 * (replaces the Python callbacks of gravhopper.py:462-473 for these static fields).
 * Units kpc, Msun, km/s.  params[1..3] = centre.  Kinds:
 *   GH_POT_POINTMASS  params[0] = M, [4] = Plummer softening b        a = -G M d / (r^2+b^2)^{3/2}
 *   GH_POT_HERNQUIST  params[0] = M, [4] = a                          a = -G M d / (r (r+a)^2)
 *   GH_POT_NFW        params[0] = 4 pi rho0 rs^3, [4] = rs            a = -G M_s (ln(1+x) - x/(1+x)) d / r^3
 *   GH_POT_LOGHALO    params[0] = v0 [km/s], [4] = rc, [5] = q        a = -v0^2 (x,y,z/q^2)/(rc^2+x^2+y^2+z^2/q^2)
 *   GH_POT_MIYAMOTO   params[0] = M, [4] = a, [5] = b
 * At most 4 per engine. */
#define GH_POT_POINTMASS 1
#define GH_POT_HERNQUIST 2
#define GH_POT_NFW 3
#define GH_POT_LOGHALO 4
#define GH_POT_MIYAMOTO 5
int gh_engine_add_potential(gh_engine *e, int kind, const double *params, int nparams);
int gh_engine_clear_potentials(gh_engine *e);

/* Single-GPU convenience: run nsteps steps back to back on the device.  Every `snapshot_every`
 * steps (and always after the last one when snapshot_every > 0) the state is copied to
 * pos_hist/vel_hist (host, (nsnap, i_count, 3) float64, row k = k-th snapshot taken); the copies
 * overlap the following steps.  snapshot_every = 1 reproduces the reference's history arrays
 * (gravhopper.py:330-336).  pos_hist/vel_hist may be NULL when snapshot_every = 0.
 * Synchronous: returns when the last snapshot has landed. */
int gh_engine_run(gh_engine *e, int64_t nsteps, double dt, double eps, double theta,
                  int algorithm, int64_t snapshot_every, double *pos_hist, double *vel_hist);

/* Copy the owned state (i_count,3)+(i_count,3) to host; synchronous. */
int gh_engine_download(gh_engine *e, double *pos, double *vel);
/* Copy the owned half-drifted positions x_half (i_count,3) to host (external-force hook). */
int gh_engine_download_xhalf(gh_engine *e, double *xhalf);
/* Change dt between steps: x_half is rebuilt from (x, v) with the new dt (= gh_engine_prepare). */
int gh_engine_set_dt(gh_engine *e, double dt);
/* Kinetic and potential energy (Msun (km/s)^2) of the owned targets against all sources, with the
 * potential consistent with the softened force law; out = {KE, PE}.  Single GPU: the total. */
int gh_engine_energy(gh_engine *e, double eps, double out[2]);
int gh_engine_synchronize(gh_engine *e);
/* Device pointers of the owned state, for zero-copy views (torch.from_dlpack-style) */
int gh_engine_state_ptrs(gh_engine *e, double **pos_dev, double **vel_dev);
/* The engine's cudaStream_t (as void*), e.g. to order an NCCL all-gather against the steps. */
int gh_engine_stream(gh_engine *e, void **stream);
/* Tree statistics of the engine's last tree step (layout as gh_tree_last_stats). */
int gh_engine_tree_stats(gh_engine *e, int64_t out[8]);
/* 1 if the last gh_engine_run page-locked the caller's history arrays (snapshot copies overlap the
 * steps), 0 if they exceeded GH_PIN_HISTORY_MAX (default 4 GiB) or registration failed (staged copies). */
int gh_engine_history_pinned(gh_engine *e, int *pinned);
/* Kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int gh_engine_launch_count(gh_engine *e, int64_t *count);
/* Device time (ms, CUDA events on the engine's stream) of the force kernel of the last step. */
int gh_engine_last_force_ms(gh_engine *e, float *ms);
/* Mean device time (ms) of the force kernel over the last `last_k` steps (at most 64 are kept);
 * *count (nullable) = the number of steps averaged. */
int gh_engine_force_ms_mean(gh_engine *e, int last_k, float *mean_ms, int *count);

/* ---- engine groups: one system on several GPUs -------------------------------------
 * Replaces nothing in the reference (it is single threaded); this is the multi-GPU form of
 * Simulation.run (gravhopper.py:293-320, :405-416).  Targets are sharded evenly over `world` ranks
 * (rank r owns particles [begin_r, begin_r + count_r), the first n % world ranks one more); each
 * step all-gathers the half-drifted positions over NCCL; direct summation tiles all sources per
 * rank; the fp32 tree is built DISTRIBUTED (every rank sorts and emits one Morton key range, the
 * entry segments are all-gathered) and walked per rank; the fp64 tree is built redundantly.
 * NCCL is loaded at run time (libnccl.so.2, or $GH_NCCL_LIB).  A group is driven by one thread. */
typedef struct gh_group gh_group;
int gh_nccl_version(int *version);
/* All GPUs of one process: engines on `devices[0..ndev)` (NULL = 0..ndev-1), ncclCommInitAll. */
int gh_group_create_local(gh_group **g, int ndev, const int *devices, int64_t n_total, int prec);
/* One rank of a multi-process job: rank 0 calls gh_group_unique_id and hands the 128 bytes to
 * every rank (e.g. with torch.distributed / MPI); every rank then calls gh_group_create_rank. */
int gh_group_unique_id(void *id128);
int gh_group_create_rank(gh_group **g, const void *id128, int rank, int world, int device,
                         int64_t n_total, int prec);
int gh_group_destroy(gh_group *g);
int gh_group_size(gh_group *g, int *nlocal, int *world, int *rank0);
/* Local engine k (rank rank0 + k) and its particle range: upload / download / potentials go
 * through the gh_engine_* calls on it; stepping goes through the group. */
int gh_group_engine(gh_group *g, int k, gh_engine **e, int64_t *begin, int64_t *count);
int gh_group_prepare(gh_group *g, double dt);
/* nsteps DKD steps of the whole system (collectives included), asynchronous. */
int gh_group_step(gh_group *g, int64_t nsteps, double dt, double eps, double theta, int algorithm);
int gh_group_synchronize(gh_group *g);
/* 0: build the fp32 tree redundantly on every rank instead of distributed (default 1; GH_TREE_DIST). */
int gh_group_set_tree_distributed(gh_group *g, int enable);
/* Device time (ms) of the phases of the last step on local engine 0: [0] source all-gather,
 * [1] bbox + keys + select + sort, [2] boundary-key exchange, [3] levels + scans + moments,
 * [4] table exchange, [5] stitch + emit, [6] entry + sorted-index all-gathers, [7] walk,
 * [8] acceleration all-gather, [9] owners' kick and drift, [10] step. */
int gh_group_phase_ms(gh_group *g, float out[11]);

#ifdef __cplusplus
}
#endif
#endif

/*
 * ggp.h -- C ABI of libggp.so, the B200 (sm_100a) backend for the Strang-splitting time step of
 * GeneralizedGrossPitaevskii.jl.
 *
 * The reference has no FFI today: its backends are reached by Julia multiple dispatch on the array
 * type (src/strang_splitting.jl:62, src/misc.jl:7,55).  The seam this library sits behind is the
 * CommonSolve triple `init / step! / solve!` of StrangSplitting:
 *
 *   ggp_plan_create   replaces  init(prob, ::StrangSplitting, tspan; ...)   src/strang_splitting.jl:32-67
 *                     (device state, exp(-i dt D(k)) / exp(-i dt/2 V(r)) tables from
 *                      get_exponential src/misc.jl:12-20, pump buffers src/misc.jl:22-27,
 *                      FFT plans src/misc.jl:53-58)
 *   ggp_set_state     replaces  u = copy.(prob.u0)                         src/strang_splitting.jl:48
 *   ggp_step          replaces  the inner loop `for _ in 1:steps_per_save; t += dt; step!(iter,t,dt)`
 *                                                                          src/fixed_time_stepping.jl:43-47
 *                     i.e. nsteps x step!                                  src/strang_splitting.jl:86-90
 *                     = potential_pump_step! (:78-84) + diffusion_step! (:69-76) + potential_pump_step!
 *                     with muladd_kernel! (src/kernels.jl:37-54), perform_ft! (src/misc.jl:60-64),
 *                     sample_noise! (src/misc.jl:44-51), evaluate_pump! (src/misc.jl:29-42)
 *   ggp_get_state     replaces  map(copy!, slice, iter.u)                  src/fixed_time_stepping.jl:48
 *   ggp_observe       (no reference counterpart; on-device ensemble observables)
 *
 * Conventions
 *   - Arrays are Julia column-major: index (i1, ..., id, b) with i1 fastest, batch (trajectory) dims
 *     flattened into one trailing index b.  One array per field component (the reference's
 *     NTuple{M,Array}).
 *   - Every host pointer is BORROWED for the duration of the call only.
 *   - Every function returns 0 on success or a negative ggp_status; the message is available from
 *     ggp_last_error() (thread-local).  No C++ exception crosses this boundary.
 *   - A plan is not thread-safe; distinct plans are independent.  ggp_step is stream-ordered and may
 *     return before the GPU finishes; ggp_get_state / ggp_observe / ggp_synchronize / destroy block.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     GGP_ERR_CUDA.
 */
#ifndef GGP_H
#define GGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGP_ABI_VERSION 5u
#define GGP_MAX_COMPONENTS 4 /* ncomp <= 2: fused kernels; 3..4: generic plan (ABI 5) */

typedef enum ggp_status {
  GGP_OK = 0,
  GGP_ERR_INVALID = -1,     /* bad descriptor / argument */
  GGP_ERR_UNSUPPORTED = -2, /* legal in the reference, not covered by this backend (e.g. an axis longer than 8192) */
  GGP_ERR_CUDA = -3,        /* CUDA runtime failure (message in ggp_last_error) */
  GGP_ERR_NCCL = -4,
  GGP_ERR_ALLOC = -5
} ggp_status;

typedef enum ggp_precision { GGP_C64 = 0, GGP_C128 = 1 } ggp_precision;

/* kind of the exp tables: what `dispersion(k,param)` / `potential(r,param)` returned in Julia */
typedef enum ggp_table_kind {
  GGP_TABLE_NONE = 0,   /* AdditiveIdentity -> multiplicativeIdentity (src/misc.jl:12) */
  GGP_TABLE_SCALAR = 1, /* Number: one complex per point */
  GGP_TABLE_DIAG = 2,   /* SVector{M}: M complex per point (elementwise product, src/kernels.jl:10) */
  GGP_TABLE_FULL = 3,   /* SMatrix{M,M}: M*M complex per point, column-major (m11,m21,m12,m22) */
  GGP_TABLE_SEP_AXES = 4 /* disp_kind only (ABI 4): a scalar exp_D given as one factor per axis,
                            exp_D(k) = prod_a disp_axes[a][k_a] -- see disp_axes below */
} ggp_table_kind;

typedef enum ggp_nl_kind {
  GGP_NL_NONE = 0,
  GGP_NL_DIAG = 1, /* G_i(u) = c_i + sum_j g_ij |u_j|^2   (every nonlinearity in test/, examples/, docs/) */
  GGP_NL_MATRIX = 2 /* ABI 5: the closure returned an SMatrix, G_ij(u) = C_ij + sum_k g_ijk |u_k|^2; the half-step applies
                       the matrix exponential cis(-dt/2 G) (src/kernels.jl:22-25,44).  Coefficients in nl_c_ext / nl_g_ext. */
} ggp_nl_kind;

typedef enum ggp_pump_kind {
  GGP_PUMP_NONE = 0,
  GGP_PUMP_SEPARABLE = 1, /* F_i(r,t) = S_i(r) a(t); a static pump is a(t) = const */
  GGP_PUMP_DENSE = 2      /* any F(r,t): the caller evaluates the pump closure on the direct grid at every half-step
                             time, exactly as evaluate_pump! does (src/misc.jl:34-42, called from
                             src/strang_splitting.jl:81), and hands the profiles to ggp_step_dense (ABI 4).
                             pump_table = the profile at tspan[1] (primed at init, src/strang_splitting.jl:58). */
} ggp_pump_kind;

typedef enum ggp_noise_kind {
  GGP_NOISE_NONE = 0,
  GGP_NOISE_CONST = 1, /* eta_i(u,r) = const (examples/truncated_wigner.jl:96, test/windowed_ft.jl:27-29) */
  GGP_NOISE_FIELD = 2  /* eta_i(u,r) = P(r) * (eta_i + sum_j alpha_ij |u_j|): field-dependent and spatially varying
                          amplitudes of docs/src/stochastic_simulations.md:68-79 */
} ggp_noise_kind;

typedef enum ggp_observable {
  GGP_OBS_DENSITY = 0,  /* out[c][r] = sum over local trajectories |u_c(r)|^2        (double, M*nspatial) */
  GGP_OBS_MOMENTUM = 1, /* out[c][k] = sum over local trajectories |fft(u_c)(k)|^2/N^2 (double, M*nspatial) */
  GGP_OBS_NORM = 2,     /* out[c]    = sum over everything |u_c|^2                    (double, M) */
  GGP_OBS_G2_MOMENTUM = 3 /* 1-D ensembles: out[c][m][n] = sum over local trajectories |F_c(m)|^2 |F_c(n)|^2 with
                             F = fft(u_c)/N  (double, M*N*N): the trajectory sum inside `G2` of
                             examples/truncated_wigner.jl:143-154; the Wigner-ordering corrections of `f` (:139-141)
                             are a host-side combination of this matrix with n(k) = GGP_OBS_MOMENTUM */
} ggp_observable;

typedef struct ggp_desc {
  uint32_t abi_version; /* GGP_ABI_VERSION */
  uint32_t struct_size; /* sizeof(ggp_desc) as seen by the caller */

  int32_t ndim;         /* number of FFT'd dims, 1..3 (length(lengths)) */
  int32_t ncomp;        /* M = length(u0), 1..GGP_MAX_COMPONENTS */
  int64_t n[3];         /* spatial sizes, n[0] fastest.  Powers of two 2..8192 run on the fused kernels; any other
                           length up to 4096 runs on the generic plan (Bluestein, ABI 5) */
  int64_t nbatch;       /* trajectories held by THIS plan (product of trailing dims / shards) */
  int64_t batch_offset; /* global index of this plan's first trajectory (Philox counters are global,
                           so results do not depend on the sharding) */
  int32_t precision;    /* ggp_precision of the fields */
  int32_t table_precision; /* ggp_precision of disp_table / pot_table / pump_table as passed in */
  int32_t device;       /* CUDA ordinal, -1 = current device */
  int32_t reserved0;
  void *stream;         /* cudaStream_t to run on; NULL = the plan creates its own */

  double dt;            /* the RESOLVED step _dt of resolve_fixed_timestepping (src/fixed_time_stepping.jl:19-21) */

  /* exp_Ddt = cis(-dt*D(k)) on the reciprocal grid, exp_Vdt = cis(-dt/2*V(r)) on the direct grid,
     exactly as get_exponential builds them (src/strang_splitting.jl:53-54).  Array-of-structs as
     Julia stores Array{SVector/SMatrix}: point-major, then the static entries column-major. */
  int32_t disp_kind;    /* ggp_table_kind */
  int32_t pot_kind;     /* ggp_table_kind */
  const void *disp_table;
  const void *pot_table;

  /* nonlinearity, registered form; complex coefficients as (re, im) */
  int32_t nl_kind;      /* ggp_nl_kind */
  int32_t nl_scalar;    /* 1: the closure returned a Number (same G for every component; may be combined
                              with a FULL potential table), 0: SVector */
  double nl_c[2][2];    /* c_i          [i][re/im] */
  double nl_g[2][2][2]; /* g_ij         [i][j][re/im] */

  /* pump */
  int32_t pump_kind;    /* ggp_pump_kind */
  int32_t pump_ncomp;   /* 1: closure returned a Number (added to every component, src/kernels.jl:17), M: SVector */
  const void *pump_table; /* S(r): complex, point-major then component */
  double pump_amp0[2];  /* a(tspan[1]) -- the value primed by evaluate_pump! at init (src/strang_splitting.jl:58) */

  /* position noise: noise = -i*sqrt(dt/2) * eta_i * xi   (src/kernels.jl:42) */
  int32_t noise_kind;   /* ggp_noise_kind */
  int32_t noise_real;   /* 1: noise_prototype is a real array (xi ~ N(0,1)); 0: complex (<|xi|^2> = 1) */
  double noise_eta[2][2]; /* eta_i [i][re/im]; if the closure returned a Number, repeat it */
  uint64_t seed;        /* Philox4x32-10 key */

  /* 3-D slab decomposition over slab_nranks processes (one per GPU), 0 or 1 = off.  n[] stays the GLOBAL
     grid; this rank's state is the z-slab n[0] x n[1] x (n[2]/slab_nranks) starting at plane
     slab_rank*n[2]/slab_nranks; pot_table / pump_table cover that z-slab; disp_table covers the y-slab
     n[0] x (n[1]/slab_nranks) x n[2] starting at row slab_rank*n[1]/slab_nranks (the layout after the
     all-to-all transpose, in which the z lines are local).  Needs ggp_comm_init before ggp_step. */
  int32_t slab_nranks;
  int32_t slab_rank;

  /* GGP_NOISE_FIELD only (ABI 3).  noise_alpha: coefficients of |u_j| (evaluated on the PRE-update field like
     G, src/kernels.jl:40-42).  noise_profile: NULL, or n[0] complex doubles P(point(k1)), k1 = 0..n[0]-1 -- the
     reference builds `point` by indexing EVERY grid axis with the first index K[1] (src/kernels.jl:27,41; SURVEY
     quirk Q2), so the spatial profile it applies is a function of the first (fastest) index only. */
  double noise_alpha[2][2][2]; /* alpha_ij [i][j][re/im] */
  const void *noise_profile;

  /* Separable dispersion (ABI 3).  A scalar exp_D that factorises as Dperp(k_1..k_{d-1}) * Dline(k_d) -- any
     dispersion that is a sum over axes, e.g. |k|^2/2 -- is held as those two factors and the full table is never
     read by the kernels.  The library checks the factorisation numerically on disp_table and accepts it if the
     largest deviation is <= tol * max|exp_D|, tol = 1e-13 (ComplexF64 plans) or 4e-6 (ComplexF32 plans).  A
     ComplexF32 problem's table is cis(-dt * fl32(D(k))): it deviates from ANY product by its own rounding,
     eps32 * |phase|, which exceeds 4e-6 on large grids (4096^2 with L = 64: 5e-6).  A host that has verified in
     Float64 that the dispersion is a sum over axes passes the measured deviation here so that the fast path is
     kept; 0 = library default. */
  double disp_sep_tol;

  /* GGP_TABLE_SEP_AXES (ABI 4): per-axis factors of a scalar exp_D, disp_axes[a] = n[a] complex numbers of
     table_precision with exp_D(k) = prod_a disp_axes[a][k_a]; disp_table is ignored.  For dispersions that are a sum
     over axes (|k|^2/2 + const ...) this replaces the full-grid table -- 16 GiB of host memory for a 1024^3
     ComplexF64 table -- by d short vectors.  Slab plans pass the GLOBAL axes; the library slices. */
  const void *disp_axes[3];

  /* Mixed-precision tables (ABI 4).  In the reference the table element type follows the user's
     types (src/misc.jl:14-17): a ComplexF32 problem stepped with a Float64 dt (or Float64 lengths) holds ComplexF64
     tables and multiplies in ComplexF64 before rounding to ComplexF32 on store (src/kernels.jl:51-53).  1 = this is
     such a problem: a ComplexF32 plan then keeps the two factors of a separable exp_D in Float64 and forms
     exp_D[k] * u~[k] in Float64 (one rounding, as the reference); full (non-separable) tables, exp_V and the
     nonlinear phase stay in the plan's precision (documented deviation).  0 = tables in plan precision. */
  int32_t mixed_precision_tables;
  int32_t reserved1;

  /* ABI 5: coefficient arrays for more than two components and for matrix-valued nonlinearities (both run on the
     generic plan: one kernel per stage of the reference's sequence instead of the fused kernels).  REQUIRED when
     ncomp > 2 or nl_kind == GGP_NL_MATRIX (the fixed-size arrays above are then ignored), optional otherwise (NULL).
       nl_c_ext         GGP_NL_DIAG: ncomp x (re,im) c_i (nl_scalar: 1 entry)      GGP_NL_MATRIX: ncomp^2 x (re,im) C_ij, [i][j]
       nl_g_ext         GGP_NL_DIAG: ncomp^2 x (re,im) g_ij, [i][j] (nl_scalar: ncomp entries g_j)
                                                                                   GGP_NL_MATRIX: ncomp^3 x (re,im) g_ijk, [i][j][k]
       noise_eta_ext    ncomp x (re,im)
       noise_alpha_ext  ncomp^2 x (re,im), [i][j]   (GGP_NOISE_FIELD) */
  const double *nl_c_ext;
  const double *nl_g_ext;
  const double *noise_eta_ext;
  const double *noise_alpha_ext;
} ggp_desc;

typedef struct ggp_plan ggp_plan;

int ggp_version(void);
int ggp_device_count(void);
const char *ggp_last_error(void);

int ggp_plan_create(const ggp_desc *desc, ggp_plan **out);
int ggp_plan_destroy(ggp_plan *plan);

/* u_host: M pointers to host arrays of nspatial*nbatch complex numbers of the plan's precision */
int ggp_set_state(ggp_plan *plan, const void *const *u_host);
int ggp_get_state(ggp_plan *plan, void *const *u_host);

/*
 * Advance nsteps Strang steps.
 *   pump_amp  : NULL (static pump: a(t) = pump_amp0), or 2*nsteps complex doubles (re,im): for step s the
 *               values a(t_s + dt/2), a(t_s + dt) where t_s is the reference's already-incremented t
 *               (src/fixed_time_stepping.jl:44-45, src/strang_splitting.jl:87,89).
 *               The library keeps the previous value (F_now) across calls.
 *   noise_host: NULL => in-kernel Philox stream; otherwise TEST MODE: 2*nsteps*M pointers to host arrays
 *               (same shape as the state; complex of the plan's precision, or real if noise_real) in the
 *               reference's draw order: step, half-step, component (src/misc.jl:44-51).
 */
int ggp_step(ggp_plan *plan, int64_t nsteps, const double *pump_amp, const void *const *noise_host);

/*
 * GGP_PUMP_DENSE plans: as ggp_step, but instead of amplitudes the caller passes the pump evaluated on the direct
 * grid: pump_profiles = 2*nsteps pointers, for step s the profiles at t_s + dt/2 and t_s + dt (the reference's
 * one-dt-late times), each nspatial * pump_ncomp complex numbers of table_precision laid out like
 * pump_table (point-major, then component).  The library keeps F_now across calls (src/misc.jl:39-42: the
 * double-buffer shuffle of evaluate_pump!).  Replaces evaluate_pump! + the pump buffers of the reference for pumps
 * that do not separate as S(r) a(t).
 */
int ggp_step_dense(ggp_plan *plan, int64_t nsteps, const void *const *pump_profiles, const void *const *noise_host);

int ggp_synchronize(ggp_plan *plan);

/*
 * Streaming save: the non-blocking form of ggp_get_state for the snapshot copy of solve!'s
 * save loop, `map(copy!, slice, iter.u)` (src/fixed_time_stepping.jl:48).  The state is snapshotted on the
 * device in stream order (after every step issued so far) and transferred to u_host on a second stream while
 * the caller already issues the next ggp_step.  u_host (M pointers, as ggp_get_state) must stay valid -- and
 * should be page-locked (ggp_host_alloc) for the transfer to be asynchronous -- until ggp_save_wait returns.
 * A further ggp_save_async waits on the device for the previous transfer; ggp_save_wait blocks the host until
 * every transfer issued so far has landed.
 */
int ggp_save_async(ggp_plan *plan, void *const *u_host);
int ggp_save_wait(ggp_plan *plan);

/*
 * Checkpoint / resume (the reference has none).  The blob holds the fields plus the
 * plan state that is not in the descriptor -- half-step counter (Philox counter word) and the pump amplitude
 * of F_now (quirk Q1) -- so that  create(desc) + ggp_checkpoint_load + ggp_step(n2)  continues a run
 * bit-identically to the uninterrupted  ggp_step(n1); ggp_step(n2).  Loading into a plan of another shape,
 * precision or trajectory shard fails with GGP_ERR_INVALID.  The time t, the position in the pump schedule
 * and ts stay with the caller (they are Julia-side state).
 */
int64_t ggp_checkpoint_bytes(ggp_plan *plan);
int ggp_checkpoint_save(ggp_plan *plan, void *blob, uint64_t capacity);
int ggp_checkpoint_load(ggp_plan *plan, const void *blob, uint64_t size);

/* Ensemble observables summed over this plan's trajectories, written as doubles to out_host.
   If a communicator was attached with ggp_comm_init the sums are all-reduced over ranks (NCCL). */
int ggp_observe(ggp_plan *plan, int kind, double *out_host);

/* Windowed first-order coherence of a 1-D ensemble (test/windowed_ft.jl:31-49, `correlation`) without downloading
   the ensemble:  out[c][i][j] (re, im) = sum over local trajectories conj(F2[j]) * F1[i],
   F_a = ifftshift(fft(fftshift(u_c .* w_a)))  for the two windows w1, w2 (N complex doubles each, values on the
   direct grid).  The caller divides by length(sol) = N * ntraj.  All-reduced over ranks like ggp_observe.
   out_host: ncomp * N * N complex doubles. */
int ggp_observe_windowed(ggp_plan *plan, const double *w1, const double *w2, double *out_host);

/* Multi-GPU (one process per GPU).  unique_id: the 128-byte ncclUniqueId produced by
   ggp_comm_unique_id on rank 0 and broadcast by the caller's own plumbing. */
int ggp_comm_unique_id(void *unique_id_128);
int ggp_comm_init(ggp_plan *plan, int nranks, int rank, const void *unique_id_128);

/* Slab decomposition without NCCL on the data path: the all-to-all transposes are fused into the strided FFT
   kernels, which store their results straight into the slabs of the ranks that own them afterwards (peer
   memory over NVLink, mapped with CUDA IPC) -- no pack, send/receive or unpack sweeps.  Every rank exports a
   blob of GGP_IPC_BLOB_BYTES, the caller gathers them (rank order, its own plumbing) and hands the
   concatenation to every rank; a host-side barrier must follow before the first ggp_step.  Replaces the
   ncclSend/ncclRecv transposes that a plan uses when only ggp_comm_init was called (GGP_SLAB_NCCL=1 keeps
   those).  New in this library: the reference has no multi-GPU path. */
#define GGP_IPC_BLOB_BYTES 512
int ggp_slab_ipc_export(ggp_plan *plan, void *blob);
int ggp_slab_ipc_attach(ggp_plan *plan, const void *blobs_in_rank_order);

/* Harness helpers (bench / tests): device pointers, device-side timing on the plan's stream,
   pinned host memory, and the number of kernels launched so far. */
void *ggp_state_device_ptr(ggp_plan *plan, int comp);
int ggp_timer_begin(ggp_plan *plan);
int ggp_timer_end(ggp_plan *plan, float *milliseconds);
int64_t ggp_launch_count(ggp_plan *plan);
void *ggp_host_alloc(uint64_t bytes);
int ggp_host_free(void *p);
/* Page-lock / unlock memory the CALLER owns (e.g. the Julia `result` arrays allocated by init,
   src/strang_splitting.jl:41-43), so that ggp_save_async into it is a true asynchronous DMA (ABI 4). */
int ggp_host_register(void *p, uint64_t bytes);
int ggp_host_unregister(void *p);
/* bytes of device memory the plan owns */
int64_t ggp_device_bytes(ggp_plan *plan);
/* Per-kernel-class device timing inside ggp_step (CUDA events around every launch on the plan's
   stream).  Classes: 0 = contiguous-axis kernel (inverse FFT_x + real-space half-steps + forward
   FFT_x), 1 = strided kernel with the dispersion multiply, 2 = strided forward-only / inverse-only
   (3-D middle axis), 3 = 1-D whole-step kernel.  ms_total / launches: arrays of 4. */
int ggp_profile_enable(ggp_plan *plan, int on);
/* Measurement aid: when bytes > 0, a buffer of that size is overwritten after EVERY kernel of ggp_step so
   that each kernel starts with a cold L2 (timing rule for working sets smaller than the 126 MB L2).
   The flushes are outside the per-kernel events of ggp_profile_read.  bytes = 0 switches it off. */
int ggp_debug_l2_flush(ggp_plan *plan, uint64_t bytes);
/* Time `count` of those flushes alone (same stream, one event pair): subtracted from a bracketed flushed run. */
int ggp_debug_flush_only(ggp_plan *plan, int64_t count, float *milliseconds);
int ggp_profile_read(ggp_plan *plan, double *ms_total, int64_t *launches);
/* Step-level timing (bench.py's headline): one CUDA-event pair around the kernels of every steady-state step --
   the strided pass(es) of step s and the contiguous-axis kernel that closes it (inverse FFT_x, trailing V/2 of step s,
   leading V/2 of step s+1, forward FFT_x) -- so programmatic-launch overlap INSIDE a step is part of the number, and,
   when flush_bytes > 0, that many bytes are overwritten BETWEEN the windows (outside every event pair) so that each
   step starts with a cold L2.  The leading kernel of a ggp_step call (V/2 + forward FFT_x only) is outside the
   windows.  ggp_profile_steps_read: total milliseconds over the windows and their number since enable. */
int ggp_profile_steps_enable(ggp_plan *plan, int on, uint64_t flush_bytes);
int ggp_profile_steps_read(ggp_plan *plan, double *ms_total, int64_t *windows);

#ifdef __cplusplus
}
#endif
#endif /* GGP_H */

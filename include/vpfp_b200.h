/* vpfp_b200.h -- C ABI of the B200-native VPFP phase-space update (libvpfp_b200.so).
 *
 * Drop-in boundary for the per-timestep hot path of VlaPy (pure Python reference; paths below
 * are relative to its source tree).  One entry point per reference operator; a maintainer binds
 * them with ctypes from the reference's factory functions.
 *
 * Conventions
 *  - every pointer is DEVICE memory owned by the caller (torch tensors: tensor.data_ptr());
 *    f is fp64, row-major (batch*nx rows, nv contiguous) with row stride `ld` doubles (ld >= nv);
 *  - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never synchronise;
 *  - return 0 on success, non-zero on error (no exceptions cross the ABI); vpfp_last_error()
 *    gives the message; VPFP_ERR_UNSUPPORTED maps to Python NotImplementedError;
 *  - the library keeps only twiddle tables and reduction scratch (cached per device, freed by
 *    vpfp_shutdown()); there is no CPU fallback anywhere.
 */
#ifndef VPFP_B200_H
#define VPFP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VPFP_ABI_VERSION 1

#define VPFP_OK 0
#define VPFP_ERR_ARG 1          /* bad argument                                   */
#define VPFP_ERR_UNSUPPORTED 2  /* size/flavour not implemented (NotImplementedError) */
#define VPFP_ERR_CUDA 3         /* a CUDA runtime call failed                      */

/* flags for the advection operators */
#define VPFP_PHASE_EXACT 0   /* per-bin sincos of theta = (k*dt)*c, the reference's own rounding */
#define VPFP_PHASE_TABLE 1   /* geometric phase tables (needs a uniform fftfreq wavenumber grid);
                              * honoured by the register-resident kernels (256 <= N <= 16384),
                              * the generic kernels always use the exact phases */
#define VPFP_FORCE_GENERIC 2 /* testing: skip the register-resident kernels */
#define VPFP_FORCE_THREE_PASS 4 /* testing / A-B timing: skip the single-pass row kernel */

/* collision operator ids (vlapy/core/collisions.py:292-317) */
#define VPFP_FP_LB 0
#define VPFP_FP_DG 1

int vpfp_abi_version(void);
const char *vpfp_last_error(void);
int vpfp_shutdown(void);

/* Number of kernels this library has enqueued since the last reset (all devices, all streams; kernels replayed
 * by a CUDA graph are counted once, at capture).  bench.py reports it as gpu_launches. */
long vpfp_launch_count(int reset);

/* Counts reallocations of the library's reduction scratch (x-mode partials, fused density partials).  Outgrown
 * blocks are retired, never freed before vpfp_shutdown, so a CUDA graph captured earlier stays valid; callers
 * that want their graphs to use the current block re-capture when this number has changed.  The scratch is
 * shared by the streams of a device: calls that use it must not run concurrently on two streams. */
unsigned long vpfp_scratch_generation(void);

/* Per-launch timing for benchmarking: when enabled, every kernel launch of this library is
 * bracketed by CUDA events on its stream.  vpfp_profile_report() synchronises the device, writes
 * one "label count total_ms" line per kernel label into buf and clears the records. */
int vpfp_profile_enable(int on);
int vpfp_profile_report(char *buf, int buflen);

/* e df/dv, exponential integrator: f_out = Re ifft_v( exp(-i kv dt e[x]) fft_v f_in ).
 * Replaces vlapy/core/vlasov.py:113-140 (step_edfdv_exponential).
 * f_in/f_out: (rows, nv) with row strides ld_in/ld_out; e: (rows); kv: (nv) = 2*pi*fftfreq.
 * rows = batch*nx (the operator is independent per row). nv must be a power of two >= 4. */
int vpfp_edfdv_exp(const double *f_in, long ld_in, double *f_out, long ld_out, const double *e,
                   const double *kv, double dt, int rows, int nv, int flags, void *stream);

/* v df/dx, exponential integrator: f_out = Re ifft_x( exp(-i kx dt v) fft_x f_in ).
 * Replaces vlapy/core/vlasov.py:83-110 (step_vdfdx_exponential).
 * f: (batch, nx, ncols) row stride ld; kx: (batch, nx) per-simulation wavenumbers; v: (ncols).
 * ncols may be a v-slice (multi-GPU v-sharded layout); it must be even, nx a power of two. */
int vpfp_vdfdx_exp(const double *f_in, long ld_in, double *f_out, long ld_out, const double *kx,
                   const double *v, double dt, int batch, int nx, int ncols, int flags,
                   void *stream);

/* v df/dx fused with the charge density of its result: n_out[b*nx + x] = trapz_v f_out[b, x, :]
 * (vlapy/core/field.py:27-36), reduced in the epilogue of the last pass instead of a separate
 * read of f.  Every field solve of the reference follows a v df/dx
 * (vlapy/core/vlasov_poisson.py:54-55, 115-116, 205-206).  edge_flags as in vpfp_moments. */
int vpfp_vdfdx_exp_density(const double *f_in, long ld_in, double *f_out, long ld_out,
                           const double *kx, const double *v, double dt, int batch, int nx,
                           int ncols, int flags, double *n_out, double dv, int edge_flags,
                           void *stream);

/* e df/dv, 2nd-order centred differences: f_out = f - e * gradient_v(f) * dt (edge_order=2).
 * Replaces vlapy/core/vlasov.py:143-165. */
int vpfp_edfdv_cd2(const double *f_in, long ld_in, double *f_out, long ld_out, const double *e,
                   double dt, double dv, int rows, int nv, void *stream);

/* Semi-Lagrangian advection (backward characteristics + cubic-spline interpolation).
 * vpfp_vdfdx_sl replaces vlapy/core/vlasov.py:42-80 (get_vdfdx_sl): f_out[i, j] = S_j(x_i - v_j dt) with S_j the
 *   not-a-knot cubic spline of column j padded with one periodic ghost row on either side; feet outside the padded
 *   axis are clamped to its ends (what scipy's RectBivariateSpline / FITPACK does for the reference).
 * vpfp_edfdv_sl replaces vlapy/core/vlasov.py:168-210 (get_edfdv_sl): f_out[i, j] = S_i(v_j - e_i dt) along v.
 * x: (nx), v: (nv) the axes (uniform, spacing dx / dv = ax[2] - ax[1] as the reference takes it); e: (nx).
 * nx >= 4; 10 <= nv <= 16386 for the v direction.  Uses library scratch of (nx + 2) * nv doubles. */
int vpfp_vdfdx_sl(const double *f_in, long ld_in, double *f_out, long ld_out, const double *x, const double *v,
                  double dt, double dx, int nx, int nv, void *stream);
int vpfp_edfdv_sl(const double *f_in, long ld_in, double *f_out, long ld_out, const double *e, const double *v,
                  double dt, double dv, int nx, int nv, void *stream);

/* v-moments per row: out[k*out_ld + row], k = 0..nmom-1:
 *   k<6: trapz_v(f v^k)   (n, j, T, q, fv4, vN; vlapy/core/step.py:164-171, field.py:27-36)
 *   k=6: trapz_v(f^2), k=7: trapz_v(f ln f)   (step.py:216-224; NaN where f <= 0 like numpy)
 * v0/v1 give trapz end weights: the first/last LOCAL column gets weight dv/2 only when it is the
 * global first/last velocity cell (flags bit0 = has global first, bit1 = has global last). */
int vpfp_moments(const double *f, long ld, const double *v, double dv, double *out, long out_ld,
                 int nmom, int rows, int ncols, int edge_flags, void *stream);

/* Spectral Poisson: e = driver + Re ifft( i one_over_kx fft(1 - n) ).
 * Replaces vlapy/core/field.py:39-88.  n, driver, e: (batch, nx); one_over_kx: (batch, nx).
 * driver may be NULL.  Any nx >= 2: powers of two use an FFT (one launch for nx >= 256: the e df/dv kernels of that length
 * in Poisson mode, two density rows per packed sequence up to nx = 2048), other lengths a direct DFT. */
int vpfp_poisson(const double *n, const double *one_over_kx, const double *driver, double *e,
                 int batch, int nx, void *stream);

/* Implicit Fokker-Planck step, one tridiagonal system in v per row:
 * moments of f_in -> LB or Dougherty diagonals -> solve (A f_out = f_in).
 * Replaces vlapy/core/collisions.py:26-160 + 222-265 via step.py:70-113.
 * v: (nv). moments_out (nullable): (8, rows) moments of f_out laid out like vpfp_moments. */
int vpfp_fp_step(const double *f_in, long ld_in, double *f_out, long ld_out, const double *v,
                 double nu, double dt, double dv, int op, double *moments_out, long mom_ld,
                 int rows, int nv, void *stream);

/* Same operator for velocity grids built by np.linspace (vlapy/initializers.py:66):
 * v_i = v0 + i*vstep for i < nv-1 and v_{nv-1} = vlast, bit for bit.  The diagonals are then affine
 * in the cell index and nothing but f is loaded.  nv must be a power of two in [128, 16384];
 * other sizes return VPFP_ERR_UNSUPPORTED (callers fall back to vpfp_fp_step). */
int vpfp_fp_step_linspace(const double *f_in, long ld_in, double *f_out, long ld_out, double v0,
                          double vstep, double vlast, double nu, double dt, double dv, int op,
                          double *moments_out, long mom_ld, int rows, int nv, void *stream);

/* The reference's explicit two-stage collision interface (the product step fuses both stages and never
 * materialises the diagonals; these serve callers of get_batched_array_maker / get_matrix_solver).
 * vpfp_fp_diagonals: f -> (a, b, c), replaces vlapy/core/collisions.py:26-83 (lb) and :86-160 (dg) as returned by
 *   get_batched_array_maker (:292-317).  a, c: (rows, nv-1) with pitches lda, ldc; b: (rows, nv).
 * vpfp_tridiag_solve: x = solve(tridiag(a, b, c), d) per row, general diagonals, no pivoting; replaces
 *   vlapy/core/collisions.py:222-265 (_batched_tridiag_solver_ behind get_matrix_solver :268-289).
 *   a[i-1] couples row i to x[i-1], c[i] couples row i to x[i+1] (the reference's storage). 8 <= nv <= 16384.
 *   x may alias d.  A pitch of 0 for a, b or c broadcasts one set of diagonals to every row. */
int vpfp_fp_diagonals(const double *f, long ld, const double *v, double nu, double dt, double dv, int op,
                      double *a, long lda, double *b, long ldb, double *c, long ldc, int rows, int nv,
                      void *stream);
int vpfp_tridiag_solve(const double *a, long lda, const double *b, long ldb, const double *c, long ldc,
                       const double *d, long ldd, double *x, long ldx, int rows, int nv, void *stream);

/* First nmodes x-Fourier modes of f per v: out[(b*nmodes + m)*ncols + j] = sum_x f[b,x,j] w^(m x)
 * as interleaved (re, im) doubles.  Replaces vlapy/core/step.py:130-135 (get_f_to_store). */
int vpfp_xmodes(const double *f, long ld, double *out, int nmodes, int batch, int nx, int ncols,
                void *stream);

/* The same sums over a slab of rows x_offset .. x_offset+nx-1 of a grid of nx_total cells
 * (x-sharded multi-GPU layout); the partial results of all slabs add up to vpfp_xmodes. */
int vpfp_xmodes_partial(const double *f, long ld, double *out, int nmodes, int batch, int nx,
                        int ncols, int x_offset, int nx_total, void *stream);

/* Ponderomotive driver E_d(x, t) summed over npulse pulses (vlapy/field_driver.py:24-50).
 * pulses: host array of 7 doubles per pulse {k0, w0, a0, t_L, t_R, t_wL, t_wR}. x, out: device. */
int vpfp_driver(const double *x, double t, const double *pulses, int npulse, double *out, int nx,
                void *stream);

/* The same driver with the time read on the device: t = (((*t_dev + incs[0]) + incs[1]) + ...),
 * summed in the reference's order (vlapy/core/vlasov_poisson.py:116-148).  Lets a whole timestep
 * be captured in a CUDA graph and replayed with a new time.  incs: host array, ninc <= 6. */
int vpfp_driver_dev(const double *x, const double *t_dev, const double *incs, int ninc,
                    const double *pulses, int npulse, double *out, int nx, void *stream);

/* Series reductions of one stored step (vlapy/core/step.py:202-224): out[0..6] =
 * mean_x of moments rows n, j, T, mean(e^2), mean(de^2), mean_x of rows f2, flogf. */
int vpfp_series(const double *moments, long mom_ld, const double *e, const double *de,
                double *out, int nx, void *stream);

/* Multi-GPU (one process per GPU): shards that peers can address, and the two advections with the
 * x<->v layout change fused into the stores of their last pass -- the result is written over
 * NVLink straight into the target ranks' shards instead of a local array + all-to-all.
 *   vpfp_ipc_alloc / open / close / free: cudaMalloc + CUDA IPC handle (64 bytes) exchange.
 *   vpfp_edfdv_exp_scatter: f_in is the local x-shard (rows, nv); scratch (rows, nv) holds the
 *     intermediate passes; the result column block q lands in peer_fv[q] (the v-shard of rank q,
 *     shape (rows*nparts, nv/nparts)) at rows [my_rank*rows, (my_rank+1)*rows).
 *   vpfp_vdfdx_exp_scatter: f_in is the local v-shard (nx, ncols); the result row block q lands in
 *     peer_fx[q] (the x-shard of rank q, shape (nx/nparts, ncols*nparts)) at columns
 *     [my_rank*ncols, (my_rank+1)*ncols); n_out (nullable) receives the partial charge density.
 * The caller orders the ranks (a barrier on the stream) before a shard is read.  Both need the
 * register-resident kernels (256 <= N <= 16384 and >= 4M cells per rank). */
int vpfp_ipc_alloc(unsigned long bytes, void **ptr, unsigned char *handle64);
int vpfp_ipc_open(const unsigned char *handle64, void **ptr);
int vpfp_ipc_close(void *ptr);
int vpfp_ipc_free(void *ptr);
int vpfp_edfdv_exp_scatter(const double *f_in, long ld_in, double *scratch, long ld_scratch,
                           const double *e, const double *kv, double dt, int rows, int nv, int flags,
                           void *const *peer_fv, int nparts, int my_rank, void *stream);
int vpfp_vdfdx_exp_scatter(const double *f_in, long ld_in, double *scratch, long ld_scratch,
                           const double *kx, const double *v, double dt, int nx, int ncols,
                           int flags, double *n_out, double dv, int edge_flags,
                           void *const *peer_fx, int nparts, int my_rank, void *stream);

/* Ensembles of independent simulations (BASELINE config 4): simulation b owns rows
 * b*nx .. b*nx+nx-1 of every per-x array.  vpfp_series_batch writes out[b*7 + k];
 * vpfp_driver_batch evaluates the driver with per-simulation grids x[batch][nx] and pulse
 * parameters pulses_dev[batch][npulse][7] (device memory); t_dev may be NULL (then t is used). */
int vpfp_series_batch(const double *moments, long mom_ld, const double *e, const double *de,
                      double *out, int nx, int batch, void *stream);
int vpfp_driver_batch(const double *x, double t, const double *t_dev, const double *incs, int ninc,
                      const double *pulses_dev, int npulse, double *out, int nx, int batch,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VPFP_B200_H */

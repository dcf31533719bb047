/*
 * gpmap_b200.h -- C-ABI of libgpmap_b200.so: the B200-native (sm_100a) GP path-modelling hot path.
 *
 * Boundary note.  The reference (/root/reference/GPmap.py, 219 lines of Python) has NO plugin,
 * operator or FFI interface and NO Gaussian-process code: its only
 * linear-algebra call is np.linalg.norm at GPmap.py:120.  The entry points below are therefore
 * new API; what each one "replaces" is the numpy/scipy call a GPmap.py-style script would make
 * for that step (the libraries GPmap.py:1,10 imports), and the data conventions are the
 * reference's own:
 *   - per-path samples are float64 arrays xs, ys[, timestamp]          (GPmap.py:15-23)
 *   - a batch is many equal-length paths                                 (GPmap.py:30-34,96,117)
 *   - coordinates live in [-5e4, 5e4]^2                                  (GPmap.py:126,155)
 *
 * Conventions
 *   - every matrix is float64, ROW-major, with an explicit leading dimension in elements;
 *     leading dimensions must be even (16-byte row pitch, a TMA requirement) and base pointers
 *     16-byte aligned.  torch.float64 CUDA tensors satisfy this.
 *   - all data pointers are DEVICE pointers owned by the caller; the library allocates nothing
 *     persistent except what gpm_create() puts in the handle (one helper stream, events).
 *   - theta = [l_1 .. l_D, signal_var, noise_var] is a HOST pointer to D+2 doubles (D = 2 or 3),
 *     k(x,z) = signal_var * exp(-0.5 * sum_d ((x_d - z_d)/l_d)^2).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 ok; -i = argument i (1-based) invalid; > 0 = cudaError_t of a failed launch
 *     or API call (text via gpm_last_error()).  Numerical failure is LAPACK-style: `info` is a
 *     device int32 the kernels set to j > 0 when pivot j (1-based) is not positive.
 *   - no CPU fallback exists: without a CUDA device every call fails with a CUDA error.
 */
#ifndef GPMAP_B200_H
#define GPMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPM_VERSION 100          /* 0.1.0 */
#define GPM_NB 128               /* factorisation block size; workspace layouts depend on it */

typedef struct gpm_handle_s* gpm_handle_t;
typedef void* gpm_stream_t;      /* cudaStream_t */

/* Regular query grid over the reference's plot window (GPmap.py:126): point m = gy*Gx + gx,
 * x = linspace(x0,x1,Gx)[gx], y = linspace(y0,y1,Gy)[gy]; t is the constant third coordinate
 * used when the model has D = 3. */
typedef struct gpm_grid {
  double x0, x1, y0, y1, t;
  int32_t gx, gy;
} gpm_grid_t;

/* flags for gpm_cov */
#define GPM_COV_FULL  0          /* write all N*N entries (8 N^2 bytes) */
#define GPM_COV_LOWER 1          /* write only the 128x128 tiles on or below the diagonal */
/* flags for gpm_predict */
#define GPM_PREDICT_MEAN      1
#define GPM_PREDICT_VAR       2
#define GPM_PREDICT_ADD_NOISE 4  /* var += noise_var (predictive variance of y rather than f) */

int         gpm_version(void);
const char* gpm_last_error(void);                     /* thread-local text of the last failure */
/* A handle carries mutable per-call state (a helper stream, an event pool, the flag arrays of the chained
 * solves): it is SINGLE-STREAM and SINGLE-THREAD.  Calls that may overlap in time -- issued on different CUDA
 * streams or from different host threads -- need one handle each (the Python binding keys its handles on
 * (device, stream)).  Every entry point runs on the handle's device and restores the caller's current device. */
int         gpm_create(gpm_handle_t* handle, int device);
int         gpm_destroy(gpm_handle_t handle);
int         gpm_sm_count(gpm_handle_t handle);
long long   gpm_launch_count(void);                   /* kernels launched by this library so far (process-wide) */
/* Debug / comparison switches of a handle (e.g. "no_lookahead", "no_separable", "var_steps", "no_path_fused";
 * the full list is kOptions in csrc/api.cu).  Each is initialised ONCE, at gpm_create, from the environment
 * variable GPM_<NAME>; no entry point reads the environment afterwards.  Returns -2 for an unknown name. */
int         gpm_set_option(gpm_handle_t handle, const char* name, int value);
int         gpm_get_option(gpm_handle_t handle, const char* name, int* value);

/* Step 1.  K = k(X,X) + noise_var*I.  Replaces  cdist(X/l, X/l,'sqeuclidean') -> np.exp -> +sigma^2 I.
 * X: N x D row-major (ldx = D).  K: N x N, leading dimension ldk. */
int gpm_cov(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
            double* K, int64_t ldk, int32_t flags, gpm_stream_t stream);

/* Materialised cross-covariance, query-major:  KsT[m, i] = k(Xs[m], X[i]),  M x N, ld = ldks.
 * Xs may be NULL, in which case points [m0, m1) of `grid` are used (M = m1 - m0). */
int gpm_cross_cov(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
                  const double* Xs, const gpm_grid_t* grid, int64_t m0, int64_t m1,
                  double* KsT, int64_t ldks, gpm_stream_t stream);

/* Step 2.  In-place lower Cholesky K = L L^T (blocked right-looking, FP64 tensor-core updates).
 * Replaces scipy.linalg.cholesky(K, lower=True).  On exit the lower triangle of K holds L; the
 * strict upper triangle is unspecified.  `ws` (gpm_potrf_workspace_bytes(N) bytes) receives the
 * inverses of the 128x128 diagonal blocks of L, which gpm_solve_lml and gpm_predict consume.
 * info: device int32, set to 0 or to the 1-based index of the first non-positive pivot. */
size_t gpm_potrf_workspace_bytes(int64_t N);
int gpm_potrf(gpm_handle_t h, double* K, int64_t N, int64_t ldk, void* ws, int32_t* info,
              gpm_stream_t stream);

/* Step 3.  alpha = K^{-1} Y (two blocked triangular solves) and the log marginal likelihood
 *   lml[r] = -0.5 Y_r^T alpha_r - sum_i log L_ii - (N/2) log(2 pi).
 * Replaces two scipy.linalg.solve_triangular calls + the LML line of R&W Alg. 2.1.
 * Y, alpha: N x R row-major (ld = R), R <= 8; alpha must not alias Y (it is the solve's work
 * buffer).  lml: R doubles (device), may be NULL. */
int gpm_solve_lml(gpm_handle_t h, const double* L, int64_t N, int64_t ldl, const void* potrf_ws,
                  const double* Y, int32_t R, double* alpha, double* lml, gpm_stream_t stream);

/* Steps 1-3 in one call (what GPmap.fit_gp issues): covariance (lower tiles) -> Cholesky -> alpha -> LML, with the
 * forward substitution fused into the factorisation, so L is read once by the solve instead of twice.
 * K (N x ldk, out): the factor L in its lower triangle.  ws: gpm_potrf_workspace_bytes(N) bytes, receives the inverted
 * diagonal blocks exactly as gpm_potrf leaves them (gpm_predict / gpm_lml_grad consume them).  Y, alpha: N x R
 * row-major, R <= 8, alpha must not alias Y.  lml: R doubles (device) or NULL.  info as in gpm_potrf.
 * Option no_fused_solve = 1 runs the three separate steps instead (same results to rounding). */
int gpm_fit(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
            const double* Y, int32_t R, double* K, int64_t ldk, void* ws, double* alpha, double* lml,
            int32_t* info, gpm_stream_t stream);

/* Step 4+5.  Posterior mean and variance at query points.
 *   mu[m, r] = sum_i k(xs_m, x_i) alpha[i, r]             (fused: K* is never stored)
 *   var[m]   = max(0, signal_var - || L^{-1} k(X, xs_m) ||^2)   (blocked TRSM on FP64 tensor cores + row-norm
 *                                                           epilogue; clamped at 0 against cancellation)
 * Replaces  Ks.T @ alpha  and  solve_triangular(L, Ks) -> column sum of squares.
 * Query points: rows [m0, m1) of Xs (M_total x D) if Xs != NULL, else points [m0, m1) of `grid`
 * (so a caller shards the grid across GPUs by choosing [m0, m1)).  mu: (m1-m0) x R, var: (m1-m0);
 * either may be NULL according to `flags`.  ws: gpm_predict_workspace_bytes(N, m1-m0) bytes. */
size_t gpm_predict_workspace_bytes(gpm_handle_t h, int64_t N, int64_t M);
int gpm_predict(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
                const double* L, int64_t ldl, const void* potrf_ws, const double* alpha, int32_t R,
                const double* Xs, const gpm_grid_t* grid, int64_t m0, int64_t m1,
                double* mu, double* var, void* ws, size_t ws_bytes, int32_t flags,
                gpm_stream_t stream);

/* Batched per-path fits: B independent paths of equal length N (GPmap.py:96,117 assume equal
 * length).  Xb: B x N x D, Yb: B x N x R, theta: HOST pointer to (D+2) doubles shared by all
 * paths (theta_stride = 0) or to B x (D+2) per-path values (theta_stride = D+2; the host array must stay
 * valid until the copy enqueued on `stream` has run).  alpha: B x N x R (must not alias Yb), lml: B x R,
 * info: B device int32.  ws: gpm_fit_batched_workspace_bytes(h, B, N) bytes.
 * Paths of N <= 112 samples (GPmap.py:189 resamples every trajectory to 33) are fitted one CTA per path entirely in
 * shared memory and B is limited only by 2^31; longer paths go through batched tile launches with B <= 65535. */
size_t gpm_fit_batched_workspace_bytes(gpm_handle_t h, int64_t B, int64_t N);
int gpm_fit_batched(gpm_handle_t h, const double* Xb, const double* Yb, int64_t B, int64_t N,
                    int32_t D, int32_t R, const double* theta, int64_t theta_stride,
                    double* alpha, double* lml, int32_t* info, void* ws, gpm_stream_t stream);

/* exact gradient of the log marginal likelihood with respect to the LOG
 * hyper-parameters,  grad[r, j] = d lml_r / d log(theta_j),  j over [l_1..l_D, signal_var, noise_var]
 * (R&W eq. 5.9).  Needs the factor and alpha of a finished fit.  K^{-1} is formed on the tensor cores
 * (2 N^3 / 3 flops).  grad: device R x (D+2).  ws: gpm_lml_grad_workspace_bytes(N) bytes. */
size_t gpm_lml_grad_workspace_bytes(int64_t N);
int gpm_lml_grad(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
                 const double* L, int64_t ldl, const void* potrf_ws, const double* alpha, int32_t R,
                 double* grad, void* ws, size_t ws_bytes, gpm_stream_t stream);

/* "Next" row of SURVEY.md section 8f: the reference's actual hot loop, trajectories.kmeansclustering
 * (GPmap.py:36-93), as a kernel pair that keeps the Lloyd iteration on the device.
 *
 * gpm_kmeans_assign -- one assignment step (GPmap.py:72-80,114-121): dist[p, c] = sum_i ||path_p[i] - centroid_c[i]||_2
 * summed sequentially in sample order, assign[p] = first c with the smallest distance (strict '<' as GPmap.py:76).
 * pxT, pyT: n x P (SAMPLE-major, so that a warp's loads are contiguous); cx, cy: k x n; dist (P x k) may be NULL.
 *
 * gpm_kmeans_lloyd -- `iters` (even) full iterations enqueued back to back: assignment, centroid update
 * (calc_mean_traj, GPmap.py:95-112: members summed in path order, then divided by the count -- bit-exact; an empty
 * cluster keeps its centroid) and the convergence test shift = sum_c calc_distance(new_c, old_c) < threshold
 * (GPmap.py:87-90).  Once converged, the remaining kernels of the call are no-ops.  px, py, pt: P x n path-major
 * copies of xs, ys, timestamp; pxT, pyT: n x P.  centroids: 3 x k x n (xs, ys, timestamp planes), in/out.
 * first != 0 resets the device-side state in ws (gpm_kmeans_workspace_bytes).  gpm_kmeans_state reads that state
 * (synchronises the stream): when *iters is odd the final centroids are the first 3*k*n doubles of ws, otherwise
 * they are in `centroids`. */
int gpm_kmeans_assign(gpm_handle_t h, const double* pxT, const double* pyT, int64_t P, int32_t n,
                      const double* cx, const double* cy, int32_t k, double* dist, int32_t* assign,
                      gpm_stream_t stream);
size_t gpm_kmeans_workspace_bytes(int64_t P, int32_t n, int32_t k);
int gpm_kmeans_lloyd(gpm_handle_t h, const double* px, const double* py, const double* pt,
                     const double* pxT, const double* pyT, int64_t P, int32_t n, int32_t k,
                     double* centroids, int32_t* assign, double threshold, int32_t iters, int32_t first,
                     void* ws, gpm_stream_t stream);
int gpm_kmeans_state(gpm_handle_t h, const void* ws, int32_t n, int32_t k, int32_t* iters,
                     int32_t* converged, double* shift, gpm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GPMAP_B200_H */

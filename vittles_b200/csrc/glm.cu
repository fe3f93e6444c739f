// GLM-family instances of the fused X^T f(X V) pass (xtfx.cuh), the dense
// prediction GEMV, and the conjugate-gradient vector kernels.
//
// Family: f(theta, w) = sum_n w_n [ b(z_n) - y_n z_n ],  z = X theta.
//   logistic: b = softplus,  poisson: b = exp,  gaussian: b = z^2/2.
// Reference call sites these kernels replace (closed forms of the autograd
// sweeps): gradient `sensitivity_lib.py:354`, Hessian-vector products behind
// `solver_lib.py:70-98` (get_cg_solver's mat_times_vec), nested JVPs
// `sensitivity_lib.py:440-492,788-807`, prediction GEMV `:245-247`.
#include "xtfx.cuh"
#include "glm.cuh"
#include <cmath>

namespace vt {

__global__ void xtfx_reduce_kernel(const double* partial, int ncta, int Dp, int D, double* out, double alpha,
                                   const double* addvec, double beta_vec) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  double s = 0.0;
  for (int k = 0; k < ncta; ++k) s += partial[(size_t)k * Dp + c];
  s *= alpha;
  if (addvec) s += beta_vec * addvec[c];
  out[c] = s;
}

namespace {

// b^{(k)}(z) for the logistic family is a degree-k polynomial in sigma(z):
// P_1 = s,  P_{k+1} = P_k'(s) * (s - s^2).  Coefficients are built on the host.
struct Poly {
  double c[GLM_MAX_BDERIV + 2];
  int deg;
};

static Poly logistic_poly(int k) {
  double cur[GLM_MAX_BDERIV + 2] = {0.0, 1.0};
  int deg = 1;
  for (int it = 1; it < k; ++it) {
    double der[GLM_MAX_BDERIV + 2] = {0};
    for (int i = 1; i <= deg; ++i) der[i - 1] = i * cur[i];
    double nxt[GLM_MAX_BDERIV + 2] = {0};
    for (int i = 0; i < deg; ++i) {
      nxt[i + 1] += der[i];
      nxt[i + 2] -= der[i];
    }
    deg += 1;
    for (int i = 0; i <= deg; ++i) cur[i] = nxt[i];
  }
  Poly p;
  p.deg = deg;
  for (int i = 0; i < GLM_MAX_BDERIV + 2; ++i) p.c[i] = i <= deg ? cur[i] : 0.0;
  return p;
}

__device__ __forceinline__ double sigmoid(double z) {
  // stable in both tails
  if (z >= 0) {
    const double e = exp(-z);
    return 1.0 / (1.0 + e);
  }
  const double e = exp(z);
  return e / (1.0 + e);
}

// k-th derivative of b at z (k >= 1)
__device__ __forceinline__ double b_deriv(int family, int k, double z, const Poly& poly) {
  if (family == GLM_LOGISTIC) {
    const double s = sigmoid(z);
    double r = poly.c[poly.deg];
    for (int i = poly.deg - 1; i >= 0; --i) r = fma(r, s, poly.c[i]);
    return r;
  }
  if (family == GLM_POISSON) return exp(z);
  return k == 1 ? z : (k == 2 ? 1.0 : 0.0);
}

struct StatsOp {
  const double* y; const double* w;
  double* z; double* resid; double* s;
  int family;
  struct Aux { double y, w; };
  __device__ Aux load(long n) const { return Aux{y[n], w ? w[n] : 1.0}; }
  __device__ double operator()(long n, const double* t, const Aux& a) const {
    const double zz = t[0];
    double mu, var;
    if (family == GLM_LOGISTIC) { mu = sigmoid(zz); var = mu * (1.0 - mu); }
    else if (family == GLM_POISSON) { mu = exp(zz); var = mu; }
    else { mu = zz; var = 1.0; }
    const double r = mu - a.y;
    if (z) z[n] = zz;
    if (resid) resid[n] = r;
    if (s) s[n] = a.w * var;
    return a.w * r;
  }
};

struct HvpOp {
  const double* s;
  struct Aux { double s; };
  __device__ Aux load(long n) const { return Aux{s[n]}; }
  __device__ double operator()(long n, const double* t, const Aux& a) const { return a.s * t[0]; }
};

// coef[n] = w_n * b^{(k)}(z_n): elementwise pre-pass of the directional derivative
__global__ void __launch_bounds__(256) glm_coef_kernel(const double* __restrict__ z, const double* __restrict__ w,
                                                       long N, int family, int k, const Poly poly, double* coef) {
  for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long)gridDim.x * blockDim.x)
    coef[n] = b_deriv(family, k, z[n], poly) * (w ? w[n] : 1.0);
}

template <int Q>
struct DirDerivOp {
  const double* coef;
  struct Aux { double c; };
  __device__ Aux load(long n) const { return Aux{coef[n]}; }
  __device__ double operator()(long n, const double* t, const Aux& a) const {
    double u = a.c;
#pragma unroll
    for (int j = 0; j < Q; ++j) u *= t[j];
    return u;
  }
};

template <class RowOp, int Q, int CPT>
int launch_xtfx(const XtfxParams& p, const RowOp& op, cudaStream_t stream, int* grid_out) {
  constexpr int R = (CPT <= 8) ? 8 / CPT : 1;
  constexpr int CTAS_PER_SM = (R * CPT <= 8) ? 2 : 1;
  const size_t smem = xtfx_smem_bytes(p.Dp, R, Q);
  VT_CUDA(cudaFuncSetAttribute(xtfx_kernel<RowOp, Q, CPT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long nblocks = (p.N + R - 1) / R;
  const long cap = (long)num_sms() * CTAS_PER_SM;
  const int grid = (int)(nblocks < cap ? nblocks : cap);
  xtfx_kernel<RowOp, Q, CPT, R><<<grid, XT_THREADS, smem, stream>>>(p, op);
  VT_LAUNCH_CHECK();
  *grid_out = grid;
  return VT_OK;
}

template <class RowOp, int Q>
int dispatch_cpt(const XtfxParams& p, const RowOp& op, cudaStream_t stream, int* grid_out) {
  const int D = p.D;
  if (D <= 512) return launch_xtfx<RowOp, Q, 1>(p, op, stream, grid_out);
  if (D <= 1024) return launch_xtfx<RowOp, Q, 2>(p, op, stream, grid_out);
  if (D <= 2048) return launch_xtfx<RowOp, Q, 4>(p, op, stream, grid_out);
  if constexpr (Q == 1) {
    if (D <= 4096) return launch_xtfx<RowOp, 1, 8>(p, op, stream, grid_out);
    if (D <= 8192) return launch_xtfx<RowOp, 1, 16>(p, op, stream, grid_out);
  }
  set_error("xtfx: D=%d with %d directions is not supported (D <= 2048, or <= 8192 for one direction)", D, Q);
  return VT_ERR_INVALID;
}

int make_params(XtfxParams& p, const double* X, long ldx, long N, int D, const double* V, double* workspace,
                size_t workspace_bytes, bool need_partial) {
  VT_REQUIRE(X && V, "xtfx: null pointer");
  VT_REQUIRE(N >= 1 && D >= 1 && ldx >= D, "xtfx: bad shape N=%ld D=%d ldx=%ld", N, D, ldx);
  p.X = X; p.ldx = ldx; p.N = N; p.D = D; p.V = V;
  p.Dp = (D + 1) & ~1;
  p.contiguous = (ldx == D);
  p.bulk = (D % 2 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (p.contiguous || ldx % 2 == 0);
  p.partial = nullptr;
  if (need_partial) {
    VT_REQUIRE(workspace && workspace_bytes >= glm_workspace_bytes(D),
               "xtfx: workspace too small: need %zu bytes", glm_workspace_bytes(D));
    p.partial = workspace;
  }
  return VT_OK;
}

}  // namespace

size_t glm_workspace_bytes(int D) { return (size_t)num_sms() * 2 * ((D + 1) & ~1) * 8; }

int glm_stats(const double* X, long ldx, long N, int D, const double* theta, const double* y, const double* w,
              int family, double* z, double* resid, double* s, double* grad, double l2, double* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  VT_REQUIRE(y, "glm_stats: y is null");
  VT_REQUIRE(family >= 0 && family <= 2, "glm_stats: unknown family %d", family);
  XtfxParams p;
  int st = make_params(p, X, ldx, N, D, theta, workspace, workspace_bytes, grad != nullptr);
  if (st != VT_OK) return st;
  StatsOp op{y, w, z, resid, s, family};
  int grid = 0;
  st = dispatch_cpt<StatsOp, 1>(p, op, stream, &grid);
  if (st != VT_OK) return st;
  if (grad) {
    xtfx_reduce_kernel<<<(D + 255) / 256, 256, 0, stream>>>(p.partial, grid, p.Dp, D, grad, 1.0, theta, l2);
    VT_LAUNCH_CHECK();
  }
  return VT_OK;
}

int glm_hvp(const double* X, long ldx, long N, int D, const double* s, const double* v, double ridge, double* out,
            double* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VT_REQUIRE(s && out, "glm_hvp: null pointer");
  XtfxParams p;
  int st = make_params(p, X, ldx, N, D, v, workspace, workspace_bytes, true);
  if (st != VT_OK) return st;
  HvpOp op{s};
  int grid = 0;
  st = dispatch_cpt<HvpOp, 1>(p, op, stream, &grid);
  if (st != VT_OK) return st;
  xtfx_reduce_kernel<<<(D + 255) / 256, 256, 0, stream>>>(p.partial, grid, p.Dp, D, out, 1.0, v, ridge);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

size_t glm_dirderiv_workspace_bytes(long N, int D) { return glm_workspace_bytes(D) + (size_t)N * 8; }

int glm_dirderiv(const double* X, long ldx, long N, int D, const double* z, const double* w, int family,
                 const double* dirs, int q, double* out, double* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
  VT_REQUIRE(z && out, "glm_dirderiv: null pointer");
  VT_REQUIRE(q >= 1 && q <= XT_MAXQ, "glm_dirderiv: number of directions must be 1..%d, got %d", XT_MAXQ, q);
  VT_REQUIRE(q + 1 <= GLM_MAX_BDERIV, "glm_dirderiv: derivative order too high");
  VT_REQUIRE(family >= 0 && family <= 2, "glm_dirderiv: unknown family %d", family);
  VT_REQUIRE(workspace && workspace_bytes >= glm_dirderiv_workspace_bytes(N, D),
             "glm_dirderiv: workspace too small: need %zu bytes", glm_dirderiv_workspace_bytes(N, D));
  XtfxParams p;
  int st = make_params(p, X, ldx, N, D, dirs, workspace, workspace_bytes, true);
  if (st != VT_OK) return st;
  double* coef = workspace + glm_workspace_bytes(D) / 8;
  const Poly poly = logistic_poly(q + 1);
  const long want = (N + 255) / 256;
  const int cgrid = (int)(want < (long)num_sms() * 8 ? want : (long)num_sms() * 8);
  glm_coef_kernel<<<cgrid, 256, 0, stream>>>(z, w, N, family, q + 1, poly, coef);
  VT_LAUNCH_CHECK();
  int grid = 0;
  switch (q) {
    case 1: { DirDerivOp<1> op{coef}; st = dispatch_cpt<DirDerivOp<1>, 1>(p, op, stream, &grid); break; }
    case 2: { DirDerivOp<2> op{coef}; st = dispatch_cpt<DirDerivOp<2>, 2>(p, op, stream, &grid); break; }
    case 3: { DirDerivOp<3> op{coef}; st = dispatch_cpt<DirDerivOp<3>, 3>(p, op, stream, &grid); break; }
    default: { DirDerivOp<4> op{coef}; st = dispatch_cpt<DirDerivOp<4>, 4>(p, op, stream, &grid); break; }
  }
  if (st != VT_OK) return st;
  xtfx_reduce_kernel<<<(D + 255) / 256, 256, 0, stream>>>(p.partial, grid, p.Dp, D, out, 1.0, nullptr, 0.0);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

// ---------------------------------------------------------------------------
// y = alpha * A x + beta * y0  for a row-major A (M x N) with N long: the
// prediction theta_hat + S (lam1 - lam0) (sensitivity_lib.py:245-247) with
// S = (D, N).  HBM-bound: one streaming read of A.  Each CTA owns a (row,
// column-chunk) pair; chunk sums are combined in a fixed order.
// ---------------------------------------------------------------------------
namespace {
constexpr int GEMV_THREADS = 256;
constexpr int GEMV_CHUNK = 256 * 2 * 16;   // columns per CTA pass (64 KB of A)

__global__ void __launch_bounds__(GEMV_THREADS) gemv_rows_kernel(const double* __restrict__ A, long lda, int M, long N,
                                                                  const double* __restrict__ x, double* partial,
                                                                  int nchunks) {
  const int row = blockIdx.y;
  const int chunk = blockIdx.x;
  const long c0 = (long)chunk * GEMV_CHUNK;
  const long c1 = min(N, c0 + GEMV_CHUNK);
  const double* a = A + (long)row * lda;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(x)) % 16 == 0);
  double s = 0.0;
  if (vec) {
    for (long c = c0 + 2 * threadIdx.x; c + 1 < c1; c += 2 * GEMV_THREADS) {
      const double2 av = ld_stream2(a + c);
      const double2 xv = *reinterpret_cast<const double2*>(x + c);
      s = fma(av.x, xv.x, fma(av.y, xv.y, s));
    }
    if (((c1 - c0) & 1) && threadIdx.x == 0) s = fma(a[c1 - 1], x[c1 - 1], s);
  } else {
    for (long c = c0 + threadIdx.x; c < c1; c += GEMV_THREADS) s = fma(a[c], x[c], s);
  }
  __shared__ double red[GEMV_THREADS / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < GEMV_THREADS / 32; ++i) t += red[i];
    partial[(size_t)row * nchunks + chunk] = t;
  }
}

__global__ void gemv_finish_kernel(const double* partial, int M, int nchunks, double alpha, const double* y0,
                                   double beta, double* y) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  double s = 0.0;
  for (int k = 0; k < nchunks; ++k) s += partial[(size_t)row * nchunks + k];
  y[row] = alpha * s + (y0 ? beta * y0[row] : 0.0);
}
}  // namespace

size_t gemv_workspace_bytes(int M, long N) {
  const long nchunks = (N + GEMV_CHUNK - 1) / GEMV_CHUNK;
  return (size_t)M * nchunks * 8;
}

int gemv_rows(const double* A, long lda, int M, long N, const double* x, double alpha, const double* y0, double beta,
              double* y, double* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VT_REQUIRE(A && x && y, "gemv: null pointer");
  VT_REQUIRE(M >= 1 && N >= 1 && lda >= N, "gemv: bad shape");
  VT_REQUIRE(M <= 65535, "gemv: at most 65535 rows");
  const long nchunks = (N + GEMV_CHUNK - 1) / GEMV_CHUNK;
  VT_REQUIRE(workspace && workspace_bytes >= gemv_workspace_bytes(M, N), "gemv: workspace too small");
  dim3 grid((unsigned)nchunks, (unsigned)M);
  gemv_rows_kernel<<<grid, GEMV_THREADS, 0, stream>>>(A, lda, M, N, x, workspace, (int)nchunks);
  VT_LAUNCH_CHECK();
  gemv_finish_kernel<<<(M + 255) / 256, 256, 0, stream>>>(workspace, M, (int)nchunks, alpha, y0, beta, y);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

// ---------------------------------------------------------------------------
// Conjugate-gradient vector kernels.  The iteration follows scipy's cg
// (the reference's solver, solver_lib.py:91-97) step for step so that the
// iteration count matches; all scalars stay on the device:
//   state = {rho, rho_prev, pq, rnorm2}
//   cg_update_p : rho = r.r (computed by previous kernel), p = r + (rho/rho_prev) p
//   cg_update_xr: alpha = rho / (p.q);  x += alpha p;  r -= alpha q;  rnorm2 = r.r
// Single-CTA kernels: D is at most a few thousand, so one block of 1024 threads
// does the vector update and the deterministic reduction in one launch.
// ---------------------------------------------------------------------------
namespace {
__device__ double block_sum_1024(double v, double* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// state[0]=rho, [1]=rho_prev, [2]=pq, [3]=rnorm2, [4]=bnorm2
__global__ void __launch_bounds__(1024) cg_init_kernel(int D, const double* b, double* x, double* r, double* state) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const double bi = b[i];
    x[i] = 0.0;
    r[i] = bi;
    s = fma(bi, bi, s);
  }
  s = block_sum_1024(s, red);
  if (threadIdx.x == 0) { state[0] = s; state[1] = 0.0; state[2] = 0.0; state[3] = s; state[4] = s; }
}

__global__ void __launch_bounds__(1024) cg_update_p_kernel(int D, const double* r, double* p, double* state, int first) {
  const double rho = state[3];   // r.r of the current residual
  const double beta = first ? 0.0 : rho / state[1];
  for (int i = threadIdx.x; i < D; i += blockDim.x) p[i] = first ? r[i] : fma(beta, p[i], r[i]);
  if (threadIdx.x == 0) state[0] = rho;
}

__global__ void __launch_bounds__(1024) cg_update_xr_kernel(int D, const double* p, const double* q, double* x,
                                                             double* r, double* state) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s = fma(p[i], q[i], s);
  const double pq = block_sum_1024(s, red);
  const double rho = state[0];
  const double alpha = rho / pq;
  double rr = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    rr = fma(ri, ri, rr);
  }
  rr = block_sum_1024(rr, red);
  if (threadIdx.x == 0) { state[1] = rho; state[2] = pq; state[3] = rr; }
}
}  // namespace

int cg_init(int D, const double* b, double* x, double* r, double* state, cudaStream_t stream) {
  VT_REQUIRE(D >= 1 && b && x && r && state, "cg_init: bad arguments");
  cg_init_kernel<<<1, 1024, 0, stream>>>(D, b, x, r, state);
  VT_LAUNCH_CHECK();
  return VT_OK;
}
int cg_update_p(int D, const double* r, double* p, double* state, int first, cudaStream_t stream) {
  cg_update_p_kernel<<<1, 1024, 0, stream>>>(D, r, p, state, first);
  VT_LAUNCH_CHECK();
  return VT_OK;
}
int cg_update_xr(int D, const double* p, const double* q, double* x, double* r, double* state, cudaStream_t stream) {
  cg_update_xr_kernel<<<1, 1024, 0, stream>>>(D, p, q, x, r, state);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

}  // namespace vt

// FP64 DMMA GEMM engine - see dgemm.cuh for the design notes.
#include "dgemm.cuh"

namespace vt {

namespace {

struct Unit {
  int m0, n0;      // tile origin
  int kit0, nkit;  // k-iteration range of this unit
  long slot;       // workspace slot (split-K)
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& tm, int& tn) {
  if (p.lower) {
    int i = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
    while ((long)(i + 1) * (i + 2) / 2 <= tile) ++i;
    while ((long)i * (i + 1) / 2 > tile) --i;
    tm = i;
    tn = tile - i * (i + 1) / 2;
  } else {
    tm = tile % p.tiles_m;
    tn = tile / p.tiles_m;
  }
}

__device__ __forceinline__ Unit decode_unit(const GemmParams& p, long u) {
  Unit r;
  const int tile = (int)(u % p.ntiles);
  const int part = (int)(u / p.ntiles);
  int tm, tn;
  tile_coords(p, tile, tm, tn);
  r.m0 = tm * BM;
  r.n0 = tn * BN;
  const int q = p.kiters / p.parts, rem = p.kiters % p.parts;
  r.kit0 = part * q + min(part, rem);
  r.nkit = q + (part < rem ? 1 : 0);
  r.slot = u;
  return r;
}

// Stage one operand tile (128 rows x 16 k) into shared memory.
template <int MODE>
__device__ __forceinline__ void issue_tile(double* __restrict__ s, const double* __restrict__ gp, long ld,
                                           int r0, int R, int k0, int K, int vec, int tid) {
  if (MODE == KC) {
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * GEMM_THREADS;
        const int row = c >> 3, ch = c & 7;
        const int gr = r0 + row, gk = k0 + ch * 2;
        int bytes = (gr < R) ? min(max((K - gk) * 8, 0), 16) : 0;
        const double* src = bytes ? gp + (long)gr * ld + gk : gp;
        cp_async16(s + row * LDKC + ch * 2, src, bytes);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * GEMM_THREADS;
        const int row = e >> 4, kc = e & 15;
        const int gr = r0 + row, gk = k0 + kc;
        const bool ok = (gr < R) && (gk < K);
        cp_async8(s + row * LDKC + kc, ok ? gp + (long)gr * ld + gk : gp, ok ? 8 : 0);
      }
    }
  } else {
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * GEMM_THREADS;
        const int krow = c >> 6, ch = c & 63;
        const int gk = k0 + krow, gr = r0 + ch * 2;
        int bytes = (gk < K) ? min(max((R - gr) * 8, 0), 16) : 0;
        const double* src = bytes ? gp + (long)gk * ld + gr : gp;
        cp_async16(s + krow * LDKS + ch * 2, src, bytes);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * GEMM_THREADS;
        const int krow = e >> 7, r = e & 127;
        const int gk = k0 + krow, gr = r0 + r;
        const bool ok = (gr < R) && (gk < K);
        cp_async8(s + krow * LDKS + r, ok ? gp + (long)gk * ld + gr : gp, ok ? 8 : 0);
      }
    }
  }
}

template <int AMODE, int BMODE, bool KSCALE>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_kernel(const GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * TILE_DOUBLES;
  double* sS = smem + 2 * STAGES * TILE_DOUBLES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp >> 2) * WM;   // 2 warps along M
  const int wn0 = (warp & 3) * WN;    // 4 warps along N

  // fragment base offsets inside a stage
  const int a_off = (AMODE == KC) ? (wm0 + g) * LDKC + tig : tig * LDKS + wm0 + g;
  const int b_off = (BMODE == KC) ? (wn0 + g) * LDKC + tig : tig * LDKS + wn0 + g;
  constexpr int A_MT = (AMODE == KC) ? 8 * LDKC : 8;
  constexpr int A_KK = (AMODE == KC) ? 4 : 4 * LDKS;
  constexpr int B_NT = (BMODE == KC) ? 8 * LDKC : 8;
  constexpr int B_KK = (BMODE == KC) ? 4 : 4 * LDKS;

  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const long total_units = (long)p.ntiles * p.parts;
  long lu = blockIdx.x;
  bool lvalid = lu < total_units;
  Unit lU = decode_unit(p, lvalid ? lu : 0);
  int lk = 0;
  long mu = lu;
  bool mvalid = lvalid;
  Unit mU = lU;
  int mk = 0;

  auto issue_next = [&](int stage) {
    if (lvalid) {
      const int k0 = (lU.kit0 + lk) * BK;
      issue_tile<AMODE>(sA + stage * TILE_DOUBLES, p.A, p.lda, lU.m0, p.M, k0, p.K, p.a_vec, tid);
      issue_tile<BMODE>(sB + stage * TILE_DOUBLES, p.B, p.ldb, lU.n0, p.N, k0, p.K, p.b_vec, tid);
      if (KSCALE && tid < BK) {
        const bool ok = k0 + tid < p.K;
        cp_async8(sS + stage * BK + tid, ok ? p.kscale + k0 + tid : p.kscale, ok ? 8 : 0);
      }
      if (++lk == lU.nkit) {
        lk = 0;
        lu += gridDim.x;
        lvalid = lu < total_units;
        if (lvalid) lU = decode_unit(p, lu);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue_next(s);

  int stage = 0;
  while (mvalid) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue_next((stage + STAGES - 1) % STAGES);

    const double* As = sA + stage * TILE_DOUBLES + a_off;
    const double* Bs = sB + stage * TILE_DOUBLES + b_off;
    const double* Ss = sS + stage * BK + tig;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = As[i * A_MT + kk * A_KK];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = Bs[j * B_NT + kk * B_KK];
      if (KSCALE) {
        const double sc = Ss[kk * 4];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] *= sc;
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    stage = (stage + 1) % STAGES;

    if (++mk == mU.nkit) {
      // ------------------------------------------------------ epilogue ----
      if (p.parts > 1) {
        double* ws = p.workspace + mU.slot * (long)(BM * BN);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const int r = wm0 + i * 8 + g, c = wn0 + j * 8 + 2 * tig;
            *reinterpret_cast<double2*>(ws + r * BN + c) = make_double2(acc[i][j][0], acc[i][j][1]);
          }
      } else {
        double cs[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int c = mU.n0 + wn0 + j * 8 + 2 * tig;
          cs[j][0] = (p.colscale && c < p.N) ? p.colscale[c] : 1.0;
          cs[j][1] = (p.colscale && c + 1 < p.N) ? p.colscale[c + 1] : 1.0;
        }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int r = mU.m0 + wm0 + i * 8 + g;
          if (r < p.M) {
            const double rs = p.alpha * (p.rowscale ? p.rowscale[r] : 1.0);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const int c = mU.n0 + wn0 + j * 8 + 2 * tig;
              if (c >= p.N) continue;
              double v0 = acc[i][j][0] * rs * cs[j][0];
              double v1 = acc[i][j][1] * rs * cs[j][1];
              double* cp = p.C + (long)r * p.ldc + c;
              const bool has1 = c + 1 < p.N;
              if (p.lower) {
                // keep only the lower triangle (exact symmetry comes from mirroring it)
                const bool k0 = r >= c, k1 = has1 && r >= c + 1;
                if (p.beta != 0.0) {
                  if (k0) v0 += p.beta * cp[0];
                  if (k1) v1 += p.beta * cp[1];
                }
                if (k0) cp[0] = v0;
                if (k1) cp[1] = v1;
                if (p.mirror) {
                  if (k0 && r != c) p.C[(long)c * p.ldc + r] = v0;
                  if (k1 && r != c + 1) p.C[(long)(c + 1) * p.ldc + r] = v1;
                }
              } else {
                if (p.beta != 0.0) {
                  v0 += p.beta * cp[0];
                  if (has1) v1 += p.beta * cp[1];
                }
                if (p.c_vec && has1) {
                  *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
                } else {
                  cp[0] = v0;
                  if (has1) cp[1] = v1;
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      mk = 0;
      mu += gridDim.x;
      mvalid = mu < total_units;
      if (mvalid) mU = decode_unit(p, mu);
    }
  }
  cp_async_wait<0>();
}

// Sum split-K partials in a fixed order (bitwise reproducible) and apply the
// same scaling/beta/lower/mirror rules as the direct epilogue.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmParams p) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long per_tile = (long)BM * BN;
  if (idx >= per_tile * p.ntiles) return;
  const int tile = (int)(idx / per_tile);
  const int e = (int)(idx % per_tile);
  int tm, tn;
  tile_coords(p, tile, tm, tn);
  const int r = tm * BM + e / BN, c = tn * BN + e % BN;
  if (r >= p.M || c >= p.N) return;
  if (p.lower && r < c) return;
  double s = 0.0;
  const double* ws = p.workspace + (long)tile * per_tile + e;
  for (int part = 0; part < p.parts; ++part) s += ws[(long)part * p.ntiles * per_tile];
  double v = s * p.alpha;
  if (p.rowscale) v *= p.rowscale[r];
  if (p.colscale) v *= p.colscale[c];
  double* cp = p.C + (long)r * p.ldc + c;
  if (p.beta != 0.0) v += p.beta * cp[0];
  cp[0] = v;
  if (p.lower && p.mirror && r != c) p.C[(long)c * p.ldc + r] = v;
}

template <int AMODE, int BMODE, bool KSCALE>
int launch_variant(const GemmParams& p, int grid, cudaStream_t stream) {
  VT_CUDA(cudaFuncSetAttribute(dgemm_kernel<AMODE, BMODE, KSCALE>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  dgemm_kernel<AMODE, BMODE, KSCALE><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

}  // namespace

int gemm_pick_parts(int ntiles, int kiters, size_t workspace_bytes) {
  const int G = num_sms();
  if (ntiles >= 4 * G) return 1;
  long pmax = kiters / 8;                                  // at least 8 k-iterations per unit
  const long by_ws = (long)(workspace_bytes / ((size_t)ntiles * BM * BN * 8));
  if (by_ws < pmax) pmax = by_ws;
  if (pmax > 4L * G) pmax = 4L * G;
  if (pmax < 2) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int P = 1; P <= pmax; ++P) {
    const long U = (long)ntiles * P;
    const double eff = (double)U / ((double)G * (double)((U + G - 1) / G));
    if (eff > best_eff + 1e-9) { best_eff = eff; best = P; }
    if (eff >= 0.985) break;
  }
  return best;
}

size_t gemm_workspace_bytes(int M, int N, int K, int lower) {
  const int tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
  const int ntiles = lower ? tm * (tm + 1) / 2 : tm * tn;
  const int kiters = (K + BK - 1) / BK;
  const int P = gemm_pick_parts(ntiles, kiters, (size_t)1 << 62);
  return P > 1 ? (size_t)ntiles * P * BM * BN * 8 : 0;
}

int gemm_launch(GemmParams p, cudaStream_t stream) {
  VT_REQUIRE(p.M >= 0 && p.N >= 0 && p.K >= 0, "gemm: negative dimension");
  if (p.M == 0 || p.N == 0) return VT_OK;
  VT_REQUIRE(p.A && p.B && p.C, "gemm: null operand");
  VT_REQUIRE(p.K > 0, "gemm: K must be positive");
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  if (p.lower) VT_REQUIRE(p.M == p.N, "gemm: lower-only output must be square");
  p.ntiles = p.lower ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  p.kiters = (p.K + BK - 1) / BK;
  p.a_vec = (p.lda % 2 == 0) && (reinterpret_cast<uintptr_t>(p.A) % 16 == 0);
  p.b_vec = (p.ldb % 2 == 0) && (reinterpret_cast<uintptr_t>(p.B) % 16 == 0);
  p.c_vec = (p.ldc % 2 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0);
  if (p.parts <= 0) p.parts = (p.workspace ? gemm_pick_parts(p.ntiles, p.kiters, p.workspace_bytes) : 1);
  if (p.parts > p.kiters) p.parts = p.kiters;
  if (p.parts > 1)
    VT_REQUIRE(p.workspace && p.workspace_bytes >= (size_t)p.ntiles * p.parts * BM * BN * 8,
               "gemm: split-K workspace too small (%zu bytes for %d parts x %d tiles)", p.workspace_bytes,
               p.parts, p.ntiles);
  const long units = (long)p.ntiles * p.parts;
  const int grid = (int)(units < num_sms() ? units : num_sms());
  int st;
  if (p.kscale) {
    VT_REQUIRE(p.amode == KS && p.bmode == KS, "gemm: kscale is only implemented for KS x KS operands");
    st = launch_variant<KS, KS, true>(p, grid, stream);
  } else if (p.amode == KC && p.bmode == KC) {
    st = launch_variant<KC, KC, false>(p, grid, stream);
  } else if (p.amode == KC && p.bmode == KS) {
    st = launch_variant<KC, KS, false>(p, grid, stream);
  } else if (p.amode == KS && p.bmode == KC) {
    st = launch_variant<KS, KC, false>(p, grid, stream);
  } else {
    st = launch_variant<KS, KS, false>(p, grid, stream);
  }
  if (st != VT_OK) return st;
  if (p.parts > 1) {
    const long total = (long)p.ntiles * BM * BN;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
    VT_LAUNCH_CHECK();
  }
  return VT_OK;
}

}  // namespace vt

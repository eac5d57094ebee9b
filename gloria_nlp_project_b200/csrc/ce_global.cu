// Small warp-level kernels: global cosine similarity (gloria_loss.py:75-80) and the bidirectional
// cross entropy with arange labels (gloria_loss.py:86-87, 164-170), forward and closed-form backward.
#include <math.h>

#include "common.cuh"

namespace gloria {
namespace {

__global__ void vec_norms(const float* __restrict__ x, float* __restrict__ n, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
  s = warp_sum(s);
  if (lane == 0) n[row] = sqrtf(s);
}

// one CTA per row a of x; warps sweep the rows b of y.  dynamic smem: D floats (x_a)
__global__ void __launch_bounds__(256) global_cos_fwd(const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ xn, const float* __restrict__ yn,
                                                      int Bc, int D, float eps, float* __restrict__ cosm) {
  extern __shared__ float xs[];
  const int a = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) xs[d] = x[(long long)a * D + d];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float na = xn[a];
  for (int b = warp; b < Bc; b += nwarps) {
    const float* yr = y + (long long)b * D;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot = fmaf(xs[d], yr[d], dot);
    dot = warp_sum(dot);
    if (lane == 0) cosm[(long long)a * Bc + b] = dot / fmaxf(na * yn[b], eps);
  }
}

// Gradient w.r.t. the "row side" of cos[a,b] = <x_a, y_b> / max(|x_a||y_b|, eps); dcos is addressed through
// (rs, cs) so the same kernel serves both sides.  dynamic smem: (D + Bc) floats.
// Both phases are sums over the other side's rows b and are latency-bound (every operand is an L2 hit), so each warp
// keeps 16 / 32 independent loads in flight: phase 1 takes two rows b per step in 256-channel pieces, phase 2 gives each
// warp one eighth of the rows b and 256 channels at a time, four rows per step, and the eight partial sums are added in
// a fixed order through shared memory (deterministic).  Before, phase 2 was one dependent chain of Bc loads per thread:
// 171 us for a [512 x 64] caption shard, whose 64 column-side blocks each walk 512 rows.
__global__ void __launch_bounds__(256) global_cos_bwd_side(const float* __restrict__ x_, const float* __restrict__ y_,
                                                           const float* __restrict__ xn_,
                                                           const float* __restrict__ yn_,
                                                           const float* __restrict__ dcos, int Bi, int Bc_, int D,
                                                           float eps, float* __restrict__ dx_, float* __restrict__ dy_) {
  // one launch for both sides: blocks [0, Bi) differentiate w.r.t. x (rows of dcos), blocks [Bi, Bi + Bc) w.r.t. y
  // (columns of dcos); the roles of the two matrices swap for the second group
  const bool side_y = (int)blockIdx.x >= Bi;
  const float* x = side_y ? y_ : x_;
  const float* y = side_y ? x_ : y_;
  const float* xn = side_y ? yn_ : xn_;
  const float* yn = side_y ? xn_ : yn_;
  const long long rs = side_y ? 1 : Bc_, cs = side_y ? Bc_ : 1;
  const int Bc = side_y ? Bi : Bc_;
  float* dx = side_y ? dy_ : dx_;
  extern __shared__ float smem[];
  float* xs = smem;
  float* dd = smem + D;
  __shared__ float red[8];
  __shared__ float part[8][256];
  const int a = side_y ? (int)blockIdx.x - Bi : (int)blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) xs[d] = x[(long long)a * D + d];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = 8;
  const float na = xn[a];
  float esum = 0.f;   // sum_b (dL/d(|x_a||y_b|)) * |y_b|   (lane 0 of each warp)
  // ---- phase 1: dd[b] = g / den, esum; rows b = warp, warp + 8, ... two at a time
  for (int b = warp; b < Bc; b += 2 * NW) {
    const int b1 = b + NW;
    const bool has1 = b1 < Bc;
    const float* y0 = y + (long long)b * D;
    const float* y1 = y + (long long)(has1 ? b1 : b) * D;
    float dot0 = 0.f, dot1 = 0.f;
    for (int d0 = 0; d0 < D; d0 += 256) {
      float v0[8], v1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = d0 + k * 32 + lane;
        v0[k] = d < D ? __ldg(y0 + d) : 0.f;
        v1[k] = d < D ? __ldg(y1 + d) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = d0 + k * 32 + lane;
        const float xv = d < D ? xs[d] : 0.f;
        dot0 = fmaf(xv, v0[k], dot0);
        dot1 = fmaf(xv, v1[k], dot1);
      }
    }
    dot0 = warp_sum(dot0);
    dot1 = warp_sum(dot1);
    if (lane == 0) {
      {
        const float g = dcos[a * rs + b * cs];
        const float prod = na * yn[b];
        const float den = fmaxf(prod, eps);
        dd[b] = g / den;
        if (prod >= eps) esum += -g * dot0 / (den * den) * yn[b];
      }
      if (has1) {
        const float g = dcos[a * rs + b1 * cs];
        const float prod = na * yn[b1];
        const float den = fmaxf(prod, eps);
        dd[b1] = g / den;
        if (prod >= eps) esum += -g * dot1 / (den * den) * yn[b1];
      }
    }
  }
  if (lane == 0) red[warp] = esum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) tot += red[w];
  const float coef = na > 0.f ? tot / na : 0.f;
  // ---- phase 2: dx[a, d] = coef x[a, d] + sum_b dd[b] y[b, d]
  const int per = (Bc + NW - 1) / NW;
  const int bs = warp * per, be = min(Bc, bs + per);
  for (int d0 = 0; d0 < D; d0 += 256) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int b = bs;
    for (; b + 4 <= be; b += 4) {
      float v[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int d = d0 + k * 32 + lane;
          v[r][k] = d < D ? __ldg(y + (long long)(b + r) * D + d) : 0.f;
        }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float w = dd[b + r];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, v[r][k], acc[k]);
      }
    }
    for (; b < be; ++b) {
      const float w = dd[b];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = d0 + k * 32 + lane;
        if (d < D) acc[k] = fmaf(w, __ldg(y + (long long)b * D + d), acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[warp][k * 32 + lane] = acc[k];
    __syncthreads();
    {
      const int d = d0 + (int)threadIdx.x;
      if (d < D) {
        float s = coef * xs[d];
#pragma unroll
        for (int w = 0; w < NW; ++w) s += part[w][threadIdx.x];
        dx[(long long)a * D + d] = s;
      }
    }
    __syncthreads();
  }
}

// lse over rows (blockIdx.y == 0) and columns (blockIdx.y == 1) of scale * m; one warp per row/column
__global__ void ce_lse(const float* __restrict__ m, int B, float scale, float* __restrict__ row_lse,
                       float* __restrict__ col_lse) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= B) return;
  const int lane = threadIdx.x & 31;
  const bool col = blockIdx.y == 1;
  const long long base = col ? idx : (long long)idx * B;
  const long long step = col ? B : 1;
  float mx = -INFINITY;
  for (int k = lane; k < B; k += 32) mx = fmaxf(mx, scale * m[base + k * step]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < B; k += 32) s += expf(scale * m[base + k * step] - mx);
  s = warp_sum(s);
  if (lane == 0) (col ? col_lse : row_lse)[idx] = mx + logf(s);
}

__global__ void __launch_bounds__(256) ce_losses(const float* __restrict__ m, int B, float scale,
                                                 const float* __restrict__ row_lse,
                                                 const float* __restrict__ col_lse, float* __restrict__ losses) {
  __shared__ float r0[8], r1[8];
  float a0 = 0.f, a1 = 0.f;
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    const float d = scale * m[(long long)k * B + k];
    a0 += row_lse[k] - d;
    a1 += col_lse[k] - d;
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = a0; r1[threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s0 += r0[w]; s1 += r1[w]; }
    losses[0] = s0 / (float)B;
    losses[1] = s1 / (float)B;
  }
}

__global__ void ce_bwd(const float* __restrict__ m, int B, float scale, const float* __restrict__ row_lse,
                       const float* __restrict__ col_lse, const float* __restrict__ g, float* __restrict__ dm) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * B) return;
  const int a = (int)(idx / B), b = (int)(idx % B);
  const float z = scale * m[idx];
  const float delta = (a == b) ? 1.f : 0.f;
  const float v = g[0] * (expf(z - row_lse[a]) - delta) + g[1] * (expf(z - col_lse[b]) - delta);
  dm[idx] = v * scale / (float)B;
}

}  // namespace
}  // namespace gloria

using namespace gloria;

extern "C" int gloria_b200_global_sim_fwd(const float* x, const float* y, int Bi, int Bc, int D, float eps,
                                          float* cosm, float* xn, float* yn, void* stream) {
  GLORIA_CHECK_ARG(x && y && cosm && xn && yn, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && D > 0 && D <= 12000, "bad sizes Bi=%d Bc=%d D=%d", Bi, Bc, D);
  cudaStream_t st = (cudaStream_t)stream;
  vec_norms<<<(Bi + 7) / 8, 256, 0, st>>>(x, xn, Bi, D);
  GLORIA_LAUNCHED("vec_norms");
  vec_norms<<<(Bc + 7) / 8, 256, 0, st>>>(y, yn, Bc, D);
  GLORIA_LAUNCHED("vec_norms");
  global_cos_fwd<<<Bi, 256, D * sizeof(float), st>>>(x, y, xn, yn, Bc, D, eps, cosm);
  GLORIA_LAUNCHED("global_cos_fwd");
  return GLORIA_OK;
}

extern "C" int gloria_b200_global_sim_bwd(const float* x, const float* y, const float* xn, const float* yn,
                                          const float* dcos, int Bi, int Bc, int D, float eps, float* dx, float* dy,
                                          void* stream) {
  GLORIA_CHECK_ARG(x && y && xn && yn && dcos && dx && dy, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && D > 0, "bad sizes Bi=%d Bc=%d D=%d", Bi, Bc, D);
  // (the kernel also holds 8.2 KB of static shared memory for its partial sums)
  GLORIA_CHECK_ARG((size_t)(D + (Bi > Bc ? Bi : Bc)) * sizeof(float) <= 39 * 1024,
                   "global_sim_bwd: D + B = %d exceeds the 39 KB shared-memory tile", D + (Bi > Bc ? Bi : Bc));
  cudaStream_t st = (cudaStream_t)stream;
  global_cos_bwd_side<<<Bi + Bc, 256, (D + (Bi > Bc ? Bi : Bc)) * sizeof(float), st>>>(x, y, xn, yn, dcos, Bi, Bc, D, eps,
                                                                                        dx, dy);
  GLORIA_LAUNCHED("global_cos_bwd_side");
  return GLORIA_OK;
}

extern "C" int gloria_b200_ce_bidir_fwd(const float* m, int B, float scale, float* losses, float* row_lse,
                                        float* col_lse, void* stream) {
  GLORIA_CHECK_ARG(m && losses && row_lse && col_lse, "null pointer");
  GLORIA_CHECK_ARG(B > 0, "bad B=%d", B);
  cudaStream_t st = (cudaStream_t)stream;
  ce_lse<<<dim3((B + 7) / 8, 2), 256, 0, st>>>(m, B, scale, row_lse, col_lse);
  GLORIA_LAUNCHED("ce_lse");
  ce_losses<<<1, 256, 0, st>>>(m, B, scale, row_lse, col_lse, losses);
  GLORIA_LAUNCHED("ce_losses");
  return GLORIA_OK;
}

extern "C" int gloria_b200_ce_bidir_bwd(const float* m, int B, float scale, const float* row_lse,
                                        const float* col_lse, const float* g, float* dm, void* stream) {
  GLORIA_CHECK_ARG(m && row_lse && col_lse && g && dm, "null pointer");
  GLORIA_CHECK_ARG(B > 0, "bad B=%d", B);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * B;
  ce_bwd<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, B, scale, row_lse, col_lse, g, dm);
  GLORIA_LAUNCHED("ce_bwd");
  return GLORIA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Row-wise cosine of two [N, D] matrices (cosine_similarity, gloria_loss.py:11-16), forward and backward.
// One warp per row.  stats [N, 3] = (dot, |x1|, |x2|) saved for the backward.
// ---------------------------------------------------------------------------------------------------------------
namespace gloria {
namespace {
__global__ void row_cosine_fwd(const float* __restrict__ a, const float* __restrict__ b, long long N, int D, float eps,
                               float* __restrict__ out, float* __restrict__ stats) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float* ar = a + row * D;
  const float* br = b + row * D;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = ar[d], y = br[d];
    dot = fmaf(x, y, dot); na = fmaf(x, x, na); nb = fmaf(y, y, nb);
  }
  dot = warp_sum(dot); na = sqrtf(warp_sum(na)); nb = sqrtf(warp_sum(nb));
  if (lane == 0) {
    out[row] = dot / fmaxf(na * nb, eps);
    if (stats) { stats[row * 3] = dot; stats[row * 3 + 1] = na; stats[row * 3 + 2] = nb; }
  }
}
__global__ void row_cosine_bwd(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ stats,
                               const float* __restrict__ g, long long N, int D, float eps, float* __restrict__ da,
                               float* __restrict__ db) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float dot = stats[row * 3], na = stats[row * 3 + 1], nb = stats[row * 3 + 2];
  const float prod = na * nb, den = fmaxf(prod, eps), gr = g[row];
  const float ddot = gr / den;
  const float dden = (prod >= eps) ? -gr * dot / (den * den) : 0.f;
  const float ca = na > 0.f ? dden * nb / na : 0.f;    // (dL/d|a|) / |a|
  const float cb = nb > 0.f ? dden * na / nb : 0.f;
  const float* ar = a + row * D;
  const float* br = b + row * D;
  for (int d = lane; d < D; d += 32) {
    const float x = ar[d], y = br[d];
    da[row * D + d] = fmaf(ddot, y, ca * x);
    db[row * D + d] = fmaf(ddot, x, cb * y);
  }
}
}  // namespace
}  // namespace gloria

extern "C" int gloria_b200_row_cosine_fwd(const float* x1, const float* x2, long long N, int D, float eps, float* out,
                                          float* stats, void* stream) {
  GLORIA_CHECK_ARG(x1 && x2 && out && N > 0 && D > 0, "bad arguments");
  gloria::row_cosine_fwd<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x1, x2, N, D, eps, out, stats);
  GLORIA_LAUNCHED("row_cosine_fwd");
  return GLORIA_OK;
}

extern "C" int gloria_b200_row_cosine_bwd(const float* x1, const float* x2, const float* stats, const float* dout,
                                          long long N, int D, float eps, float* dx1, float* dx2, void* stream) {
  GLORIA_CHECK_ARG(x1 && x2 && stats && dout && dx1 && dx2 && N > 0 && D > 0, "bad arguments");
  gloria::row_cosine_bwd<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x1, x2, stats, dout, N, D, eps, dx1, dx2);
  GLORIA_LAUNCHED("row_cosine_bwd");
  return GLORIA_OK;
}

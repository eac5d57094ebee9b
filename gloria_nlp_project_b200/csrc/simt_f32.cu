// fp32 mode of the GLoRIA local similarity: CUDA-core FFMA kernels, fp32 everywhere.
//
// This is the exact-arithmetic mode (logits within 1e-5 of the un-autocast reference) and the mode that carries
// every optional output (diagonal attention maps, word-mean attention for the entropy / KL / no-attn regularisers).
// It follows the reference op graph (gloria/loss/gloria_loss.py:19-63, 116-160) but batched over all
// (image, caption) pairs of a caption chunk, with no transposing copies of the context and no autograd tape:
// the backward recomputes scores / P / A / C per chunk and applies the closed-form gradients.
//
// Pair index inside a chunk of `nc` captions starting at i0:  p = j * nc + (i - i0)   (image-major).
#include <math.h>

#include "common.cuh"

namespace gloria {
namespace {

// -------------------------------------------------------------------------------------------------------------
// Generic strided fp32 GEMM with two batch indices and an inner reduction index:
//   C[b0,b1] (M x N)  (+)=  sum_{r < R}  A[b0,b1,r] (M x K_r)  *  B[b0,b1,r] (K_r x N)
// Optional device-side limits keep padded caption rows out of the arithmetic:
//   mlim: M_eff = min(M, mlim[mlim_off + b1])      klim: K_r = min(K, klim[klim_off + r])
// -------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K, R;
  long long sAm, sAk, sAb0, sAb1, sAr;
  long long sBk, sBn, sBb0, sBb1, sBr;
  long long sCm, sCn, sCb0, sCb1;
  int nb0, nb1;
  float beta;
  const int* mlim;
  const int* klim;
  int mlim_off, klim_off;
};

constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int b0 = b / g.nb1, b1 = b % g.nb1;
  const int m0 = blockIdx.z * GBM, n0 = blockIdx.y * GBN;
  int Meff = g.M;
  if (g.mlim) Meff = min(Meff, max(g.mlim[g.mlim_off + b1], 0));
  if (m0 >= Meff) return;
  const float* Ab = g.A + b0 * g.sAb0 + b1 * g.sAb1;
  const float* Bb = g.B + b0 * g.sBb0 + b1 * g.sBb1;
  float* Cb = g.C + b0 * g.sCb0 + b1 * g.sCb1;
  const bool a_kfast = (g.sAk == 1);
  const bool b_nfast = (g.sBn == 1);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r = 0; r < g.R; ++r) {
    int Keff = g.K;
    if (g.klim) Keff = min(Keff, max(g.klim[g.klim_off + r], 0));
    const float* Ar = Ab + r * g.sAr;
    const float* Br = Bb + r * g.sBr;
    for (int k0 = 0; k0 < Keff; k0 += GBK) {
#pragma unroll
      for (int e = tid; e < GBM * GBK; e += 256) {
        int m, k;
        if (a_kfast) { k = e % GBK; m = e / GBK; } else { m = e % GBM; k = e / GBM; }
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < Meff && gk < Keff) ? __ldg(Ar + gm * g.sAm + gk * g.sAk) : 0.f;
      }
#pragma unroll
      for (int e = tid; e < GBN * GBK; e += 256) {
        int n, k;
        if (b_nfast) { n = e % GBN; k = e / GBN; } else { k = e % GBK; n = e / GBK; }
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < g.N && gk < Keff) ? __ldg(Br + gk * g.sBk + gn * g.sBn) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < GBK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= Meff) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float* p = Cb + gm * g.sCm + gn * g.sCn;
      float v = acc[i][j];
      if (g.beta != 0.f) v += g.beta * (*p);
      *p = v;
    }
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.nb0 * g.nb1 <= 0) return GLORIA_OK;
  dim3 grid((unsigned)(g.nb0 * g.nb1), (unsigned)((g.N + GBN - 1) / GBN), (unsigned)((g.M + GBM - 1) / GBM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>(g);
  GLORIA_LAUNCHED("gemm_f32_kernel");
  return GLORIA_OK;
}

// -------------------------------------------------------------------------------------------------------------
// words [Bc, D, Lw]  ->  Wt [Bc, Lw, D] (word-major rows) and |W_l| (gloria_loss.py:14)
// -------------------------------------------------------------------------------------------------------------
__global__ void transpose_dl_to_ld(const float* __restrict__ in, float* __restrict__ out, int D, int L) {
  // in [b][D][L] -> out [b][L][D]
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const float* ib = in + (long long)b * D * L;
  float* ob = out + (long long)b * D * L;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, l = l0 + threadIdx.x;
    t[r][threadIdx.x] = (d < D && l < L) ? ib[(long long)d * L + l] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    if (l < L && d < D) ob[(long long)l * D + d] = t[threadIdx.x][r];
  }
}

__global__ void row_norms(const float* __restrict__ x, float* __restrict__ n, long long rows, int D) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
  s = warp_sum(s);
  if (lane == 0) n[row] = sqrtf(s);
}

// d_words [Bc, D, Lw] = transpose of dWt [Bc, Lw, D], zero outside [off, off + cap_len)
__global__ void unpack_dwords(const float* __restrict__ dWt, float* __restrict__ dwords,
                              const int* __restrict__ cap_lens, int D, int Lw, int Lcap, int off, int accumulate) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int L = min(max(cap_lens[b], 0), Lcap);
  const float* ib = dWt + (long long)b * D * Lw;
  float* ob = dwords + (long long)b * D * Lw;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    const bool live = (l < Lw && d < D && l >= off && l < off + L);
    t[r][threadIdx.x] = live ? ib[(long long)l * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, l = l0 + threadIdx.x;
    if (d < D && l < Lw) {
      float* o = ob + (long long)d * Lw + l;
      *o = accumulate ? *o + t[threadIdx.x][r] : t[threadIdx.x][r];
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// Double softmax (gloria_loss.py:42-53).  One CTA per pair.  sc [P, Lcap, S] holds the scores S_[l][s] on entry
// and the word-softmax P on exit; at [P, Lcap, S] receives A (rows l >= cap_len zero-filled).
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) double_softmax_fwd(float* __restrict__ sc, float* __restrict__ at,
                                                          const int* __restrict__ cap_lens, int i0, int nc, int Bc,
                                                          int Lcap, int S, float temp1,
                                                          float* __restrict__ attn_diag,
                                                          float* __restrict__ attn_mean, int diag_only) {
  const int p = blockIdx.x;
  const int i = diag_only ? i0 + p : i0 + p % nc;
  const int j = diag_only ? i : p / nc;
  const int L = min(max(cap_lens[i], 0), Lcap);
  float* s = sc + (long long)p * Lcap * S;
  float* a = at + (long long)p * Lcap * S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // softmax #1: over the caption's words, for every region (thread-per-region, coalesced over s)
  for (int x = tid; x < S; x += blockDim.x) {
    float m = -INFINITY;
    for (int l = 0; l < L; ++l) m = fmaxf(m, s[(long long)l * S + x]);
    float den = 0.f;
    for (int l = 0; l < L; ++l) den += expf(s[(long long)l * S + x] - m);
    const float inv = 1.f / den;
    for (int l = 0; l < L; ++l) s[(long long)l * S + x] = expf(s[(long long)l * S + x] - m) * inv;
  }
  __syncthreads();
  // softmax #2: x temp1, over the regions, for every word (warp-per-word)
  float* diag = (attn_diag != nullptr && j == i) ? attn_diag + (long long)i * Lcap * S : nullptr;
  for (int l = warp; l < Lcap; l += nwarps) {
    float* ar = a + (long long)l * S;
    if (l < L) {
      const float* pr = s + (long long)l * S;
      float z = 0.f;
      for (int x = lane; x < S; x += 32) z += expf(temp1 * pr[x]);
      z = warp_sum(z);
      const float inv = 1.f / z;
      for (int x = lane; x < S; x += 32) {
        const float v = expf(temp1 * pr[x]) * inv;
        ar[x] = v;
        if (diag) diag[(long long)l * S + x] = v;
      }
    } else {
      for (int x = lane; x < S; x += 32) {
        ar[x] = 0.f;
        if (diag) diag[(long long)l * S + x] = 0.f;
      }
    }
  }
  if (attn_mean != nullptr) {
    __syncthreads();
    float* mo = attn_mean + ((long long)j * Bc + i) * S;
    const float invL = L > 0 ? 1.f / (float)L : 0.f;
    for (int x = tid; x < S; x += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc += a[(long long)l * S + x];
      mo[x] = acc * invL;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// Per-word cosine (gloria_loss.py:11-16,150) + aggregation over words (:153-158 / gloria_model.py:198-201).
// One CTA per pair.  In backward mode (dsim != nullptr) it also emits the per-word coefficients
//   ddot, beta = dnc/nc, gamma = dnw/nw   of the closed-form backward.
// dynamic smem: 3 * Lcap floats
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cosine_agg(const float* __restrict__ cx, const float* __restrict__ Wt,
                                                  const float* __restrict__ wn, const int* __restrict__ cap_lens,
                                                  int i0, int nc, int Bc, int Lcap, int Lw, int off, int D,
                                                  float temp2, int agg, float eps, float* __restrict__ sim,
                                                  const float* __restrict__ dsim, float* __restrict__ coef) {
  extern __shared__ float sm[];
  float* r_s = sm;
  float* dot_s = sm + Lcap;
  float* nc_s = sm + 2 * Lcap;
  const int p = blockIdx.x;
  const int j = p / nc, i = i0 + p % nc;
  const int L = min(max(cap_lens[i], 0), Lcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int l = warp; l < L; l += nwarps) {
    const float* w = Wt + ((long long)i * Lw + off + l) * D;
    const float* c = cx + ((long long)p * Lcap + l) * D;
    float dot = 0.f, c2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float cv = c[d];
      dot = fmaf(w[d], cv, dot);
      c2 = fmaf(cv, cv, c2);
    }
    dot = warp_sum(dot);
    c2 = warp_sum(c2);
    if (lane == 0) {
      const float ncv = sqrtf(c2);
      const float den = fmaxf(wn[(long long)i * Lw + off + l] * ncv, eps);
      r_s[l] = dot / den;
      dot_s[l] = dot;
      nc_s[l] = ncv;
    }
  }
  __syncthreads();
  if (warp != 0) return;
  float m = -INFINITY;
  for (int l = lane; l < L; l += 32) m = fmaxf(m, r_s[l]);
  m = warp_max(m);
  float sum = 0.f;
  for (int l = lane; l < L; l += 32) sum += expf(temp2 * (r_s[l] - m));
  sum = warp_sum(sum);
  if (lane == 0) {
    float v;
    if (agg == GLORIA_AGG_MAX) v = temp2 * m;
    else {
      v = temp2 * m + logf(sum);
      if (agg == GLORIA_AGG_MEAN) v -= logf((float)L);
    }
    if (sim != nullptr) sim[(long long)j * Bc + i] = v;
  }
  if (dsim == nullptr) return;
  const float g = dsim[(long long)j * Bc + i];
  float* cf = coef + (long long)p * 3 * Lcap;
  for (int l = lane; l < L; l += 32) {
    const float q = expf(temp2 * (r_s[l] - m)) / sum;
    const float dr = g * temp2 * q;
    const float nwv = wn[(long long)i * Lw + off + l], ncv = nc_s[l], dot = dot_s[l];
    const float prod = nwv * ncv;
    const float den = fmaxf(prod, eps);
    const float ddot = dr / den;
    const float dden = (prod >= eps) ? -dr * dot / (den * den) : 0.f;
    const float beta = ncv > 0.f ? dden * nwv / ncv : 0.f;    // (dL/d|C|) / |C|
    const float gamma = nwv > 0.f ? dden * ncv / nwv : 0.f;   // (dL/d|W|) / |W|
    cf[l] = ddot;
    cf[Lcap + l] = beta;
    cf[2 * Lcap + l] = gamma;
  }
}

// dC = ddot * W + beta * C (in place over C), and the direct word gradient
// dWt[i][off+l][:] = sum_j ddot * C + gamma * W.   grid (Lcap, nc); one CTA owns word l of caption i0+ii.
__global__ void __launch_bounds__(256) bwd_context_grad(float* __restrict__ cx, const float* __restrict__ Wt,
                                                        float* __restrict__ dWt, const float* __restrict__ coef,
                                                        const int* __restrict__ cap_lens, int i0, int nc, int Bi,
                                                        int Lcap, int Lw, int off, int D) {
  const int l = blockIdx.x, ii = blockIdx.y, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  if (l >= L) return;
  const float* w = Wt + ((long long)i * Lw + off + l) * D;
  float* dw = dWt + ((long long)i * Lw + off + l) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float wv = w[d];
    float acc = 0.f;
    for (int j = 0; j < Bi; ++j) {
      const long long p = (long long)j * nc + ii;
      const float* cf = coef + p * 3 * Lcap;
      const float ddot = cf[l], beta = cf[Lcap + l], gamma = cf[2 * Lcap + l];
      float* c = cx + (p * Lcap + l) * D + d;
      const float cv = *c;
      *c = fmaf(ddot, wv, beta * cv);
      acc += fmaf(ddot, cv, gamma * wv);
    }
    dw[d] = acc;
  }
}

// Backward of the two softmaxes.  da [P, Lcap, S] holds dA on entry and dS_[l][s] on exit.
__global__ void __launch_bounds__(256) double_softmax_bwd(float* __restrict__ da, const float* __restrict__ at,
                                                          const float* __restrict__ pr,
                                                          const int* __restrict__ cap_lens, int i0, int nc, int Bc,
                                                          int Lcap, int S, float temp1,
                                                          const float* __restrict__ d_attn_diag,
                                                          const float* __restrict__ d_attn_mean, int diag_only) {
  const int p = blockIdx.x;
  const int i = diag_only ? i0 + p : i0 + p % nc;
  const int j = diag_only ? i : p / nc;
  const int L = min(max(cap_lens[i], 0), Lcap);
  float* g = da + (long long)p * Lcap * S;
  const float* a = at + (long long)p * Lcap * S;
  const float* pp = pr + (long long)p * Lcap * S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* ed = (d_attn_diag != nullptr && j == i) ? d_attn_diag + (long long)i * Lcap * S : nullptr;
  const float* em = (d_attn_mean != nullptr) ? d_attn_mean + ((long long)j * Bc + i) * S : nullptr;
  const float invL = L > 0 ? 1.f / (float)L : 0.f;
  // softmax #2 backward: dZ = A * (dA - sum_s A dA);  dP = temp1 * dZ
  for (int l = warp; l < L; l += nwarps) {
    float* gr = g + (long long)l * S;
    const float* ar = a + (long long)l * S;
    float rs = 0.f;
    for (int x = lane; x < S; x += 32) {
      float v = gr[x];
      if (ed) v += ed[(long long)l * S + x];
      if (em) v += em[x] * invL;
      gr[x] = v;
      rs = fmaf(ar[x], v, rs);
    }
    rs = warp_sum(rs);
    for (int x = lane; x < S; x += 32) gr[x] = temp1 * ar[x] * (gr[x] - rs);
  }
  __syncthreads();
  // softmax #1 backward: dS = P * (dP - sum_l P dP)
  for (int x = tid; x < S; x += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < L; ++l) t = fmaf(pp[(long long)l * S + x], g[(long long)l * S + x], t);
    for (int l = 0; l < L; ++l) {
      const long long o = (long long)l * S + x;
      g[o] = pp[o] * (g[o] - t);
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// host orchestration
// -------------------------------------------------------------------------------------------------------------
struct Plan {
  int nc;                 // captions per chunk
  size_t off_wt, off_wn, off_dwt, off_sc, off_at, off_cx, off_da, off_coef, total;
};

size_t per_caption_bytes(int Bi, int D, int S, int Lcap) {
  // sc, at, da: Lcap*S each; cx: Lcap*D; coef: 3*Lcap   (per pair), times Bi pairs per caption
  return (size_t)Bi * ((size_t)Lcap * (3 * (size_t)S + D) + 3 * (size_t)Lcap) * sizeof(float) + 1024;
}

size_t fixed_bytes(int Bc, int D, int Lw) {
  // Wt, dWt [Bc, Lw, D], wn [Bc, Lw]
  return 2 * align_up((size_t)Bc * Lw * D * sizeof(float), 256) + align_up((size_t)Bc * Lw * sizeof(float), 256) + 4096;
}

Plan make_plan(int Bi, int Bc, int D, int S, int Lw, int Lcap, size_t bytes) {
  Plan pl{};
  const size_t fixed = fixed_bytes(Bc, D, Lw);
  const size_t per = per_caption_bytes(Bi, D, S, Lcap);
  if (bytes < fixed + per) { pl.nc = 0; return pl; }
  size_t nc = (bytes - fixed) / per;
  if (nc > (size_t)Bc) nc = Bc;
  pl.nc = (int)nc;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 256); return r; };
  pl.off_wt = take((size_t)Bc * Lw * D * sizeof(float));
  pl.off_dwt = take((size_t)Bc * Lw * D * sizeof(float));
  pl.off_wn = take((size_t)Bc * Lw * sizeof(float));
  const size_t P = (size_t)Bi * nc;
  pl.off_sc = take(P * Lcap * S * sizeof(float));
  pl.off_at = take(P * Lcap * S * sizeof(float));
  pl.off_da = take(P * Lcap * S * sizeof(float));
  pl.off_cx = take(P * Lcap * D * sizeof(float));
  pl.off_coef = take(P * 3 * Lcap * sizeof(float));
  pl.total = o;
  if (pl.total > bytes) pl.nc = 0;
  return pl;
}

int check_common(const void* ctx, const void* words, const void* cap_lens, int Bi, int Bc, int D, int S, int Lw,
                 int Lcap, int word_off, int agg) {
  GLORIA_CHECK_ARG(ctx && words && cap_lens, "null input pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && D > 0 && S > 0 && Lw > 0, "non-positive size (Bi=%d Bc=%d D=%d S=%d Lw=%d)",
                   Bi, Bc, D, S, Lw);
  GLORIA_CHECK_ARG(word_off >= 0 && Lcap > 0 && word_off + Lcap <= Lw,
                   "caption window [%d, %d) exceeds the word axis (%d)", word_off, word_off + Lcap, Lw);
  GLORIA_CHECK_ARG(agg == GLORIA_AGG_SUM || agg == GLORIA_AGG_MEAN || agg == GLORIA_AGG_MAX, "bad agg %d", agg);
  return GLORIA_OK;
}

int prepack_words(const float* words, float* Wt, float* wn, int Bc, int D, int Lw, cudaStream_t st) {
  dim3 grid((Lw + 31) / 32, (D + 31) / 32, Bc), block(32, 8);
  transpose_dl_to_ld<<<grid, block, 0, st>>>(words, Wt, D, Lw);
  GLORIA_LAUNCHED("transpose_dl_to_ld");
  const long long rows = (long long)Bc * Lw;
  row_norms<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(Wt, wn, rows, D);
  GLORIA_LAUNCHED("row_norms");
  return GLORIA_OK;
}

// scores, P, A, C for the chunk [i0, i0+nc)
int chunk_forward(const float* ctx, const float* Wt, const int32_t* cap_lens, int Bi, int Bc, int D, int S, int Lw,
                  int Lcap, int off, float temp1, int i0, int nc, float* sc, float* at, float* cx, float* attn_diag,
                  float* attn_mean, cudaStream_t st) {
  int rc;
  {  // S_[p][l][s] = sum_d Wt[i][off+l][d] * ctx[j][d][s]           (bmm #1, gloria_loss.py:40)
    GemmArgs g{};
    g.A = Wt + ((long long)i0 * Lw + off) * D; g.B = ctx; g.C = sc;
    g.M = Lcap; g.N = S; g.K = D; g.R = 1;
    g.sAm = D; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)Lw * D; g.sAr = 0;
    g.sBk = S; g.sBn = 1; g.sBb0 = (long long)D * S; g.sBb1 = 0; g.sBr = 0;
    g.sCm = S; g.sCn = 1; g.sCb0 = (long long)nc * Lcap * S; g.sCb1 = (long long)Lcap * S;
    g.nb0 = Bi; g.nb1 = nc; g.beta = 0.f; g.mlim = cap_lens; g.mlim_off = i0;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  double_softmax_fwd<<<(unsigned)(Bi * nc), 256, 0, st>>>(sc, at, cap_lens, i0, nc, Bc, Lcap, S, temp1, attn_diag,
                                                         attn_mean, 0);
  GLORIA_LAUNCHED("double_softmax_fwd");
  {  // C[p][l][d] = sum_s A[p][l][s] * ctx[j][d][s]                  (bmm #2, gloria_loss.py:59)
    GemmArgs g{};
    g.A = at; g.B = ctx; g.C = cx;
    g.M = Lcap; g.N = D; g.K = S; g.R = 1;
    g.sAm = S; g.sAk = 1; g.sAb0 = (long long)nc * Lcap * S; g.sAb1 = (long long)Lcap * S;
    g.sBk = 1; g.sBn = S; g.sBb0 = (long long)D * S; g.sBb1 = 0;
    g.sCm = D; g.sCn = 1; g.sCb0 = (long long)nc * Lcap * D; g.sCb1 = (long long)Lcap * D;
    g.nb0 = Bi; g.nb1 = nc; g.beta = 0.f; g.mlim = cap_lens; g.mlim_off = i0;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  return GLORIA_OK;
}

}  // namespace
}  // namespace gloria

using namespace gloria;

extern "C" size_t gloria_b200_local_f32_workspace(int Bi, int Bc, int D, int S, int Lw, int Lcap, size_t budget) {
  if (Bi <= 0 || Bc <= 0 || D <= 0 || S <= 0 || Lw <= 0 || Lcap <= 0) return 0;
  const size_t fixed = fixed_bytes(Bc, D, Lw), per = per_caption_bytes(Bi, D, S, Lcap);
  size_t want = fixed + per * (size_t)Bc;
  if (budget != 0 && want > budget) {
    size_t nc = budget > fixed + per ? (budget - fixed) / per : 1;
    if (nc < 1) nc = 1;
    want = fixed + per * nc;
  }
  return want;
}

extern "C" int gloria_b200_local_sim_fwd_f32(const float* ctx, const float* words, const int32_t* cap_lens, int Bi,
                                             int Bc, int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                             float temp2, int agg, float eps, float* sim, float* attn_diag,
                                             float* attn_mean, void* workspace, size_t workspace_bytes,
                                             void* stream) {
  int rc = check_common(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, agg);
  if (rc) return rc;
  GLORIA_CHECK_ARG(sim != nullptr && workspace != nullptr, "null output / workspace");
  GLORIA_CHECK_ARG(attn_diag == nullptr || Bi == Bc, "attn_diag needs Bi == Bc (got %d x %d)", Bi, Bc);
  cudaStream_t st = (cudaStream_t)stream;
  const Plan pl = make_plan(Bi, Bc, D, S, Lw, Lcap, workspace_bytes);
  if (pl.nc < 1) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B too small", workspace_bytes);
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  float* wn = (float*)(ws + pl.off_wn);
  float* sc = (float*)(ws + pl.off_sc);
  float* at = (float*)(ws + pl.off_at);
  float* cx = (float*)(ws + pl.off_cx);
  if ((rc = prepack_words(words, Wt, wn, Bc, D, Lw, st))) return rc;
  for (int i0 = 0; i0 < Bc; i0 += pl.nc) {
    const int nc = min(pl.nc, Bc - i0);
    if ((rc = chunk_forward(ctx, Wt, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, temp1, i0, nc, sc, at, cx,
                            attn_diag, attn_mean, st)))
      return rc;
    cosine_agg<<<(unsigned)(Bi * nc), 256, 3 * Lcap * sizeof(float), st>>>(
        cx, Wt, wn, cap_lens, i0, nc, Bc, Lcap, Lw, word_off, D, temp2, agg, eps, sim, nullptr, nullptr);
    GLORIA_LAUNCHED("cosine_agg");
  }
  return GLORIA_OK;
}

extern "C" int gloria_b200_local_sim_bwd_f32(const float* ctx, const float* words, const int32_t* cap_lens, int Bi,
                                             int Bc, int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                             float temp2, int agg, float eps, const float* dsim,
                                             const float* d_attn_diag, const float* d_attn_mean, float* d_ctx,
                                             float* d_words, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, agg);
  if (rc) return rc;
  GLORIA_CHECK_ARG(dsim && d_ctx && d_words && workspace, "null gradient / workspace pointer");
  GLORIA_CHECK_ARG(d_attn_diag == nullptr || Bi == Bc, "d_attn_diag needs Bi == Bc (got %d x %d)", Bi, Bc);
  if (agg == GLORIA_AGG_MAX) return fail(GLORIA_ERR_UNSUPPORTED, "backward of agg=max is not part of the path");
  cudaStream_t st = (cudaStream_t)stream;
  const Plan pl = make_plan(Bi, Bc, D, S, Lw, Lcap, workspace_bytes);
  if (pl.nc < 1) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B too small", workspace_bytes);
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  float* dWt = (float*)(ws + pl.off_dwt);
  float* wn = (float*)(ws + pl.off_wn);
  float* sc = (float*)(ws + pl.off_sc);
  float* at = (float*)(ws + pl.off_at);
  float* da = (float*)(ws + pl.off_da);
  float* cx = (float*)(ws + pl.off_cx);
  float* coef = (float*)(ws + pl.off_coef);
  if ((rc = prepack_words(words, Wt, wn, Bc, D, Lw, st))) return rc;
  GLORIA_CUDA(cudaMemsetAsync(d_ctx, 0, (size_t)Bi * D * S * sizeof(float), st));
  GLORIA_CUDA(cudaMemsetAsync(dWt, 0, (size_t)Bc * Lw * D * sizeof(float), st));
  for (int i0 = 0; i0 < Bc; i0 += pl.nc) {
    const int nc = min(pl.nc, Bc - i0);
    // recompute forward intermediates of the chunk
    if ((rc = chunk_forward(ctx, Wt, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, temp1, i0, nc, sc, at, cx, nullptr,
                            nullptr, st)))
      return rc;
    cosine_agg<<<(unsigned)(Bi * nc), 256, 3 * Lcap * sizeof(float), st>>>(
        cx, Wt, wn, cap_lens, i0, nc, Bc, Lcap, Lw, word_off, D, temp2, agg, eps, nullptr, dsim, coef);
    GLORIA_LAUNCHED("cosine_agg(bwd)");
    bwd_context_grad<<<dim3((unsigned)Lcap, (unsigned)nc), 256, 0, st>>>(cx, Wt, dWt, coef, cap_lens, i0, nc, Bi,
                                                                          Lcap, Lw, word_off, D);
    GLORIA_LAUNCHED("bwd_context_grad");
    {  // dA[p][l][s] = sum_d dC[p][l][d] * ctx[j][d][s]
      GemmArgs g{};
      g.A = cx; g.B = ctx; g.C = da;
      g.M = Lcap; g.N = S; g.K = D; g.R = 1;
      g.sAm = D; g.sAk = 1; g.sAb0 = (long long)nc * Lcap * D; g.sAb1 = (long long)Lcap * D;
      g.sBk = S; g.sBn = 1; g.sBb0 = (long long)D * S; g.sBb1 = 0;
      g.sCm = S; g.sCn = 1; g.sCb0 = (long long)nc * Lcap * S; g.sCb1 = (long long)Lcap * S;
      g.nb0 = Bi; g.nb1 = nc; g.beta = 0.f; g.mlim = cap_lens; g.mlim_off = i0;
      if ((rc = launch_gemm(g, st))) return rc;
    }
    {  // d_ctx[j][d][s] += sum_{i in chunk} sum_l dC[p][l][d] * A[p][l][s]
      GemmArgs g{};
      g.A = cx; g.B = at; g.C = d_ctx;
      g.M = D; g.N = S; g.K = Lcap; g.R = nc;
      g.sAm = 1; g.sAk = D; g.sAb0 = (long long)nc * Lcap * D; g.sAb1 = 0; g.sAr = (long long)Lcap * D;
      g.sBk = S; g.sBn = 1; g.sBb0 = (long long)nc * Lcap * S; g.sBb1 = 0; g.sBr = (long long)Lcap * S;
      g.sCm = S; g.sCn = 1; g.sCb0 = (long long)D * S; g.sCb1 = 0;
      g.nb0 = Bi; g.nb1 = 1; g.beta = 1.f; g.klim = cap_lens; g.klim_off = i0;
      if ((rc = launch_gemm(g, st))) return rc;
    }
    double_softmax_bwd<<<(unsigned)(Bi * nc), 256, 0, st>>>(da, at, sc, cap_lens, i0, nc, Bc, Lcap, S, temp1,
                                                           d_attn_diag, d_attn_mean, 0);
    GLORIA_LAUNCHED("double_softmax_bwd");
    {  // dWt[i][off+l][d] += sum_j sum_s dS[p][l][s] * ctx[j][d][s]
      GemmArgs g{};
      g.A = da; g.B = ctx; g.C = dWt + ((long long)i0 * Lw + word_off) * D;
      g.M = Lcap; g.N = D; g.K = S; g.R = Bi;
      g.sAm = S; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)Lcap * S; g.sAr = (long long)nc * Lcap * S;
      g.sBk = 1; g.sBn = S; g.sBb0 = 0; g.sBb1 = 0; g.sBr = (long long)D * S;
      g.sCm = D; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)Lw * D;
      g.nb0 = 1; g.nb1 = nc; g.beta = 1.f; g.mlim = cap_lens; g.mlim_off = i0;
      if ((rc = launch_gemm(g, st))) return rc;
    }
    {  // d_ctx[j][d][s] += sum_{i in chunk} sum_l Wt[i][off+l][d] * dS[p][l][s]
      GemmArgs g{};
      g.A = Wt + ((long long)i0 * Lw + word_off) * D; g.B = da; g.C = d_ctx;
      g.M = D; g.N = S; g.K = Lcap; g.R = nc;
      g.sAm = 1; g.sAk = D; g.sAb0 = 0; g.sAb1 = 0; g.sAr = (long long)Lw * D;
      g.sBk = S; g.sBn = 1; g.sBb0 = (long long)nc * Lcap * S; g.sBb1 = 0; g.sBr = (long long)Lcap * S;
      g.sCm = S; g.sCn = 1; g.sCb0 = (long long)D * S; g.sCb1 = 0;
      g.nb0 = Bi; g.nb1 = 1; g.beta = 1.f; g.klim = cap_lens; g.klim_off = i0;
      if ((rc = launch_gemm(g, st))) return rc;
    }
  }
  dim3 grid((Lw + 31) / 32, (D + 31) / 32, Bc), block(32, 8);
  unpack_dwords<<<grid, block, 0, st>>>(dWt, d_words, cap_lens, D, Lw, Lcap, word_off, 0);
  GLORIA_LAUNCHED("unpack_dwords");
  return GLORIA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Diagonal pairs only: attention maps A_ii (att_maps, gloria_loss.py:141-143) and the gradient that flows back
// through them (supervised-attention loss, gloria_model.py:143-147).  B pairs instead of B^2.
// workspace: Wt [B,Lw,D], dWt [B,Lw,D], wn [B,Lw], sc / at / da [B,Lcap,S]
// ---------------------------------------------------------------------------------------------------------------
namespace gloria {
namespace {
struct DiagPlan { size_t off_wt, off_dwt, off_wn, off_sc, off_at, off_da, off_tc, tc_bytes, total; };
DiagPlan diag_plan(int B, int D, int S, int Lw, int Lcap) {
  DiagPlan pl{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += gloria::align_up(n, 256); return r; };
  pl.off_wt = take((size_t)B * Lw * D * sizeof(float));
  pl.off_dwt = take((size_t)B * Lw * D * sizeof(float));
  pl.off_wn = take((size_t)B * Lw * sizeof(float));
  pl.off_sc = take((size_t)B * Lcap * S * sizeof(float));
  pl.off_at = take((size_t)B * Lcap * S * sizeof(float));
  pl.off_da = take((size_t)B * Lcap * S * sizeof(float));
  pl.tc_bytes = gloria::f32tc_diag_scores_workspace(B, D, S, Lcap);      // operand pieces of the tensor-core score GEMM
  pl.off_tc = take(pl.tc_bytes);
  pl.total = o;
  return pl;
}

// scores + double softmax of the pairs (i, i); leaves P in sc and A in at (and attn_diag if given)
int diag_forward(const float* ctx, const float* Wt, const int32_t* cap_lens, int B, int D, int S, int Lw, int Lcap,
                 int off, float temp1, float* sc, float* at, float* attn_diag, void* tc_ws, size_t tc_bytes, cudaStream_t st) {
  // S_[i][l][s] = sum_d Wt[i][off+l][d] ctx[i][d][s]: one batch per pair on the split-precision tensor-core GEMM
  // (tc_f32.cu; fp32 accuracy) where the shape is covered, else on the CUDA-core GEMM above
  int rc;
  if (tc_bytes > 0) {
    if ((rc = f32tc_diag_scores(ctx, Wt, cap_lens, B, D, S, Lw, Lcap, off, sc, tc_ws, tc_bytes, st))) return rc;
  } else {
    GemmArgs g{};
    g.A = Wt + (long long)off * D; g.B = ctx; g.C = sc;
    g.M = Lcap; g.N = S; g.K = D; g.R = 1;
    g.sAm = D; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)Lw * D; g.sAr = 0;
    g.sBk = S; g.sBn = 1; g.sBb0 = 0; g.sBb1 = (long long)D * S; g.sBr = 0;
    g.sCm = S; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)Lcap * S;
    g.nb0 = 1; g.nb1 = B; g.beta = 0.f; g.mlim = cap_lens; g.mlim_off = 0;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  double_softmax_fwd<<<(unsigned)B, 256, 0, st>>>(sc, at, cap_lens, 0, B, B, Lcap, S, temp1, attn_diag, nullptr, 1);
  GLORIA_LAUNCHED("double_softmax_fwd(diag)");
  return GLORIA_OK;
}
}  // namespace
}  // namespace gloria

extern "C" size_t gloria_b200_diag_attn_workspace(int B, int D, int S, int Lw, int Lcap) {
  if (B <= 0 || D <= 0 || S <= 0 || Lw <= 0 || Lcap <= 0) return 0;
  return diag_plan(B, D, S, Lw, Lcap).total;
}

extern "C" int gloria_b200_diag_attn_fwd_f32(const float* ctx, const float* words, const int32_t* cap_lens, int B,
                                             int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                             float* attn_diag, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(ctx, words, cap_lens, B, B, D, S, Lw, Lcap, word_off, GLORIA_AGG_SUM);
  if (rc) return rc;
  GLORIA_CHECK_ARG(attn_diag && workspace, "null output / workspace");
  const DiagPlan pl = diag_plan(B, D, S, Lw, Lcap);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  if ((rc = prepack_words(words, Wt, (float*)(ws + pl.off_wn), B, D, Lw, st))) return rc;
  return diag_forward(ctx, Wt, cap_lens, B, D, S, Lw, Lcap, word_off, temp1, (float*)(ws + pl.off_sc),
                      (float*)(ws + pl.off_at), attn_diag, ws + pl.off_tc, pl.tc_bytes, st);
}

extern "C" int gloria_b200_diag_attn_bwd_f32(const float* ctx, const float* words, const int32_t* cap_lens, int B,
                                             int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                             const float* d_attn_diag, float* d_ctx, float* d_words, int accumulate,
                                             void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(ctx, words, cap_lens, B, B, D, S, Lw, Lcap, word_off, GLORIA_AGG_SUM);
  if (rc) return rc;
  GLORIA_CHECK_ARG(d_attn_diag && d_ctx && d_words && workspace, "null gradient / workspace pointer");
  const DiagPlan pl = diag_plan(B, D, S, Lw, Lcap);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  float* dWt = (float*)(ws + pl.off_dwt);
  float* sc = (float*)(ws + pl.off_sc);
  float* at = (float*)(ws + pl.off_at);
  float* da = (float*)(ws + pl.off_da);
  if ((rc = prepack_words(words, Wt, (float*)(ws + pl.off_wn), B, D, Lw, st))) return rc;
  if ((rc = diag_forward(ctx, Wt, cap_lens, B, D, S, Lw, Lcap, word_off, temp1, sc, at, nullptr, ws + pl.off_tc, pl.tc_bytes, st)))
    return rc;
  GLORIA_CUDA(cudaMemsetAsync(da, 0, (size_t)B * Lcap * S * sizeof(float), st));
  GLORIA_CUDA(cudaMemsetAsync(dWt, 0, (size_t)B * Lw * D * sizeof(float), st));
  double_softmax_bwd<<<(unsigned)B, 256, 0, st>>>(da, at, sc, cap_lens, 0, B, B, Lcap, S, temp1, d_attn_diag, nullptr, 1);
  GLORIA_LAUNCHED("double_softmax_bwd(diag)");
  {  // dWt[i][off+l][d] = sum_s dS[i][l][s] * ctx[i][d][s]
    GemmArgs g{};
    g.A = da; g.B = ctx; g.C = dWt + (long long)word_off * D;
    g.M = Lcap; g.N = D; g.K = S; g.R = 1;
    g.sAm = S; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)Lcap * S; g.sAr = 0;
    g.sBk = 1; g.sBn = S; g.sBb0 = 0; g.sBb1 = (long long)D * S; g.sBr = 0;
    g.sCm = D; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)Lw * D;
    g.nb0 = 1; g.nb1 = B; g.beta = 0.f; g.mlim = cap_lens; g.mlim_off = 0;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  {  // d_ctx[i][d][s] (+)= sum_l Wt[i][off+l][d] * dS[i][l][s]
    GemmArgs g{};
    g.A = Wt + (long long)word_off * D; g.B = da; g.C = d_ctx;
    g.M = D; g.N = S; g.K = Lcap; g.R = 1;
    g.sAm = 1; g.sAk = D; g.sAb0 = (long long)Lw * D; g.sAb1 = 0; g.sAr = 0;
    g.sBk = S; g.sBn = 1; g.sBb0 = (long long)Lcap * S; g.sBb1 = 0; g.sBr = 0;
    g.sCm = S; g.sCn = 1; g.sCb0 = (long long)D * S; g.sCb1 = 0;
    g.nb0 = B; g.nb1 = 1; g.beta = accumulate ? 1.f : 0.f;
    // K limit per image: caption i's length (klim is indexed by r; R == 1, so offset by the image index instead)
    g.klim = nullptr;   // rows l >= cap_len of dS are zero (double_softmax_bwd never writes them; da was cleared)
    if ((rc = launch_gemm(g, st))) return rc;
  }
  dim3 grid((Lw + 31) / 32, (D + 31) / 32, B), block(32, 8);
  unpack_dwords<<<grid, block, 0, st>>>(dWt, d_words, cap_lens, D, Lw, Lcap, word_off, accumulate);
  GLORIA_LAUNCHED("unpack_dwords(diag)");
  return GLORIA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// attention_fn (gloria_loss.py:19-63) as a stand-alone paired operator: query[b] attends to context[b].
//   query [B, D, L] (word axis contiguous), context [B, D, S]  ->  wctx [B, D, L], attn [B, L, S]
// Backward takes d_wctx [B, D, L] and d_attn [B, L, S] (either may be NULL) and overwrites d_query, d_ctx.
// workspace: Wt, dWt [B,L,D]; wn [B,L]; sc, at, da [B,L,S]; cx [B,L,D]
// ---------------------------------------------------------------------------------------------------------------
namespace gloria {
namespace {
struct AttnPlan { size_t off_wt, off_dwt, off_wn, off_sc, off_at, off_da, off_cx, off_lens, off_tc, tc_bytes, total; };
AttnPlan attn_plan(int B, int D, int S, int L) {
  AttnPlan pl{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += gloria::align_up(n, 256); return r; };
  pl.off_wt = take((size_t)B * L * D * 4);
  pl.off_dwt = take((size_t)B * L * D * 4);
  pl.off_wn = take((size_t)B * L * 4);
  pl.off_sc = take((size_t)B * L * S * 4);
  pl.off_at = take((size_t)B * L * S * 4);
  pl.off_da = take((size_t)B * L * S * 4);
  pl.off_cx = take((size_t)B * L * D * 4);
  pl.off_lens = take((size_t)B * 4);
  pl.tc_bytes = gloria::f32tc_diag_scores_workspace(B, D, S, L);
  pl.off_tc = take(pl.tc_bytes);
  pl.total = o;
  return pl;
}
__global__ void fill_int(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// [B, L, D] -> [B, D, L]
__global__ void transpose_ld_to_dl(const float* __restrict__ in, float* __restrict__ out, int D, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const float* ib = in + (long long)b * D * L;
  float* ob = out + (long long)b * D * L;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    t[r][threadIdx.x] = (l < L && d < D) ? ib[(long long)l * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, l = l0 + threadIdx.x;
    if (d < D && l < L) ob[(long long)d * L + l] = t[threadIdx.x][r];
  }
}
// cx[b][l][d] = sum_s at[b][l][s] * ctx[b][d][s]   (paired)
int paired_context(const float* at, const float* ctx, float* cx, int B, int D, int S, int L, cudaStream_t st) {
  GemmArgs g{};
  g.A = at; g.B = ctx; g.C = cx;
  g.M = L; g.N = D; g.K = S; g.R = 1;
  g.sAm = S; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)L * S;
  g.sBk = 1; g.sBn = S; g.sBb0 = 0; g.sBb1 = (long long)D * S;
  g.sCm = D; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)L * D;
  g.nb0 = 1; g.nb1 = B; g.beta = 0.f;
  return launch_gemm(g, st);
}
}  // namespace
}  // namespace gloria

extern "C" size_t gloria_b200_attention_workspace(int B, int D, int S, int L) {
  if (B <= 0 || D <= 0 || S <= 0 || L <= 0) return 0;
  return attn_plan(B, D, S, L).total;
}

extern "C" int gloria_b200_attention_fwd_f32(const float* query, const float* ctx, int B, int D, int S, int L,
                                             float temp1, float* wctx, float* attn, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  GLORIA_CHECK_ARG(query && ctx && wctx && attn && workspace, "null pointer");
  GLORIA_CHECK_ARG(B > 0 && D > 0 && S > 0 && L > 0, "non-positive size");
  const AttnPlan pl = attn_plan(B, D, S, L);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  float* cx = (float*)(ws + pl.off_cx);
  int* lens = (int*)(ws + pl.off_lens);
  int rc;
  fill_int<<<(B + 255) / 256, 256, 0, st>>>(lens, B, L);
  GLORIA_LAUNCHED("fill_int");
  if ((rc = prepack_words(query, Wt, (float*)(ws + pl.off_wn), B, D, L, st))) return rc;
  if ((rc = diag_forward(ctx, Wt, lens, B, D, S, L, L, 0, temp1, (float*)(ws + pl.off_sc), (float*)(ws + pl.off_at),
                         attn, ws + pl.off_tc, pl.tc_bytes, st)))
    return rc;
  if ((rc = paired_context((float*)(ws + pl.off_at), ctx, cx, B, D, S, L, st))) return rc;
  dim3 grid((L + 31) / 32, (D + 31) / 32, B), block(32, 8);
  transpose_ld_to_dl<<<grid, block, 0, st>>>(cx, wctx, D, L);
  GLORIA_LAUNCHED("transpose_ld_to_dl");
  return GLORIA_OK;
}

extern "C" int gloria_b200_attention_bwd_f32(const float* query, const float* ctx, int B, int D, int S, int L,
                                             float temp1, const float* d_wctx, const float* d_attn, float* d_query,
                                             float* d_ctx, void* workspace, size_t workspace_bytes, void* stream) {
  GLORIA_CHECK_ARG(query && ctx && d_query && d_ctx && workspace, "null pointer");
  GLORIA_CHECK_ARG(B > 0 && D > 0 && S > 0 && L > 0, "non-positive size");
  const AttnPlan pl = attn_plan(B, D, S, L);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* Wt = (float*)(ws + pl.off_wt);
  float* dWt = (float*)(ws + pl.off_dwt);
  float* sc = (float*)(ws + pl.off_sc);
  float* at = (float*)(ws + pl.off_at);
  float* da = (float*)(ws + pl.off_da);
  float* dCt = (float*)(ws + pl.off_cx);       // d_wctx transposed to [B, L, D]
  int* lens = (int*)(ws + pl.off_lens);
  int rc;
  fill_int<<<(B + 255) / 256, 256, 0, st>>>(lens, B, L);
  GLORIA_LAUNCHED("fill_int");
  if ((rc = prepack_words(query, Wt, (float*)(ws + pl.off_wn), B, D, L, st))) return rc;
  if ((rc = diag_forward(ctx, Wt, lens, B, D, S, L, L, 0, temp1, sc, at, nullptr, ws + pl.off_tc, pl.tc_bytes, st))) return rc;
  GLORIA_CUDA(cudaMemsetAsync(da, 0, (size_t)B * L * S * sizeof(float), st));
  GLORIA_CUDA(cudaMemsetAsync(d_ctx, 0, (size_t)B * D * S * sizeof(float), st));
  if (d_wctx) {
    dim3 grid((L + 31) / 32, (D + 31) / 32, B), block(32, 8);
    transpose_dl_to_ld<<<grid, block, 0, st>>>(d_wctx, dCt, D, L);
    GLORIA_LAUNCHED("transpose_dl_to_ld");
    {  // dA[b][l][s] = sum_d dC[b][l][d] * ctx[b][d][s]
      GemmArgs g{};
      g.A = dCt; g.B = ctx; g.C = da;
      g.M = L; g.N = S; g.K = D; g.R = 1;
      g.sAm = D; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)L * D;
      g.sBk = S; g.sBn = 1; g.sBb0 = 0; g.sBb1 = (long long)D * S;
      g.sCm = S; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)L * S;
      g.nb0 = 1; g.nb1 = B; g.beta = 0.f;
      if ((rc = launch_gemm(g, st))) return rc;
    }
    {  // d_ctx[b][d][s] = sum_l dC[b][l][d] * A[b][l][s]
      GemmArgs g{};
      g.A = dCt; g.B = at; g.C = d_ctx;
      g.M = D; g.N = S; g.K = L; g.R = 1;
      g.sAm = 1; g.sAk = D; g.sAb0 = (long long)L * D; g.sAb1 = 0;
      g.sBk = S; g.sBn = 1; g.sBb0 = (long long)L * S; g.sBb1 = 0;
      g.sCm = S; g.sCn = 1; g.sCb0 = (long long)D * S; g.sCb1 = 0;
      g.nb0 = B; g.nb1 = 1; g.beta = 0.f;
      if ((rc = launch_gemm(g, st))) return rc;
    }
  }
  // softmax backward; d_attn (if any) is added to dA inside the kernel
  double_softmax_bwd<<<(unsigned)B, 256, 0, st>>>(da, at, sc, lens, 0, B, B, L, S, temp1, d_attn, nullptr, 1);
  GLORIA_LAUNCHED("double_softmax_bwd(attention)");
  {  // dWt[b][l][d] = sum_s dS[b][l][s] * ctx[b][d][s]
    GemmArgs g{};
    g.A = da; g.B = ctx; g.C = dWt;
    g.M = L; g.N = D; g.K = S; g.R = 1;
    g.sAm = S; g.sAk = 1; g.sAb0 = 0; g.sAb1 = (long long)L * S;
    g.sBk = 1; g.sBn = S; g.sBb0 = 0; g.sBb1 = (long long)D * S;
    g.sCm = D; g.sCn = 1; g.sCb0 = 0; g.sCb1 = (long long)L * D;
    g.nb0 = 1; g.nb1 = B; g.beta = 0.f;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  {  // d_ctx[b][d][s] += sum_l Wt[b][l][d] * dS[b][l][s]
    GemmArgs g{};
    g.A = Wt; g.B = da; g.C = d_ctx;
    g.M = D; g.N = S; g.K = L; g.R = 1;
    g.sAm = 1; g.sAk = D; g.sAb0 = (long long)L * D; g.sAb1 = 0;
    g.sBk = S; g.sBn = 1; g.sBb0 = (long long)L * S; g.sBb1 = 0;
    g.sCm = S; g.sCn = 1; g.sCb0 = (long long)D * S; g.sCb1 = 0;
    g.nb0 = B; g.nb1 = 1; g.beta = 1.f;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  dim3 grid((L + 31) / 32, (D + 31) / 32, B), block(32, 8);
  transpose_ld_to_dl<<<grid, block, 0, st>>>(dWt, d_query, D, L);
  GLORIA_LAUNCHED("transpose_ld_to_dl");
  return GLORIA_OK;
}

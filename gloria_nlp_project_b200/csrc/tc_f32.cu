// fp32 mode of the GLoRIA local similarity ON THE TENSOR CORES (sm_100a): every bmm of the reference op graph
// (gloria/loss/gloria_loss.py:40,59 and their autograd) runs on the CTA-pair tcgen05 GEMM of tc_gemm.cu with
// SPLIT-PRECISION operands, the softmaxes / cosine / aggregation are streaming fp32 kernels in between.
//
// Split precision: an fp32 value x is carried as three bf16 pieces x = p0 + p1 + p2 (p0 = bf16(x), p1 = bf16(x - p0),
// p2 = bf16(x - p0 - p1): 24 significant bits, every piece product is exact in the fp32 accumulator).  A GEMM sums
//   6 terms  p2.q0 + p1.q1 + p0.q2 + p1.q0 + p0.q1 + p0.q0      (everything down to 2^-24: the forward, 1e-5 gate) or
//   3 terms  p1.q0 + p0.q1 + p0.q0                              (2^-16: gradient GEMMs)
// as runs of k-blocks into ONE accumulator, smallest term first -- tensor memory accumulates round-toward-zero
// (profiles/r02_tmem_accumulation_rounding_probe.txt), so the large term must come last: K = 768 then lands at
// 9e-7 rms of the exact result, the same as an fp32 FFMA loop.
//
// Layout (chunk of nc captions [i0, i0+nc), all Bi images; Sq = round_up(S, 64), Lp = round_up(Lcap, 8),
// NC = round_up(nc * Lp, 64)): the big matrices have rows (j, s) and columns (ii, l), like X^T of the bf16 backward.
//   Rt_p  [3][Bi*Sq, D]  bf16   region features, rows (j, s)            (A of the score GEMM, B of GEMMs 2 and 5)
//   Rn_p  [3][Bi*D, Sq]  bf16   region features, rows (j, d)            (B of GEMM 3)
//   Wn_p  [3][D, NC]     bf16   words of the chunk, columns (ii, l)     (B of the score GEMM)
//   Wt_p  [3][NC, D]     bf16   words of the chunk, rows (ii, l)        (B of GEMM 6)
//   SC    [Bi*Sq, NC]    fp32   scores, then P (word softmax)           GEMM 1:  SC = Rt Wn
//   AT_p  [3][Bi*Sq, NC] bf16   attention A (region softmax)
//   CX    [Bi][NC, D]    fp32   context                                 GEMM 2:  CX_j = AT_j^T Rt_j          (batched over j)
//   dC_p  [3][Bi*NC, D]  bf16   dL/dC                                   (cosine backward)
//   DAt   [Bi][NC, Sq]   fp32   dL/dA, transposed blocks                GEMM 3:  DAt_j = dC_j Rn_j           (batched)
//   dRt   [Bi*Sq, D]     fp32   region gradient                         GEMM 4:  dRt_j += AT_j dC_j          (batched)
//   DS_p  [3][Bi*Sq, NC] bf16   dL/dscores                              (softmax backward)
//   dWc   [NC, D]        fp32   word gradient of the chunk              GEMM 5:  dWc = DS^T Rt
//                                                                       GEMM 6:  dRt += DS Wt
// Everything inside the tensor-map extents is written (padding = exact zeros): the GEMM reads whole tiles.
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gloria {
namespace f32tc {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split3(float x, bf16& p0, bf16& p1, bf16& p2) {
  p0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(p1);
  p2 = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ float join3(bf16 p0, bf16 p1, bf16 p2) {
  return (__bfloat162float(p2) + __bfloat162float(p1)) + __bfloat162float(p0);     // exact: the pieces do not overlap
}

// ctx [Bi, D, S] fp32 -> Rt_p [3][Bi*Sq, D] (transposed; rows s >= S zero) and Rn_p [3][Bi*D, Sq] (columns s >= S zero; or null)
__global__ void split_ctx(const float* __restrict__ ctx, bf16* __restrict__ rt, bf16* __restrict__ rn, int Bi, int D, int S, int Sq) {
  __shared__ float t[32][33];
  const int j = blockIdx.z, s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const size_t rt_plane = (size_t)Bi * Sq * D, rn_plane = (size_t)Bi * D * Sq;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, s = s0 + threadIdx.x;
    const float v = (s < S) ? ctx[((size_t)j * D + d) * S + s] : 0.f;
    t[r][threadIdx.x] = v;
    if (rn != nullptr) {
      bf16 p0, p1, p2;
      split3(v, p0, p1, p2);
      const size_t o = ((size_t)j * D + d) * Sq + s;
      rn[o] = p0; rn[rn_plane + o] = p1; rn[2 * rn_plane + o] = p2;
    }
  }
  if (rt == nullptr) return;
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = s0 + r, d = d0 + threadIdx.x;
    bf16 p0, p1, p2;
    split3(t[threadIdx.x][r], p0, p1, p2);
    const size_t o = ((size_t)j * Sq + s) * D + d;
    rt[o] = p0; rt[rt_plane + o] = p1; rt[2 * rt_plane + o] = p2;
  }
}

// words [Bc, D, Lw] -> Wt32 [Bc, Lw, D]
__global__ void transpose_words(const float* __restrict__ in, float* __restrict__ out, int D, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const float* ib = in + (size_t)b * D * L;
  float* ob = out + (size_t)b * D * L;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, l = l0 + threadIdx.x;
    t[r][threadIdx.x] = (d < D && l < L) ? ib[(size_t)d * L + l] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    if (l < L && d < D) ob[(size_t)l * D + d] = t[threadIdx.x][r];
  }
}
__global__ void word_norms(const float* __restrict__ x, float* __restrict__ n, long long rows, int D) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
  s = warp_sum(s);
  if (lane == 0) n[row] = sqrtf(s);
}

// Wt32 rows of the chunk -> Wt_p [3][NC, D] (or null) and Wn_p [3][D, NC]; column c = ii * Lp + l, zero beyond the caption
__global__ void split_words(const float* __restrict__ wt32, const int* __restrict__ cap_lens, bf16* __restrict__ wnp,
                            bf16* __restrict__ wtp, int i0, int nc, int Lw, int Lcap, int Lp, int off, int D, int NC) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const size_t plane = (size_t)NC * D;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, d = d0 + threadIdx.x;
    const int ii = c / Lp, l = c - ii * Lp;
    float v = 0.f;
    if (ii < nc) {
      const int L = min(max(cap_lens[i0 + ii], 0), Lcap);
      if (l < L) v = wt32[((size_t)(i0 + ii) * Lw + off + l) * D + d];
    }
    t[r][threadIdx.x] = v;
    if (wtp != nullptr) {
      bf16 p0, p1, p2;
      split3(v, p0, p1, p2);
      const size_t o = (size_t)c * D + d;
      wtp[o] = p0; wtp[plane + o] = p1; wtp[2 * plane + o] = p2;
    }
  }
  if (wnp == nullptr) return;
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, c = c0 + threadIdx.x;
    bf16 p0, p1, p2;
    split3(t[threadIdx.x][r], p0, p1, p2);
    const size_t o = (size_t)d * NC + c;
    wnp[o] = p0; wnp[plane + o] = p1; wnp[2 * plane + o] = p2;
  }
}

// tmp [B][Lp][Sq] -> sc [B][Lcap][S]  (scores of the diagonal pairs in the layout of simt_f32.cu's diagonal kernels)
__global__ void repack_diag_scores(const float* __restrict__ tmp, float* __restrict__ sc, int Lcap, int Lp, int S, int Sq) {
  const int l = blockIdx.x, i = blockIdx.y;
  const float* src = tmp + ((size_t)i * Lp + l) * Sq;
  float* dst = sc + ((size_t)i * Lcap + l) * S;
  for (int x = threadIdx.x; x < S; x += blockDim.x) dst[x] = src[x];
}

// Double softmax of one (image, caption) block (gloria_loss.py:42-53).  SC block: scores on entry, P (word softmax) on exit.
// AT_p block: A = softmax_s(temp1 P) as bf16 pieces, zero in padded rows / columns.  512 threads;
// dynamic smem: S * (Lp + 1) + 5 * 128 floats
constexpr int SM_THREADS = 512;
constexpr int SM_PARTS = SM_THREADS / 128;      // threads per word in the column passes
constexpr int RB = 4;                           // region rows a warp keeps in flight
__global__ void __launch_bounds__(SM_THREADS) softmax_fwd(float* __restrict__ sc, bf16* __restrict__ atp,
                                                          const int* __restrict__ cap_lens, int i0, int nc, int Bi, int Bc, int S,
                                                          int Sq, int Lcap, int Lp, int NC, float temp1,
                                                          float* __restrict__ attn_diag, float* __restrict__ attn_mean) {
  extern __shared__ float sm[];
  const int LP1 = Lp + 1;
  float* tile = sm;                          // E[s][l]
  float* zpart = sm + (size_t)S * LP1;       // [SM_PARTS][128]
  float* invz = zpart + SM_PARTS * 128;      // [128]
  const int p = blockIdx.x, j = p / nc, ii = p - j * nc, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SM_THREADS >> 5;
  float* scb = sc + (size_t)j * Sq * NC + (size_t)ii * Lp;
  // softmax #1 over the caption's words: a warp owns RB region rows at a time (coalesced along l, <= 4 words per lane)
  for (int s0 = warp * RB; s0 < S; s0 += nwarps * RB) {
    float v[RB][4], m[RB], den[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float* row = scb + (size_t)(s0 + r) * NC;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int l = lane + 32 * k;
        v[r][k] = (s0 + r < S && l < L) ? row[l] : -INFINITY;
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) m[r] = warp_max(fmaxf(fmaxf(v[r][0], v[r][1]), fmaxf(v[r][2], v[r][3])));
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float e = (v[r][k] == -INFINITY) ? 0.f : expf(v[r][k] - m[r]);
        v[r][k] = e;
        d += e;
      }
      den[r] = warp_sum(d);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (s0 + r < S) {
        float* row = scb + (size_t)(s0 + r) * NC;
        const float inv = 1.f / den[r];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int l = lane + 32 * k;
          if (l < Lp) {
            const float P = (l < L) ? v[r][k] * inv : 0.f;
            row[l] = P;
            tile[(size_t)(s0 + r) * LP1 + l] = (l < L) ? expf(temp1 * P) : 0.f;
          }
        }
      }
    }
  }
  __syncthreads();
  // softmax #2 over the regions: Z_l (SM_PARTS threads per word, rows interleaved)
  {
    const int l = tid & 127, part = tid >> 7;
    float z0 = 0.f, z1 = 0.f;
    if (l < L) {
      int s = part;
      for (; s + SM_PARTS < S; s += 2 * SM_PARTS) {
        z0 += tile[(size_t)s * LP1 + l];
        z1 += tile[(size_t)(s + SM_PARTS) * LP1 + l];
      }
      if (s < S) z0 += tile[(size_t)s * LP1 + l];
    }
    zpart[part * 128 + l] = z0 + z1;
  }
  __syncthreads();
  if (tid < 128) {
    float z = 0.f;
#pragma unroll
    for (int q = 0; q < SM_PARTS; ++q) z += zpart[q * 128 + tid];
    invz[tid] = (tid < L) ? 1.f / z : 0.f;
  }
  __syncthreads();
  // A as bf16 pieces, two words per thread
  const size_t plane = (size_t)Bi * Sq * NC;
  const int half_lp = Lp >> 1;
  bf16* ab = atp + (size_t)j * Sq * NC + (size_t)ii * Lp;
  for (int s = warp; s < Sq; s += nwarps) {
    for (int h = lane; h < half_lp; h += 32) {
      const int l = 2 * h;
      float a0 = 0.f, a1 = 0.f;
      if (s < S) {
        a0 = tile[(size_t)s * LP1 + l] * invz[l];
        a1 = tile[(size_t)s * LP1 + l + 1] * invz[l + 1];
      }
      bf16 x0, x1, x2, y0, y1, y2;
      split3(a0, x0, x1, x2);
      split3(a1, y0, y1, y2);
      const size_t o = (size_t)s * NC + l;
      *reinterpret_cast<__nv_bfloat162*>(ab + o) = __nv_bfloat162(x0, y0);
      *reinterpret_cast<__nv_bfloat162*>(ab + plane + o) = __nv_bfloat162(x1, y1);
      *reinterpret_cast<__nv_bfloat162*>(ab + 2 * plane + o) = __nv_bfloat162(x2, y2);
    }
  }
  if (attn_diag != nullptr && j == i) {       // att_maps of the diagonal pair, [Bc, Lcap, S] (gloria_loss.py:141-143)
    float* dg = attn_diag + (size_t)i * Lcap * S;
    for (int l = warp; l < Lcap; l += nwarps) {
      const float iz = (l < L) ? invz[l] : 0.f;
      for (int x = lane; x < S; x += 32) dg[(size_t)l * S + x] = (l < L) ? tile[(size_t)x * LP1 + l] * iz : 0.f;
    }
  }
  if (attn_mean != nullptr) {                 // word-mean attention (gloria_loss.py:132), [Bi, Bc, S]
    float* mo = attn_mean + ((size_t)j * Bc + i) * S;
    const float invL = L > 0 ? 1.f / (float)L : 0.f;
    for (int x = tid; x < S; x += SM_THREADS) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc += tile[(size_t)x * LP1 + l] * invz[l];
      mo[x] = acc * invL;
    }
  }
}

// Per-word cosine (gloria_loss.py:11-16,150) + aggregation over words (:153-158); in backward mode also the per-word
// coefficients ddot, beta = dnc / nc, gamma = dnw / nw of the closed-form backward.  One CTA per pair; smem 3 * Lp floats.
__global__ void __launch_bounds__(256) cosine_agg(const float* __restrict__ cx, const float* __restrict__ wt32,
                                                  const float* __restrict__ wn, const int* __restrict__ cap_lens, int i0, int nc,
                                                  int Bc, int Lcap, int Lp, int NC, int Lw, int off, int D, float temp2, int agg,
                                                  float eps, float* __restrict__ sim, const float* __restrict__ dsim,
                                                  float* __restrict__ coef) {
  extern __shared__ float sm[];
  float* r_s = sm;
  float* dot_s = sm + Lp;
  float* nc_s = sm + 2 * Lp;
  const int p = blockIdx.x, j = p / nc, ii = p - j * nc, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int l = warp; l < L; l += nwarps) {
    const float4* w = reinterpret_cast<const float4*>(wt32 + ((size_t)i * Lw + off + l) * D);
    const float4* c = reinterpret_cast<const float4*>(cx + ((size_t)j * NC + (size_t)ii * Lp + l) * D);
    float dot = 0.f, c2 = 0.f;
#pragma unroll 6
    for (int d = lane; d < (D >> 2); d += 32) {        // D % 64 == 0: rows are 16-byte aligned
      const float4 cv = c[d], wv = w[d];
      dot = fmaf(wv.x, cv.x, dot); dot = fmaf(wv.y, cv.y, dot); dot = fmaf(wv.z, cv.z, dot); dot = fmaf(wv.w, cv.w, dot);
      c2 = fmaf(cv.x, cv.x, c2); c2 = fmaf(cv.y, cv.y, c2); c2 = fmaf(cv.z, cv.z, c2); c2 = fmaf(cv.w, cv.w, c2);
    }
    dot = warp_sum(dot);
    c2 = warp_sum(c2);
    if (lane == 0) {
      const float ncv = sqrtf(c2);
      const float den = fmaxf(wn[(size_t)i * Lw + off + l] * ncv, eps);
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
  if (lane == 0 && sim != nullptr) {
    float v;
    if (agg == GLORIA_AGG_MAX) v = temp2 * m;
    else {
      v = temp2 * m + logf(sum);
      if (agg == GLORIA_AGG_MEAN) v -= logf((float)L);
    }
    sim[(size_t)j * Bc + i] = v;
  }
  if (dsim == nullptr) return;
  const float g = dsim[(size_t)j * Bc + i];
  float* cf = coef + (size_t)p * 3 * Lp;
  for (int l = lane; l < Lp; l += 32) {
    float ddot = 0.f, beta = 0.f, gamma = 0.f;
    if (l < L) {
      const float q = expf(temp2 * (r_s[l] - m)) / sum;
      const float dr = g * temp2 * q;
      const float nwv = wn[(size_t)i * Lw + off + l], ncv = nc_s[l], dot = dot_s[l];
      const float prod = nwv * ncv;
      const float den = fmaxf(prod, eps);
      ddot = dr / den;
      const float dden = (prod >= eps) ? -dr * dot / (den * den) : 0.f;
      beta = ncv > 0.f ? dden * nwv / ncv : 0.f;      // (dL/d|C|) / |C|
      gamma = nwv > 0.f ? dden * ncv / nwv : 0.f;     // (dL/d|W|) / |W|
    }
    cf[l] = ddot;
    cf[Lp + l] = beta;
    cf[2 * Lp + l] = gamma;
  }
}

// dC = ddot W + beta C as bf16 pieces (zero rows beyond the caption), and the direct word gradient
// dWt32[i][off+l][:] = sum_j ddot C + gamma W.   grid (Lp, nc): one CTA owns word l of caption i0 + ii; a thread owns four
// channels and every ngroups-th image, the groups' partial sums meet in shared memory.  dynamic smem: blockDim float4.
constexpr int CG_THREADS = 512;
__global__ void __launch_bounds__(CG_THREADS) context_grad(const float* __restrict__ cx, const float* __restrict__ wt32,
                                                           float* __restrict__ dwt32, const float* __restrict__ coef,
                                                           const int* __restrict__ cap_lens, bf16* __restrict__ dcp, int i0, int nc,
                                                           int Bi, int Lcap, int Lp, int NC, int Lw, int off, int D) {
  extern __shared__ float4 part[];
  const int l = blockIdx.x, ii = blockIdx.y, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  const bool live = l < L;
  const int nq = D >> 2;                                   // float4 chunks per row
  const int ngroups = max(1, min(CG_THREADS / nq, Bi));
  const int dq = threadIdx.x % nq, jg = threadIdx.x / nq;
  const size_t plane = (size_t)Bi * NC * D;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (jg < ngroups) {
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) wv = reinterpret_cast<const float4*>(wt32 + ((size_t)i * Lw + off + l) * D)[dq];
#pragma unroll 4
    for (int j = jg; j < Bi; j += ngroups) {
      const size_t row = (size_t)j * NC + (size_t)ii * Lp + l;
      float4 dc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const float* cf = coef + ((size_t)j * nc + ii) * 3 * Lp;
        const float ddot = cf[l], beta = cf[Lp + l], gamma = cf[2 * Lp + l];
        const float4 cv = reinterpret_cast<const float4*>(cx + row * D)[dq];
        dc.x = fmaf(ddot, wv.x, beta * cv.x); dc.y = fmaf(ddot, wv.y, beta * cv.y);
        dc.z = fmaf(ddot, wv.z, beta * cv.z); dc.w = fmaf(ddot, wv.w, beta * cv.w);
        acc.x += fmaf(ddot, cv.x, gamma * wv.x); acc.y += fmaf(ddot, cv.y, gamma * wv.y);
        acc.z += fmaf(ddot, cv.z, gamma * wv.z); acc.w += fmaf(ddot, cv.w, gamma * wv.w);
      }
      bf16 a0, a1, a2, b0, b1, b2, c0, c1, c2, d0, d1, d2;
      split3(dc.x, a0, a1, a2); split3(dc.y, b0, b1, b2); split3(dc.z, c0, c1, c2); split3(dc.w, d0, d1, d2);
      const size_t o = row * D + 4 * (size_t)dq;
      __nv_bfloat162 q0[2] = {__nv_bfloat162(a0, b0), __nv_bfloat162(c0, d0)};
      __nv_bfloat162 q1[2] = {__nv_bfloat162(a1, b1), __nv_bfloat162(c1, d1)};
      __nv_bfloat162 q2[2] = {__nv_bfloat162(a2, b2), __nv_bfloat162(c2, d2)};
      *reinterpret_cast<uint2*>(dcp + o) = *reinterpret_cast<uint2*>(q0);
      *reinterpret_cast<uint2*>(dcp + plane + o) = *reinterpret_cast<uint2*>(q1);
      *reinterpret_cast<uint2*>(dcp + 2 * plane + o) = *reinterpret_cast<uint2*>(q2);
    }
  }
  if (!live) return;                                       // (uniform over the CTA)
  part[threadIdx.x] = acc;
  __syncthreads();
  if (jg == 0) {
    for (int q = 1; q < ngroups; ++q) {
      const float4 o = part[q * nq + dq];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    reinterpret_cast<float4*>(dwt32 + ((size_t)i * Lw + off + l) * D)[dq] = acc;
  }
}

// Backward of the two softmaxes for one (image, caption) block.  DAt block [Lp, Sq]: dL/dA (transposed); AT_p: A;
// SC block: P.  Output DS_p block: dL/dscores as bf16 pieces (zero in padded rows / columns).  512 threads;
// dynamic smem: S * (Lp + 1) + 5 * 128 floats
__global__ void __launch_bounds__(SM_THREADS) softmax_bwd(const float* __restrict__ dat, const bf16* __restrict__ atp,
                                                          const float* __restrict__ sc, bf16* __restrict__ dsp,
                                                          const int* __restrict__ cap_lens, int i0, int nc, int Bi, int Bc, int S,
                                                          int Sq, int Lcap, int Lp, int NC, float temp1,
                                                          const float* __restrict__ d_attn_diag,
                                                          const float* __restrict__ d_attn_mean) {
  extern __shared__ float sm[];
  const int LP1 = Lp + 1;
  float* tile = sm;                          // g[s][l]
  float* zpart = sm + (size_t)S * LP1;       // [SM_PARTS][128]
  float* rs = zpart + SM_PARTS * 128;        // [128]
  const int p = blockIdx.x, j = p / nc, ii = p - j * nc, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SM_THREADS >> 5;
  const float* db = dat + ((size_t)j * NC + (size_t)ii * Lp) * Sq;
  const float* ed = (d_attn_diag != nullptr && j == i) ? d_attn_diag + (size_t)i * Lcap * S : nullptr;
  const float* em = (d_attn_mean != nullptr) ? d_attn_mean + ((size_t)j * Bc + i) * S : nullptr;
  const float invL = L > 0 ? 1.f / (float)L : 0.f;
  // dL/dA of the block, transposed into g[s][l]: a warp owns a word row of DAt (coalesced along s)
  for (int l = warp; l < L; l += nwarps) {
    const float* src = db + (size_t)l * Sq;
#pragma unroll 4
    for (int x = lane; x < S; x += 32) {
      float v = src[x];
      if (ed) v += ed[(size_t)l * S + x];
      if (em) v += em[x] * invL;
      tile[(size_t)x * LP1 + l] = v;
    }
  }
  __syncthreads();
  // softmax #2 backward: dZ = A (dA - sum_s A dA);  dP = temp1 dZ      (SM_PARTS threads per word, rows interleaved)
  const size_t plane = (size_t)Bi * Sq * NC;
  const bf16* ab = atp + (size_t)j * Sq * NC + (size_t)ii * Lp;
  const int lc = tid & 127, part = tid >> 7;
  {
    float acc = 0.f;
    if (lc < L) {
#pragma unroll 4
      for (int s = part; s < S; s += SM_PARTS) {
        const size_t o = (size_t)s * NC + lc;
        acc = fmaf(join3(ab[o], ab[plane + o], ab[2 * plane + o]), tile[(size_t)s * LP1 + lc], acc);
      }
    }
    zpart[part * 128 + lc] = acc;
  }
  __syncthreads();
  if (tid < 128) {
    float z = 0.f;
#pragma unroll
    for (int q = 0; q < SM_PARTS; ++q) z += zpart[q * 128 + tid];
    rs[tid] = z;
  }
  __syncthreads();
  if (lc < L) {
    const float r = rs[lc];
#pragma unroll 4
    for (int s = part; s < S; s += SM_PARTS) {
      const size_t o = (size_t)s * NC + lc;
      const float a = join3(ab[o], ab[plane + o], ab[2 * plane + o]);
      float* t = tile + (size_t)s * LP1 + lc;
      *t = temp1 * a * (*t - r);
    }
  }
  __syncthreads();
  // softmax #1 backward: dS = P (dP - sum_l P dP): a warp owns RB region rows at a time; bf16 pieces out
  const float* pb = sc + (size_t)j * Sq * NC + (size_t)ii * Lp;
  bf16* ob = dsp + (size_t)j * Sq * NC + (size_t)ii * Lp;
  for (int s0 = warp * RB; s0 < Sq; s0 += nwarps * RB) {
    float P[RB][4], g[RB][4], t[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int s = s0 + r;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int l = lane + 32 * k;
        const bool in = s < S && l < L;
        P[r][k] = in ? pb[(size_t)s * NC + l] : 0.f;
        g[r][k] = in ? tile[(size_t)s * LP1 + l] : 0.f;
        d = fmaf(P[r][k], g[r][k], d);
      }
      t[r] = d;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) t[r] = warp_sum(t[r]);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int s = s0 + r;
      if (s < Sq) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int l = lane + 32 * k;
          if (l < Lp) {
            bf16 p0, p1, p2;
            split3(P[r][k] * (g[r][k] - t[r]), p0, p1, p2);
            const size_t o = (size_t)s * NC + l;
            ob[o] = p0; ob[plane + o] = p1; ob[2 * plane + o] = p2;
          }
        }
      }
    }
  }
}

// dWt32[i0+ii][off+l][:] += dWc[ii*Lp + l][:]  for the live words of the chunk
__global__ void add_word_grad(const float* __restrict__ dwc, float* __restrict__ dwt32, const int* __restrict__ cap_lens, int i0,
                              int Lcap, int Lp, int Lw, int off, int D) {
  const int l = blockIdx.x, ii = blockIdx.y, i = i0 + ii;
  const int L = min(max(cap_lens[i], 0), Lcap);
  if (l >= L) return;
  const float* src = dwc + ((size_t)ii * Lp + l) * D;
  float* dst = dwt32 + ((size_t)i * Lw + off + l) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] += src[d];
}

// dRt [Bi*Sq, D] -> d_ctx [Bi, D, S]
__global__ void unpack_dctx(const float* __restrict__ drt, float* __restrict__ d_ctx, int D, int S, int Sq) {
  __shared__ float t[32][33];
  const int j = blockIdx.z, s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = s0 + r, d = d0 + threadIdx.x;
    t[r][threadIdx.x] = drt[((size_t)j * Sq + s) * D + d];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, s = s0 + threadIdx.x;
    if (s < S) d_ctx[((size_t)j * D + d) * S + s] = t[threadIdx.x][r];
  }
}
// dWt32 [Bc, Lw, D] -> d_words [Bc, D, Lw], zero outside [off, off + cap_len)
__global__ void unpack_dwords(const float* __restrict__ dwt, float* __restrict__ dwords, const int* __restrict__ cap_lens, int D,
                              int Lw, int Lcap, int off) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int L = min(max(cap_lens[b], 0), Lcap);
  const float* ib = dwt + (size_t)b * D * Lw;
  float* ob = dwords + (size_t)b * D * Lw;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    t[r][threadIdx.x] = (l < Lw && d < D && l >= off && l < off + L) ? ib[(size_t)l * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, l = l0 + threadIdx.x;
    if (d < D && l < Lw) ob[(size_t)d * Lw + l] = t[threadIdx.x][r];
  }
}

// -------------------------------------------------------------------------------------------------------------
// host orchestration
// -------------------------------------------------------------------------------------------------------------
struct Dims {
  int Bi, Bc, D, S, Sq, Lw, Lcap, Lp, off;
};
struct Plan {
  int nc, NCmax;
  size_t rt, rn, wt32, wn, dwt32, drt, wnp, wtp, sc, atp, cx, coef, dcp, dat, dsp, dwc, total;
};

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

static size_t softmax_smem(int S, int Lp) { return ((size_t)S * (Lp + 1) + (SM_PARTS + 1) * 128) * sizeof(float); }

static size_t fixed_bytes(const Dims& d, bool bwd) {
  size_t f = align_up((size_t)3 * d.Bi * d.Sq * d.D * 2, 256) + align_up((size_t)d.Bc * d.Lw * d.D * 4, 256) +
             align_up((size_t)d.Bc * d.Lw * 4, 256);
  if (bwd) f += align_up((size_t)3 * d.Bi * d.D * d.Sq * 2, 256) + align_up((size_t)d.Bc * d.Lw * d.D * 4, 256) +
                align_up((size_t)d.Bi * d.Sq * d.D * 4, 256);
  return f + 8192;
}
// bytes per column (ii, l) of the chunk matrices
static size_t column_bytes(const Dims& d, bool bwd) {
  size_t c = (size_t)6 * d.D + (size_t)d.Bi * d.Sq * 4 + (size_t)d.Bi * d.Sq * 6 + (size_t)d.Bi * d.D * 4;
  if (bwd) c += (size_t)6 * d.D + (size_t)d.Bi * 12 + (size_t)d.Bi * d.D * 6 + (size_t)d.Bi * d.Sq * 4 + (size_t)d.Bi * d.Sq * 6 +
                (size_t)d.D * 4;
  return c;
}
static Plan make_plan(const Dims& d, size_t bytes, bool bwd) {
  Plan pl{};
  const size_t fixed = fixed_bytes(d, bwd), col = column_bytes(d, bwd);
  const size_t min_cols = (size_t)round_up(d.Lp, 64);
  if (bytes < fixed + col * min_cols + 16 * 256) { pl.nc = 0; return pl; }
  size_t cols = (bytes - fixed - 16 * 256) / col;
  size_t nc = cols >= 64 ? (cols - 63) / d.Lp : 0;          // NC = round_up(nc * Lp, 64) <= nc * Lp + 63
  if (nc < 1) nc = 1;
  if (nc > (size_t)d.Bc) nc = d.Bc;
  pl.nc = (int)nc;
  const size_t NC = (size_t)round_up((int)nc * d.Lp, 64);
  pl.NCmax = (int)NC;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 256); return r; };
  pl.rt = take((size_t)3 * d.Bi * d.Sq * d.D * 2);
  pl.wt32 = take((size_t)d.Bc * d.Lw * d.D * 4);
  pl.wn = take((size_t)d.Bc * d.Lw * 4);
  pl.wnp = take((size_t)3 * d.D * NC * 2);
  pl.sc = take((size_t)d.Bi * d.Sq * NC * 4);
  pl.atp = take((size_t)3 * d.Bi * d.Sq * NC * 2);
  pl.cx = take((size_t)d.Bi * NC * d.D * 4);
  if (bwd) {
    pl.rn = take((size_t)3 * d.Bi * d.D * d.Sq * 2);
    pl.dwt32 = take((size_t)d.Bc * d.Lw * d.D * 4);
    pl.drt = take((size_t)d.Bi * d.Sq * d.D * 4);
    pl.wtp = take((size_t)3 * NC * d.D * 2);
    pl.coef = take((size_t)d.Bi * nc * 3 * d.Lp * 4);
    pl.dcp = take((size_t)3 * d.Bi * NC * d.D * 2);
    pl.dat = take((size_t)d.Bi * NC * d.Sq * 4);
    pl.dsp = take((size_t)3 * d.Bi * d.Sq * NC * 2);
    pl.dwc = take(NC * d.D * 4);
  }
  pl.total = o;
  if (pl.total > bytes) pl.nc = 0;
  return pl;
}

static int supported(int D, int S, int Lcap) {
  if (D <= 0 || S <= 0 || Lcap <= 0) return 1;
  if (D % 64 || D > 2048) return 1;
  const int Lp = round_up(Lcap, 8);
  if (Lp > 128) return 1;
  if (softmax_smem(S, Lp) > 220 * 1024) return 1;
  return 0;
}

static int check_common(const void* ctx, const void* words, const void* cap_lens, int Bi, int Bc, int D, int S, int Lw, int Lcap,
                        int word_off, int agg) {
  GLORIA_CHECK_ARG(ctx && words && cap_lens, "null input pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && D > 0 && S > 0 && Lw > 0, "non-positive size (Bi=%d Bc=%d D=%d S=%d Lw=%d)", Bi, Bc, D, S, Lw);
  GLORIA_CHECK_ARG(word_off >= 0 && Lcap > 0 && word_off + Lcap <= Lw, "caption window [%d, %d) exceeds the word axis (%d)",
                   word_off, word_off + Lcap, Lw);
  GLORIA_CHECK_ARG(agg == GLORIA_AGG_SUM || agg == GLORIA_AGG_MEAN || agg == GLORIA_AGG_MAX, "bad agg %d", agg);
  if (supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "fp32 tensor-core path needs D %% 64 == 0 and cap_len <= 128 (D=%d S=%d Lcap=%d)", D, S, Lcap);
  if ((long long)Bi * round_up(S, 64) >= (1 << 24)) return fail(GLORIA_ERR_UNSUPPORTED, "too many region rows");
  return GLORIA_OK;
}

// pieces of the region features and the fp32 word rows + norms: once per call
static int prepare(const float* ctx, const float* words, const Dims& d, const Plan& pl, char* ws, bool bwd, cudaStream_t st) {
  bf16* rt = (bf16*)(ws + pl.rt);
  bf16* rn = bwd ? (bf16*)(ws + pl.rn) : nullptr;
  split_ctx<<<dim3(d.Sq / 32, d.D / 32, d.Bi), dim3(32, 8), 0, st>>>(ctx, rt, rn, d.Bi, d.D, d.S, d.Sq);
  GLORIA_LAUNCHED("f32tc::split_ctx");
  float* wt32 = (float*)(ws + pl.wt32);
  transpose_words<<<dim3((d.Lw + 31) / 32, (d.D + 31) / 32, d.Bc), dim3(32, 8), 0, st>>>(words, wt32, d.D, d.Lw);
  GLORIA_LAUNCHED("f32tc::transpose_words");
  const long long rows = (long long)d.Bc * d.Lw;
  word_norms<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(wt32, (float*)(ws + pl.wn), rows, d.D);
  GLORIA_LAUNCHED("f32tc::word_norms");
  return GLORIA_OK;
}

// scores, P, A, C of the chunk [i0, i0 + nc)
static int chunk_forward(const Dims& d, const Plan& pl, char* ws, const int32_t* cap_lens, int i0, int nc, int NC, float temp1,
                         int score_terms, int ctx_terms, bool bwd, float* attn_diag, float* attn_mean, cudaStream_t st) {
  int rc;
  bf16* rt = (bf16*)(ws + pl.rt);
  bf16* wnp = (bf16*)(ws + pl.wnp);
  bf16* wtp = bwd ? (bf16*)(ws + pl.wtp) : nullptr;
  float* sc = (float*)(ws + pl.sc);
  bf16* atp = (bf16*)(ws + pl.atp);
  float* cx = (float*)(ws + pl.cx);
  split_words<<<dim3(NC / 32, d.D / 32), dim3(32, 8), 0, st>>>((const float*)(ws + pl.wt32), cap_lens, wnp, wtp, i0, nc, d.Lw,
                                                              d.Lcap, d.Lp, d.off, d.D, NC);
  GLORIA_LAUNCHED("f32tc::split_words");
  {  // GEMM 1: SC[(j,s), (ii,l)] = sum_d R[(j,s), d] W[d, (ii,l)]            (bmm #1, gloria_loss.py:40)
    tc::GemmEx e{};
    e.A = rt; e.B = wnp; e.C = sc;
    e.M = d.Bi * d.Sq; e.N = NC; e.K = d.D; e.ldc = NC; e.a_kmajor = true; e.ksplit = 1;
    e.nb = 1; e.nterms = score_terms;
    e.a_prows = (long long)d.Bi * d.Sq; e.b_prows = d.D;
    e.a_rows = 3LL * d.Bi * d.Sq; e.b_rows = 3LL * d.D;
    if ((rc = tc::acc_gemm_ex(e, st))) return rc;
  }
  softmax_fwd<<<(unsigned)(d.Bi * nc), SM_THREADS, softmax_smem(d.S, d.Lp), st>>>(sc, atp, cap_lens, i0, nc, d.Bi, d.Bc, d.S, d.Sq, d.Lcap,
                                                                          d.Lp, NC, temp1, attn_diag, attn_mean);
  GLORIA_LAUNCHED("f32tc::softmax_fwd");
  if (NC > nc * d.Lp) {      // padding columns of A: exact zeros in every plane
    GLORIA_CUDA(cudaMemset2DAsync(atp + (size_t)nc * d.Lp, (size_t)NC * 2, 0, (size_t)(NC - nc * d.Lp) * 2, (size_t)3 * d.Bi * d.Sq, st));
  }
  {  // GEMM 2: CX_j[(ii,l), d] = sum_s A[(j,s), (ii,l)] R[(j,s), d]           (bmm #2, gloria_loss.py:59)
    tc::GemmEx e{};
    e.A = atp; e.B = rt; e.C = cx;
    e.M = NC; e.N = d.D; e.K = d.Sq; e.ldc = d.D; e.a_kmajor = false; e.ksplit = 1;
    e.nb = d.Bi; e.a_brows = d.Sq; e.b_brows = d.Sq; e.c_bstride = (long long)NC * d.D;
    e.nterms = ctx_terms;
    e.a_prows = (long long)d.Bi * d.Sq; e.b_prows = (long long)d.Bi * d.Sq;
    e.a_rows = 3LL * d.Bi * d.Sq; e.b_rows = 3LL * d.Bi * d.Sq;
    if ((rc = tc::acc_gemm_ex(e, st))) return rc;
  }
  return GLORIA_OK;
}

}  // namespace f32tc

// Scores of the B diagonal pairs, sc[i][l][s] = sum_d Wt32[i][off+l][d] ctx[i][d][s] ([B, Lcap, S], fp32 accuracy), on the
// split-precision tensor-core GEMM: one batch per pair (att_maps of local_loss, gloria_loss.py:141-143; get_attn_maps,
// gloria_model.py:209-211).  Returns 0 bytes when the shape is not covered (the caller then uses its CUDA-core GEMM).
size_t f32tc_diag_scores_workspace(int B, int D, int S, int Lcap) {
  using namespace f32tc;
  if (supported(D, S, Lcap)) return 0;
  const int Lp = round_up(Lcap, 8), Sq = round_up(S, 64), NC = round_up(B * Lp, 64);
  return align_up((size_t)3 * NC * D * 2, 256) + align_up((size_t)3 * B * D * Sq * 2, 256) + align_up((size_t)B * Lp * Sq * 4, 256);
}
int f32tc_diag_scores(const float* ctx, const float* wt32, const int32_t* cap_lens, int B, int D, int S, int Lw, int Lcap, int off,
                      float* sc, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace f32tc;
  const size_t need = f32tc_diag_scores_workspace(B, D, S, Lcap);
  if (need == 0 || ws == nullptr || ws_bytes < need) return fail(GLORIA_ERR_WORKSPACE, "diag scores: workspace %zu B < %zu B", ws_bytes, need);
  const int Lp = round_up(Lcap, 8), Sq = round_up(S, 64), NC = round_up(B * Lp, 64);
  char* w = (char*)ws;
  bf16* wtp = (bf16*)w;
  bf16* rn = (bf16*)(w + align_up((size_t)3 * NC * D * 2, 256));
  float* tmp = (float*)((char*)rn + align_up((size_t)3 * B * D * Sq * 2, 256));
  split_ctx<<<dim3(Sq / 32, D / 32, B), dim3(32, 8), 0, st>>>(ctx, nullptr, rn, B, D, S, Sq);
  GLORIA_LAUNCHED("f32tc::split_ctx(diag)");
  split_words<<<dim3(NC / 32, D / 32), dim3(32, 8), 0, st>>>(wt32, cap_lens, nullptr, wtp, 0, B, Lw, Lcap, Lp, off, D, NC);
  GLORIA_LAUNCHED("f32tc::split_words(diag)");
  tc::GemmEx e{};
  e.A = wtp; e.B = rn; e.C = tmp;
  e.M = Lp; e.N = Sq; e.K = D; e.ldc = Sq; e.a_kmajor = true; e.ksplit = 1;
  e.nb = B; e.a_brows = Lp; e.b_brows = D; e.c_bstride = (long long)Lp * Sq;
  e.nterms = 6;
  e.a_prows = NC; e.b_prows = (long long)B * D;
  e.a_rows = 3LL * NC; e.b_rows = 3LL * B * D;
  int rc = tc::acc_gemm_ex(e, st);
  if (rc) return rc;
  repack_diag_scores<<<dim3((unsigned)Lcap, (unsigned)B), 128, 0, st>>>(tmp, sc, Lcap, Lp, S, Sq);
  GLORIA_LAUNCHED("f32tc::repack_diag_scores");
  return GLORIA_OK;
}

}  // namespace gloria

using namespace gloria;
using namespace gloria::f32tc;

extern "C" int gloria_b200_f32tc_supported(int D, int S, int Lcap) { return supported(D, S, Lcap); }

extern "C" size_t gloria_b200_local_f32tc_workspace(int Bi, int Bc, int D, int S, int Lw, int Lcap, size_t budget, int backward) {
  if (Bi <= 0 || Bc <= 0 || D <= 0 || S <= 0 || Lw <= 0 || Lcap <= 0) return 0;
  Dims d{Bi, Bc, D, S, round_up(S, 64), Lw, Lcap, round_up(Lcap, 8), 0};
  const bool bwd = backward != 0;
  const size_t fixed = fixed_bytes(d, bwd), col = column_bytes(d, bwd);
  size_t want = fixed + col * ((size_t)Bc * d.Lp + 64) + 16 * 256;
  const size_t least = fixed + col * ((size_t)d.Lp + 64) + 16 * 256;
  if (budget != 0 && want > budget) want = budget > least ? budget : least;
  return want;
}

static int set_smem(const void* fn, size_t bytes) {
  GLORIA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return GLORIA_OK;
}

// keep = true: the workspace is laid out for the backward, must hold all captions in one chunk, and is left holding the
// forward's state (operand pieces, P, A, C) for gloria_b200_local_sim_bwd_f32tc(..., state_from_forward = 1)
static int forward_impl(const float* ctx, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D, int S, int Lw,
                        int Lcap, int word_off, float temp1, float temp2, int agg, float eps, float* sim, float* attn_diag,
                        float* attn_mean, void* workspace, size_t workspace_bytes, bool keep, cudaStream_t st) {
  int rc = check_common(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, agg);
  if (rc) return rc;
  GLORIA_CHECK_ARG(sim != nullptr && workspace != nullptr, "null output / workspace");
  GLORIA_CHECK_ARG(attn_diag == nullptr || Bi == Bc, "attn_diag needs Bi == Bc (got %d x %d)", Bi, Bc);
  Dims d{Bi, Bc, D, S, round_up(S, 64), Lw, Lcap, round_up(Lcap, 8), word_off};
  const Plan pl = make_plan(d, workspace_bytes, keep);
  if (pl.nc < 1 || (keep && pl.nc < Bc))
    return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B too small%s", workspace_bytes, keep ? " to keep the forward's state" : "");
  if ((rc = set_smem((const void*)softmax_fwd, softmax_smem(S, d.Lp)))) return rc;
  char* ws = (char*)workspace;
  if ((rc = prepare(ctx, words, d, pl, ws, keep, st))) return rc;
  for (int i0 = 0; i0 < Bc; i0 += pl.nc) {
    const int nc = min(pl.nc, Bc - i0);
    const int NC = round_up(nc * d.Lp, 64);
    // scores with all six piece products (they pass through two softmaxes); the context GEMM with three: its 2^-17
    // relative error per product averages out over the D channels of the cosine
    if ((rc = chunk_forward(d, pl, ws, cap_lens, i0, nc, NC, temp1, 6, 3, keep, attn_diag, attn_mean, st))) return rc;
    cosine_agg<<<(unsigned)(Bi * nc), 256, 3 * d.Lp * sizeof(float), st>>>(
        (const float*)(ws + pl.cx), (const float*)(ws + pl.wt32), (const float*)(ws + pl.wn), cap_lens, i0, nc, Bc, Lcap, d.Lp, NC, Lw,
        word_off, D, temp2, agg, eps, sim, nullptr, nullptr);
    GLORIA_LAUNCHED("f32tc::cosine_agg");
  }
  return GLORIA_OK;
}

extern "C" int gloria_b200_local_sim_fwd_f32tc(const float* ctx, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D,
                                               int S, int Lw, int Lcap, int word_off, float temp1, float temp2, int agg, float eps,
                                               float* sim, float* attn_diag, float* attn_mean, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  return forward_impl(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, temp1, temp2, agg, eps, sim, attn_diag, attn_mean,
                      workspace, workspace_bytes, false, (cudaStream_t)stream);
}

extern "C" int gloria_b200_local_sim_fwd_f32tc_train(const float* ctx, const float* words, const int32_t* cap_lens, int Bi, int Bc,
                                                     int D, int S, int Lw, int Lcap, int word_off, float temp1, float temp2, int agg,
                                                     float eps, float* sim, float* attn_diag, float* attn_mean, void* workspace,
                                                     size_t workspace_bytes, void* stream) {
  return forward_impl(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, temp1, temp2, agg, eps, sim, attn_diag, attn_mean,
                      workspace, workspace_bytes, true, (cudaStream_t)stream);
}

extern "C" int gloria_b200_local_sim_bwd_f32tc(const float* ctx, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D,
                                               int S, int Lw, int Lcap, int word_off, float temp1, float temp2, int agg, float eps,
                                               const float* dsim, const float* d_attn_diag, const float* d_attn_mean, float* d_ctx,
                                               float* d_words, void* workspace, size_t workspace_bytes, int state_from_forward,
                                               void* stream) {
  int rc = check_common(ctx, words, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, agg);
  if (rc) return rc;
  GLORIA_CHECK_ARG(dsim && d_ctx && d_words && workspace, "null gradient / workspace pointer");
  GLORIA_CHECK_ARG(d_attn_diag == nullptr || Bi == Bc, "d_attn_diag needs Bi == Bc (got %d x %d)", Bi, Bc);
  if (agg == GLORIA_AGG_MAX) return fail(GLORIA_ERR_UNSUPPORTED, "backward of agg=max is not part of the path");
  cudaStream_t st = (cudaStream_t)stream;
  Dims d{Bi, Bc, D, S, round_up(S, 64), Lw, Lcap, round_up(Lcap, 8), word_off};
  const Plan pl = make_plan(d, workspace_bytes, true);
  if (pl.nc < 1) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B too small", workspace_bytes);
  const bool have_state = state_from_forward != 0;
  if (have_state && pl.nc < Bc) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B cannot be a kept forward state", workspace_bytes);
  if ((rc = set_smem((const void*)softmax_fwd, softmax_smem(S, d.Lp)))) return rc;
  if ((rc = set_smem((const void*)softmax_bwd, softmax_smem(S, d.Lp)))) return rc;
  char* ws = (char*)workspace;
  bf16* rt = (bf16*)(ws + pl.rt);
  bf16* rn = (bf16*)(ws + pl.rn);
  float* wt32 = (float*)(ws + pl.wt32);
  float* wn = (float*)(ws + pl.wn);
  float* dwt32 = (float*)(ws + pl.dwt32);
  float* drt = (float*)(ws + pl.drt);
  bf16* wtp = (bf16*)(ws + pl.wtp);
  float* sc = (float*)(ws + pl.sc);
  bf16* atp = (bf16*)(ws + pl.atp);
  float* cx = (float*)(ws + pl.cx);
  float* coef = (float*)(ws + pl.coef);
  bf16* dcp = (bf16*)(ws + pl.dcp);
  float* dat = (float*)(ws + pl.dat);
  bf16* dsp = (bf16*)(ws + pl.dsp);
  float* dwc = (float*)(ws + pl.dwc);
  if (!have_state && (rc = prepare(ctx, words, d, pl, ws, true, st))) return rc;
  GLORIA_CUDA(cudaMemsetAsync(drt, 0, (size_t)Bi * d.Sq * D * sizeof(float), st));
  GLORIA_CUDA(cudaMemsetAsync(dwt32, 0, (size_t)Bc * Lw * D * sizeof(float), st));
  for (int i0 = 0; i0 < Bc; i0 += pl.nc) {
    const int nc = min(pl.nc, Bc - i0);
    const int NC = round_up(nc * d.Lp, 64);
    const int tail = NC - nc * d.Lp;
    // recompute the forward intermediates of the chunk unless the forward left them here
    if (!have_state && (rc = chunk_forward(d, pl, ws, cap_lens, i0, nc, NC, temp1, 6, 3, true, nullptr, nullptr, st))) return rc;
    cosine_agg<<<(unsigned)(Bi * nc), 256, 3 * d.Lp * sizeof(float), st>>>(cx, wt32, wn, cap_lens, i0, nc, Bc, Lcap, d.Lp, NC, Lw,
                                                                          word_off, D, temp2, agg, eps, nullptr, dsim, coef);
    GLORIA_LAUNCHED("f32tc::cosine_agg(bwd)");
    context_grad<<<dim3((unsigned)d.Lp, (unsigned)nc), CG_THREADS, CG_THREADS * sizeof(float4), st>>>(cx, wt32, dwt32, coef, cap_lens, dcp, i0, nc, Bi, Lcap, d.Lp, NC,
                                                                     Lw, word_off, D);
    GLORIA_LAUNCHED("f32tc::context_grad");
    if (tail > 0)            // padding rows of dC: exact zeros in every plane
      GLORIA_CUDA(cudaMemset2DAsync(dcp + (size_t)nc * d.Lp * D, (size_t)NC * D * 2, 0, (size_t)tail * D * 2, (size_t)3 * Bi, st));
    {  // GEMM 3: DAt_j[(ii,l), s] = sum_d dC_j[(ii,l), d] R_j[d, s]
      tc::GemmEx e{};
      e.A = dcp; e.B = rn; e.C = dat;
      e.M = NC; e.N = d.Sq; e.K = D; e.ldc = d.Sq; e.a_kmajor = true; e.ksplit = 1;
      e.nb = Bi; e.a_brows = NC; e.b_brows = D; e.c_bstride = (long long)NC * d.Sq;
      e.nterms = 3;
      e.a_prows = (long long)Bi * NC; e.b_prows = (long long)Bi * D;
      e.a_rows = 3LL * Bi * NC; e.b_rows = 3LL * Bi * D;
      if ((rc = tc::acc_gemm_ex(e, st))) return rc;
    }
    {  // GEMM 4: dRt_j[s, d] += sum_(ii,l) A[(j,s), (ii,l)] dC_j[(ii,l), d]
      tc::GemmEx e{};
      e.A = atp; e.B = dcp; e.C = drt;
      e.M = d.Sq; e.N = D; e.K = NC; e.ldc = D; e.a_kmajor = true; e.ksplit = 1; e.accumulate = true;
      e.nb = Bi; e.a_brows = d.Sq; e.b_brows = NC; e.c_bstride = (long long)d.Sq * D;
      e.nterms = 3;
      e.a_prows = (long long)Bi * d.Sq; e.b_prows = (long long)Bi * NC;
      e.a_rows = 3LL * Bi * d.Sq; e.b_rows = 3LL * Bi * NC;
      if ((rc = tc::acc_gemm_ex(e, st))) return rc;
    }
    softmax_bwd<<<(unsigned)(Bi * nc), SM_THREADS, softmax_smem(S, d.Lp), st>>>(dat, atp, sc, dsp, cap_lens, i0, nc, Bi, Bc, S, d.Sq, Lcap, d.Lp,
                                                                        NC, temp1, d_attn_diag, d_attn_mean);
    GLORIA_LAUNCHED("f32tc::softmax_bwd");
    if (tail > 0)
      GLORIA_CUDA(cudaMemset2DAsync(dsp + (size_t)nc * d.Lp, (size_t)NC * 2, 0, (size_t)tail * 2, (size_t)3 * Bi * d.Sq, st));
    {  // GEMM 5: dWc[(ii,l), d] = sum_(j,s) DS[(j,s), (ii,l)] R[(j,s), d]
      tc::GemmEx e{};
      e.A = dsp; e.B = rt; e.C = dwc;
      e.M = NC; e.N = D; e.K = Bi * d.Sq; e.ldc = D; e.a_kmajor = false; e.ksplit = 0;
      e.nb = 1; e.nterms = 3;
      e.a_prows = (long long)Bi * d.Sq; e.b_prows = (long long)Bi * d.Sq;
      e.a_rows = 3LL * Bi * d.Sq; e.b_rows = 3LL * Bi * d.Sq;
      if ((rc = tc::acc_gemm_ex(e, st))) return rc;
    }
    add_word_grad<<<dim3((unsigned)d.Lp, (unsigned)nc), 256, 0, st>>>(dwc, dwt32, cap_lens, i0, Lcap, d.Lp, Lw, word_off, D);
    GLORIA_LAUNCHED("f32tc::add_word_grad");
    {  // GEMM 6: dRt[(j,s), d] += sum_(ii,l) DS[(j,s), (ii,l)] W[(ii,l), d]
      tc::GemmEx e{};
      e.A = dsp; e.B = wtp; e.C = drt;
      e.M = Bi * d.Sq; e.N = D; e.K = NC; e.ldc = D; e.a_kmajor = true; e.ksplit = 1; e.accumulate = true;
      e.nb = 1; e.nterms = 3;
      e.a_prows = (long long)Bi * d.Sq; e.b_prows = NC;
      e.a_rows = 3LL * Bi * d.Sq; e.b_rows = 3LL * NC;
      if ((rc = tc::acc_gemm_ex(e, st))) return rc;
    }
  }
  unpack_dctx<<<dim3(d.Sq / 32, D / 32, Bi), dim3(32, 8), 0, st>>>(drt, d_ctx, D, S, d.Sq);
  GLORIA_LAUNCHED("f32tc::unpack_dctx");
  unpack_dwords<<<dim3((Lw + 31) / 32, (D + 31) / 32, Bc), dim3(32, 8), 0, st>>>(dwt32, d_words, cap_lens, D, Lw, Lcap, word_off);
  GLORIA_LAUNCHED("f32tc::unpack_dwords");
  return GLORIA_OK;
}

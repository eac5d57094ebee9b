// bf16 tensor-core mode of the GLoRIA local similarity (sm_100a: TMA + tcgen05 + TMEM).
//
// One "unit" = (caption i, image j).  Per unit, inside one persistent CTA:
//   GEMM1  D1[s, l] = sum_d Rt[j][s][d] * Wt[i][l][d]            M = 128 regions (x NT tiles), N = LPAD, K = D
//          -> TMEM (lanes = regions, double buffered), so softmax #1 over the caption's words is thread-local
//   softmax warps:  P = softmax_l(D1);  E = exp(temp1 * P)  (un-normalised softmax #2 numerator, in [1, e^temp1])
//          -> bf16, written to shared memory as the MN-major A operand of GEMM2
//   GEMM2  D2[l, d] = sum_s E[s, l] * Rn[j][d][s]                M = 128 words, N = 128 channels (x D/128), K = Spad
//          -> TMEM (lanes = words, double buffered): C' = Z_l * weightedContext
//   epilogue warps: dot_l = <W_l, C'_l>, |C'_l|^2 thread-local over d; cos_l is invariant to the softmax-#2
//          normaliser Z_l, so it is never needed; then the temp2 log-sum-exp over words.
// Only sim[Bi, Bc] reaches HBM.  gloria_loss.py:19-63 + :144-158 per pair.
#include "common.cuh"
#include "tc_common.cuh"

namespace gloria {
namespace tc {

constexpr int SLOT = 16384;              // one operand tile of one k-block (128 rows x 128 B)
constexpr int NSLOT = 6;                 // TMA ring: GEMM1 takes two slots per k-block (Rt, Wt), GEMM2 one (Rn)
constexpr int E_BYTES = 2 * MAX_NT * TILE * 128;   // [2 word blocks of 64][Spad regions][128 B]
constexpr int OFF_E = NSLOT * SLOT;
constexpr int OFF_BAR = OFF_E + E_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + barriers/scratch + alignment slack
constexpr int NTHREADS = 384;            // warps 0-3 control, 4-7 softmax, 8-11 epilogue

// barrier indices (8 B each) inside the barrier block
enum { B_FULL = 0, B_EMPTY = NSLOT, B_D1F = 2 * NSLOT, B_D1E = B_D1F + 2, B_EF = B_D1E + 2, B_EE, B_D2F, B_D2E = B_D2F + 2,
       B_COUNT = B_D2E + 2 };

struct FwdParams {
  const __half* wt;          // [Bc, LPAD, D] fp16 (the score GEMM's operands are fp16, see pack_ctx)
  const float* wnorm;        // [Bc, LPAD]
  const int* cap_lens;
  float* sim;                // [Bi, Bc]
  float* stats;              // [Bi, Bc, 2, LPAD] (dot', |C'|^2 of the un-normalised context) or nullptr
  int Bi, Bc, D, S, NT;
  int n_caps, per;           // SEG > 0 ("packed" mode): Bc counts word tiles; tile g holds captions g*per .. g*per+per-1 (of
                             // n_caps) in 16-word segments; sim is [Bi, n_caps]
  float t1_log2e;            // temp1 * log2(e)
  float temp2;
  int agg;
  float eps_s;               // eps * S  (clamp on the un-normalised context, see header comment)
  long long* dbg;            // phase clocks (only read when built with -DGLORIA_PHASE_CLOCKS)
};

// SEG = 0: one caption per word tile.  SEG = 16 ("packed" mode, inference only): LPAD / 16 captions of at most 16 words
// share one word tile -- zero-shot prompts are a handful of words, and the kernel's cost is set by the image tiles it
// streams per (image, word tile), not by the tile's width.  Softmax #1 and the word aggregation then run per segment;
// everything in between is per word and does not change.
template <int LPAD, int SEG>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_rt, const __grid_constant__ CUtensorMap tm_wt,
              const __grid_constant__ CUtensorMap tm_rn, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  float* red = reinterpret_cast<float*>(smem + OFF_BAR + B_COUNT * 8 + 16);   // 8 floats of LSE scratch
  auto bar = [&](int idx) { return bars + 8u * (uint32_t)idx; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Spad = p.NT * TILE;
  const int nkb1 = p.D / KBLK;        // k-blocks of GEMM1
  const int nkb2 = Spad / KBLK;       // k-blocks of GEMM2
  const int nchunk = p.D / TILE;      // D2 chunks

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(B_D1F + b), 1); mbar_init(bar(B_D1E + b), 128);
      mbar_init(bar(B_D2F + b), 1); mbar_init(bar(B_D2E + b), 128);
    }
    mbar_init(bar(B_EF), 128 * p.NT);
    mbar_init(bar(B_EE), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_rt); tma_prefetch_desc(&tm_wt); tma_prefetch_desc(&tm_rn);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int slot = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* tm, int x, int y, uint32_t bytes) {
        mbar_wait(bar(B_EMPTY + slot), ph ^ 1);
        mbar_expect_tx(bar(B_FULL + slot), bytes);
        tma_load_2d(base + slot * SLOT, tm, x, y, bar(B_FULL + slot));
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
      };
      auto load_g1 = [&](int i, int j, int t) {
        for (int kb = 0; kb < nkb1; ++kb) {
          load(&tm_rt, kb * KBLK, j * Spad + t * TILE, TILE * 128);
          load(&tm_wt, kb * KBLK, i * LPAD, LPAD * 128);
        }
      };
      auto load_g2 = [&](int j) {
        for (int c = 0; c < nchunk; ++c)
          for (int kb = 0; kb < nkb2; ++kb) load(&tm_rn, kb * KBLK, j * p.D + c * TILE, TILE * 128);
      };
      // same order as the MMA issuer (see there)
      UnitIter it(p.Bi, p.Bc);
      bool has = it.next();
      int ci = it.cap(), cj = it.j;
      if (has)
        for (int t = 0; t < p.NT; ++t) load_g1(ci, cj, t);
      while (has) {
        const bool hasn = it.next();
        const int ni = it.cap(), nj = it.j;
        if (hasn) load_g1(ni, nj, 0);
        load_g2(cj);
        if (hasn)
          for (int t = 1; t < p.NT; ++t) load_g1(ni, nj, t);
        has = hasn; ci = ni; cj = nj;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
#ifdef GLORIA_PHASE_CLOCKS
      long long wt_full = 0, wt_d1e = 0, wt_ef = 0, wt_d2e = 0, t_all = clock64();
#define TIMED_WAIT(acc, ...) do { long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; } while (0)
#else
#define TIMED_WAIT(acc, ...) do { __VA_ARGS__; } while (0)
#endif
      constexpr uint32_t idesc1 = make_idesc(TILE, LPAD, 0, 0, 0);   // fp16: A = Rh tile (K-major), B = Wh tile (K-major)
      constexpr uint32_t idesc2 = make_idesc(TILE, TILE, 1, 0);   // A = E (MN-major), B = Rn tile (K-major)
      const uint32_t e_lbo = (uint32_t)Spad * 128u;               // between the two 64-word blocks of E
      int slot = 0; uint32_t ph = 0;
      uint32_t g1 = 0, g2 = 0, nu = 0;
      // descriptors are built once; per MMA only the 16-byte-unit address field is advanced (the issuing thread's
      // instruction count, not the tensor pipe, bounded this loop before)
      const uint64_t d_slot0 = make_smem_desc(base, 16, 1024);            // K-major tile in slot 0
      const uint64_t d_e0 = make_smem_desc(base + OFF_E, e_lbo, 1024);    // E, MN-major
      auto take = [&]() {
        TIMED_WAIT(wt_full, mbar_wait(bar(B_FULL + slot), ph));
        const int s = slot;
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
        return s;
      };
      auto gemm1 = [&]() {                       // one score tile into the next D1 buffer
        const uint32_t b = g1 & 1;
        TIMED_WAIT(wt_d1e, mbar_wait(bar(B_D1E + b), ((g1 >> 1) & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < nkb1; ++kb) {
          const int sa = take();
          const int sb = take();
          tc_fence_after();
          const uint64_t ad = d_slot0 + (uint64_t)(sa * (SLOT >> 4)), bd = d_slot0 + (uint64_t)(sb * (SLOT >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + b * TILE, ad + 2 * k, bd + 2 * k, idesc1, (uint32_t)((kb | k) != 0));
          umma_commit(bar(B_EMPTY + sa));
          umma_commit(bar(B_EMPTY + sb));
        }
        umma_commit(bar(B_D1F + b));
        ++g1;
      };
      auto gemm2 = [&]() {                       // the whole context GEMM of the current pair
        TIMED_WAIT(wt_ef, mbar_wait(bar(B_EF), nu & 1));          // every E tile of this unit is in shared memory
        tc_fence_after();
        for (int c = 0; c < nchunk; ++c) {
          const uint32_t b2 = g2 & 1;
          TIMED_WAIT(wt_d2e, mbar_wait(bar(B_D2E + b2), ((g2 >> 1) & 1) ^ 1));
          tc_fence_after();
          for (int kb = 0; kb < nkb2; ++kb) {
            const int sa = take();
            tc_fence_after();
            const uint64_t bd = d_slot0 + (uint64_t)(sa * (SLOT >> 4));
            const uint64_t ed = d_e0 + (uint64_t)(kb * (KBLK / 8) * 64);        // 8 regions per 1024-byte atom
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + 2 * TILE + b2 * TILE, ed + 128 * k, bd + 2 * k, idesc2, (uint32_t)((kb | k) != 0));
            umma_commit(bar(B_EMPTY + sa));
          }
          umma_commit(bar(B_D2F + b2));
          ++g2;
        }
        umma_commit(bar(B_EE));                  // GEMM2 has finished reading E
        ++nu;
      };
      // Issue order: the tensor pipe runs MMAs in issue order, so the NEXT pair's first score tile is issued before
      // this pair's context GEMM (it fills the wait for the last softmax tile), its other tiles after it.
      UnitIter it(p.Bi, p.Bc);
      bool has = it.next();
      if (has)
        for (int t = 0; t < p.NT; ++t) gemm1();
      while (has) {
        const bool hasn = it.next();
        if (hasn) gemm1();
        gemm2();
        if (hasn)
          for (int t = 1; t < p.NT; ++t) gemm1();
        has = hasn;
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (p.dbg) {
        long long* d = p.dbg + (size_t)blockIdx.x * 32;
        d[0] = clock64() - t_all; d[1] = wt_full; d[2] = wt_d1e; d[3] = wt_ef; d[4] = wt_d2e; d[5] = nu;
      }
#endif
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ softmax warps (TMEM lanes = regions)
    const int q = warp & 3;
    const int row = q * 32 + lane;               // region row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t g1 = 0, nu = 0;
    Units u(p.Bi, p.Bc);
    while (u.next_caption()) {
      const int L = SEG ? LPAD : min(max(p.cap_lens[u.i], 0), LPAD);
      constexpr int NSEG = SEG ? LPAD / (SEG ? SEG : 1) : 1;
      int Lseg[NSEG];
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
        const int c = u.i * p.per + sg;
        Lseg[sg] = (SEG && sg < p.per && c < p.n_caps) ? min(max(p.cap_lens[c], 0), SEG) : 0;
      }
      for (int j = u.j; j < u.j_end; ++j) {
        for (int t = 0; t < p.NT; ++t) {
          const uint32_t b = g1 & 1;
          mbar_wait(bar(B_D1F + b), (g1 >> 1) & 1);
          tc_fence_after();
          float x[LPAD];
#pragma unroll
          for (int c = 0; c < LPAD / 16; ++c) tmem_ld16(tmem + lane_addr + b * TILE + c * 16, x + c * 16);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(bar(B_D1E + b));           // D1 buffer may be overwritten by the next tile's GEMM1
          ++g1;
          const int s_glob = t * TILE + row;
          const bool live_row = s_glob < p.S;
          if constexpr (SEG == 0) {
            // softmax #1 over the caption's words (gloria_loss.py:42-43), true running max
            float m = -INFINITY;
#pragma unroll
            for (int l = 0; l < LPAD; ++l) m = (l < L) ? fmaxf(m, x[l]) : m;
            const float mb = m * LOG2E;
            float sum = 0.f;
#pragma unroll
            for (int l = 0; l < LPAD; ++l) {
              const float e = (l < L) ? ex2(fmaf(x[l], LOG2E, -mb)) : 0.f;
              x[l] = e;
              sum += e;
            }
            // numerator of softmax #2 (:51-52): exp(temp1 * P); padded regions / words contribute nothing
            const float sc = p.t1_log2e / sum;
#pragma unroll
            for (int l = 0; l < LPAD; ++l) x[l] = (live_row && l < L) ? ex2(x[l] * sc) : 0.f;
          } else {
            // the same per 16-word segment = per caption of the tile
#pragma unroll
            for (int sg = 0; sg < NSEG; ++sg) {
              const int Ls = Lseg[sg];
              float* xs = x + sg * SEG;
              float m = -INFINITY;
#pragma unroll
              for (int l = 0; l < SEG; ++l) m = (l < Ls) ? fmaxf(m, xs[l]) : m;
              const float mb = m * LOG2E;
              float sum = 0.f;
#pragma unroll
              for (int l = 0; l < SEG; ++l) {
                const float e = (l < Ls) ? ex2(fmaf(xs[l], LOG2E, -mb)) : 0.f;
                xs[l] = e;
                sum += e;
              }
              const float sc = p.t1_log2e / sum;
#pragma unroll
              for (int l = 0; l < SEG; ++l) xs[l] = (live_row && l < Ls) ? ex2(xs[l] * sc) : 0.f;
            }
          }
          if (t == 0) mbar_wait(bar(B_EE), (nu & 1) ^ 1);   // previous unit's GEMM2 no longer reads E
          // MN-major SWIZZLE_128B A operand: [word block of 64][region][64 words], 16-B chunk ^= region % 8
          const uint32_t rowaddr = base + OFF_E + (uint32_t)(s_glob >> 3) * 1024u + (uint32_t)(s_glob & 7) * 128u;
#pragma unroll
          for (int wb = 0; wb < (LPAD + 63) / 64; ++wb) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int l0 = wb * 64 + c * 8;
              if (l0 < LPAD) {
                const uint32_t addr = rowaddr + (uint32_t)wb * ((uint32_t)Spad * 128u) + (uint32_t)((c ^ (s_glob & 7)) << 4);
                sts128(addr, pack_bf16(x[l0], x[l0 + 1]), pack_bf16(x[l0 + 2], x[l0 + 3]),
                       pack_bf16(x[l0 + 4], x[l0 + 5]), pack_bf16(x[l0 + 6], x[l0 + 7]));
              }
            }
          }
          fence_proxy_async_smem();              // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive(bar(B_EF));
        }
        ++nu;
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue warps (TMEM lanes = words)
    const int q = warp & 3;
    const int l = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int et = threadIdx.x - 256;            // 0..127
    uint32_t g2 = 0;
    Units u(p.Bi, p.Bc);
    while (u.next_caption()) {
      // packed mode: word l of the tile is word l % SEG of caption u.i * per + l / SEG
      const int cap = SEG ? u.i * p.per + l / (SEG ? SEG : 1) : u.i;
      const bool cap_ok = !SEG || (l < LPAD && l / (SEG ? SEG : 1) < p.per && cap < p.n_caps);
      const int L = !cap_ok ? 0 : min(max(p.cap_lens[cap], 0), SEG ? SEG : LPAD);
      const bool live = (SEG ? l % (SEG ? SEG : 1) : l) < L;
      const float nw = (l < LPAD) ? p.wnorm[(size_t)u.i * LPAD + l] : 0.f;
      const uint4* wrow = reinterpret_cast<const uint4*>(p.wt + ((size_t)u.i * LPAD + (l < LPAD ? l : 0)) * p.D);
      for (int j = u.j; j < u.j_end; ++j) {
        float dot = 0.f, c2 = 0.f;
        for (int c = 0; c < nchunk; ++c) {
          const uint32_t b2 = g2 & 1;
          mbar_wait(bar(B_D2F + b2), (g2 >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < TILE / 16; ++h) {
            uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
            if (live) {
              w0 = __ldg(wrow + (c * TILE + h * 16) / 8);
              w1 = __ldg(wrow + (c * TILE + h * 16) / 8 + 1);
            }
            float v[16];
            tmem_ld16(tmem + lane_addr + 2 * TILE + b2 * TILE + h * 16, v);
            tmem_ld_wait();
            const uint32_t wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&wr[k]));
              const float wa = w2.x, wb = w2.y;
              dot = fmaf(v[2 * k], wa, dot);
              dot = fmaf(v[2 * k + 1], wb, dot);
              c2 = fmaf(v[2 * k], v[2 * k], c2);
              c2 = fmaf(v[2 * k + 1], v[2 * k + 1], c2);
            }
          }
          tc_fence_before();
          mbar_arrive(bar(B_D2E + b2));
          ++g2;
        }
        if (p.stats != nullptr && l < LPAD) {       // saved for the backward (per word: <W, C'> and |C'|^2)
          float* sp = p.stats + ((size_t)j * p.Bc + u.i) * 2 * LPAD;
          sp[l] = live ? dot : 0.f;
          sp[LPAD + l] = live ? c2 : 0.f;
        }
        // cosine (gloria_loss.py:11-16) and the temp2 aggregation over words (:153-158 / gloria_model.py:198-201)
        const float den = fmaxf(nw * sqrtf(c2), p.eps_s);
        const float v = live ? p.temp2 * (dot / den) : -INFINITY;
        if constexpr (SEG != 0) {
          // aggregation over the 16 lanes of this word's segment
          float mx = v;
#pragma unroll
          for (int o = SEG / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          float ex = live ? __expf(v - mx) : 0.f;
#pragma unroll
          for (int o = SEG / 2; o > 0; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
          if (cap_ok && l % SEG == 0) {
            float r;
            if (p.agg == GLORIA_AGG_MAX) r = mx;
            else {
              r = mx + logf(ex);
              if (p.agg == GLORIA_AGG_MEAN) r -= logf((float)L);
            }
            p.sim[(size_t)j * p.n_caps + cap] = r;
          }
          continue;
        }
        float mx = warp_max(v);
        if (lane == 0) red[q] = mx;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        float ex = live ? __expf(v - mx) : 0.f;
        ex = warp_sum(ex);
        if (lane == 0) red[4 + q] = ex;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const float tot = red[4] + red[5] + red[6] + red[7];
          float r;
          if (p.agg == GLORIA_AGG_MAX) r = mx;
          else {
            r = mx + logf(tot);
            if (p.agg == GLORIA_AGG_MEAN) r -= logf((float)L);
          }
          p.sim[(size_t)j * p.Bc + u.i] = r;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // red[] reusable
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prepack: fp32 native layouts -> padded bf16 TMA-legal layouts
// ---------------------------------------------------------------------------------------------------------------
// ctx [Bi, D, S] -> Rn [Bi, D, Spad] (bf16), Rt [Bi, Spad, D] (bf16) and Rh [Bi, Spad, D] (fp16)
// grid (Spad/64, D/64, Bi), block (32, 8).  The score GEMM runs on the fp16 copies: the word softmax amplifies operand
// rounding (scores of unit-variance 768-d features have std 27.7) and fp16 carries 3 more mantissa bits than bf16 at
// the same tensor rate -- it is also the dtype the reference's AMP runs this bmm in.
__global__ void __launch_bounds__(256) pack_ctx(const float* __restrict__ ctx, __nv_bfloat16* __restrict__ Rn,
                                                 __nv_bfloat16* __restrict__ Rt, __half* __restrict__ Rh, int D, int S, int Spad,
                                                 int sp) {
  // one 64 (regions) x 64 (channels) tile per block: every thread first issues its 16 loads (two 128-byte rows per warp
  // and channel), then writes the s-major copy (bf16 pairs exchanged by shuffle: 128-byte rows) and, through shared
  // memory, 128-byte rows (2 channels per thread) along d of both transposed copies
  __shared__ float t[64][65];
  const int b = blockIdx.z, s0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float v[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* row = ctx + ((size_t)b * D + d0 + ty + 8 * k) * S;
    v[k][0] = (s0 + tx < S) ? __ldg(row + s0 + tx) : 0.f;
    v[k][1] = (s0 + 32 + tx < S) ? __ldg(row + s0 + 32 + tx) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k;
    t[r][tx] = v[k][0];
    t[r][tx + 32] = v[k][1];
    if (Rn != nullptr) {
      // lane pairs trade halves so that every lane stores one bf16 pair: even lanes write columns (tx, tx+1) of the first
      // 32, odd lanes columns (tx-1+32, tx+32) of the second
      const float mine = (tx & 1) ? v[k][1] : v[k][0];
      const float send = (tx & 1) ? v[k][0] : v[k][1];
      const float got = __shfl_xor_sync(0xffffffffu, send, 1);
      const int col = (tx & 1) ? s0 + 32 + tx - 1 : s0 + tx;
      const __nv_bfloat162 pr = (tx & 1) ? __floats2bfloat162_rn(got, mine) : __floats2bfloat162_rn(mine, got);
      *reinterpret_cast<__nv_bfloat162*>(Rn + ((size_t)b * D + d0 + r) * Spad + col) = pr;
    }
  }
  __syncthreads();
  const int d = d0 + 2 * tx;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k;
    const int so = s0 + r;
    const float v0 = t[2 * tx][r], v1 = t[2 * tx + 1][r];
    if (Rt != nullptr && so < sp)
      *reinterpret_cast<__nv_bfloat162*>(Rt + ((size_t)b * sp + so) * D + d) = __floats2bfloat162_rn(v0, v1);
    *reinterpret_cast<__half2*>(Rh + ((size_t)b * Spad + so) * D + d) = __floats2half2_rn(v0, v1);
  }
}

// words [Bc, D, Lw] -> Wt [Bc, lpb, D] bf16 and Wh [Bc, LPAD, D] fp16 (rows >= cap_len zero);
// grid (ceil(LPAD/32), D/32, Bc), block (32, 8): one 32 x 32 transpose tile per block
__global__ void pack_words(const float* __restrict__ words, const int* __restrict__ cap_lens,
                           __nv_bfloat16* __restrict__ Wt, __half* __restrict__ Wh, int D, int Lw, int lpad, int lpb,
                           int lcap, int off) {
  __shared__ float t[32][33];
  const int i = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int L = min(max(cap_lens[i], 0), lcap);
  const int lr = l0 + threadIdx.x;                 // word handled by this thread in the read phase
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int d = d0 + r;
    t[r][threadIdx.x] = (lr < L && d < D) ? words[((size_t)i * D + d) * Lw + off + lr] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int l = l0 + r, d = d0 + threadIdx.x;
    if (l < lpad && d < D) {
      if (l < lpb) Wt[((size_t)i * lpb + l) * D + d] = __float2bfloat16_rn(t[threadIdx.x][r]);   // GEMM pitch
      Wh[((size_t)i * lpad + l) * D + d] = __float2half_rn(t[threadIdx.x][r]);                  // tile pitch
    }
  }
}

// wnorm [Bc, LPAD] = |W_l| from the fp32 input (0 beyond the caption);  grid (ceil(LPAD/32), 1, Bc), block (32, 8):
// the 8 thread rows split the channels, reads are coalesced along the word axis
__global__ void word_norms(const float* __restrict__ words, const int* __restrict__ cap_lens, float* __restrict__ wnorm,
                           int D, int Lw, int lpad, int lcap, int off) {
  __shared__ float part[8][32];
  const int i = blockIdx.z, lr = blockIdx.x * 32 + threadIdx.x;
  const int L = min(max(cap_lens[i], 0), lcap);
  float ss = 0.f;
  if (lr < L) {
    const float* w = words + (size_t)i * D * Lw + off + lr;
#pragma unroll 8
    for (int d = threadIdx.y; d < D; d += 8) {
      const float v = __ldg(w + (size_t)d * Lw);
      ss = fmaf(v, v, ss);
    }
  }
  part[threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y == 0 && lr < lpad) {
    float tot = 0.f;
    for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x];
    wnorm[(size_t)i * lpad + lr] = sqrtf(tot);
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)ptr;
  }();
  return fn;
}

// 2-D bf16 tensor [rows, inner] (inner contiguous), box [box_rows, 64], SWIZZLE_128B
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {KBLK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GLORIA_OK;
}

int make_map3(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t mid, uint64_t rows, uint64_t mid_pitch,
              uint64_t row_pitch, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {inner, mid, rows};
  cuuint64_t strides[2] = {mid_pitch * 2, row_pitch * 2};
  cuuint32_t box[3] = {KBLK, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return GLORIA_OK;
}

int make_map4(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t mid, uint64_t rows, uint64_t imgs,
              uint64_t mid_pitch, uint64_t row_pitch, uint64_t img_pitch, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {inner, mid, rows, imgs};
  cuuint64_t strides[3] = {mid_pitch * 2, row_pitch * 2, img_pitch * 2};
  cuuint32_t box[4] = {KBLK, 1, box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GLORIA_ERR_DRIVER, "cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
  return GLORIA_OK;
}

template <int LPAD, int SEG = 0>
int launch_fwd(const CUtensorMap& rt, const CUtensorMap& wt, const CUtensorMap& rn, const FwdParams& p, int grid,
               cudaStream_t st) {
  GLORIA_CUDA(cudaFuncSetAttribute(tc_fwd_kernel<LPAD, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  timer_record(GLORIA_TIMER_TC_FWD, 0, st);
  tc_fwd_kernel<LPAD, SEG><<<grid, NTHREADS, SMEM_BYTES, st>>>(rt, wt, rn, p);
  timer_record(GLORIA_TIMER_TC_FWD, 1, st);
  GLORIA_LAUNCHED("tc_fwd_kernel");
  return GLORIA_OK;
}

}  // namespace tc
}  // namespace gloria

using namespace gloria;
using namespace gloria::tc;

// Development aid: device buffer of per-CTA phase clocks (8 x int64 per CTA), used only by -DGLORIA_PHASE_CLOCKS builds.
void* gloria::tc::g_phase_clock_buffer = nullptr;
extern "C" void gloria_b200_debug_phase_clocks(void* device_buffer) { gloria::tc::g_phase_clock_buffer = device_buffer; }

extern "C" int gloria_b200_tc_spad(int S) { return (S / TILE + 1) * TILE; }
extern "C" int gloria_b200_tc_lpad(int Lcap) { return (Lcap + 15) / 16 * 16; }
extern "C" int gloria_b200_tc_lp(int Lcap) { return (Lcap + 7) / 8 * 8; }
extern "C" int gloria_b200_tc_sp(int S) { return (S + 15) / 16 * 16; }

extern "C" int gloria_b200_tc_supported(int D, int S, int Lcap) {
  if (D < TILE || D % TILE != 0 || S < 1 || S >= MAX_NT * TILE || Lcap < 1 || Lcap > TILE) return GLORIA_ERR_UNSUPPORTED;
  return GLORIA_OK;
}

// The two halves of the prepack are separate entry points so that a caption-sharded caller can pack image parts as
// their all_gather lands (ctx_n = NULL skips the copy only the inference forward / recompute backward read, ctx_t = NULL
// the copy only the training path reads).
extern "C" int gloria_b200_tc_prepack_ctx(const float* ctx, int Bi, int D, int S, void* ctx_h, void* ctx_t, void* ctx_n,
                                          void* stream) {
  GLORIA_CHECK_ARG(ctx && ctx_h, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0, "bad sizes");
  if (gloria_b200_tc_supported(D, S, 1)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d", D, S);
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S);
  pack_ctx<<<dim3(Spad / 64, D / 64, Bi), dim3(32, 8), 0, st>>>(ctx, (__nv_bfloat16*)ctx_n, (__nv_bfloat16*)ctx_t,
                                                               (__half*)ctx_h, D, S, Spad, gloria_b200_tc_sp(S));
  GLORIA_LAUNCHED("pack_ctx");
  return GLORIA_OK;
}

extern "C" int gloria_b200_tc_prepack_words(const float* words, const int32_t* cap_lens, int Bc, int D, int Lw, int Lcap,
                                            int word_off, void* words_h, void* words_t, float* wnorm, void* stream) {
  GLORIA_CHECK_ARG(words && cap_lens && words_h && words_t && wnorm, "null pointer");
  GLORIA_CHECK_ARG(Bc > 0 && word_off >= 0 && word_off + Lcap <= Lw, "bad sizes");
  if (gloria_b200_tc_supported(D, 1, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d Lcap=%d", D, Lcap);
  cudaStream_t st = (cudaStream_t)stream;
  const int lpad = gloria_b200_tc_lpad(Lcap);
  pack_words<<<dim3((lpad + 31) / 32, D / 32, Bc), dim3(32, 8), 0, st>>>(words, cap_lens, (__nv_bfloat16*)words_t,
                                                                         (__half*)words_h, D, Lw, lpad,
                                                                         gloria_b200_tc_lp(Lcap), Lcap, word_off);
  GLORIA_LAUNCHED("pack_words");
  word_norms<<<dim3((lpad + 31) / 32, 1, Bc), dim3(32, 8), 0, st>>>(words, cap_lens, wnorm, D, Lw, lpad, Lcap, word_off);
  GLORIA_LAUNCHED("word_norms");
  return GLORIA_OK;
}

extern "C" int gloria_b200_tc_prepack(const float* ctx, const float* words, const int32_t* cap_lens, int Bi, int Bc,
                                      int D, int S, int Lw, int Lcap, int word_off, void* ctx_h, void* ctx_t,
                                      void* ctx_n, void* words_h, void* words_t, float* wnorm, void* stream) {
  GLORIA_CHECK_ARG(ctx && words && cap_lens && ctx_h && ctx_t && ctx_n && words_h && words_t && wnorm, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && word_off >= 0 && word_off + Lcap <= Lw, "bad sizes");
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  int rc;
  if ((rc = gloria_b200_tc_prepack_ctx(ctx, Bi, D, S, ctx_h, ctx_t, ctx_n, stream))) return rc;
  return gloria_b200_tc_prepack_words(words, cap_lens, Bc, D, Lw, Lcap, word_off, words_h, words_t, wnorm, stream);
}

extern "C" int gloria_b200_tc_local_sim_fwd(const void* ctx_h, const void* ctx_n, const void* words_h,
                                            const float* wnorm, const int32_t* cap_lens, int Bi, int Bc, int D, int S,
                                            int Lcap, float temp1, float temp2, int agg, float eps, float* sim,
                                            float* stats, void* stream) {
  GLORIA_CHECK_ARG(ctx_h && ctx_n && words_h && wnorm && cap_lens && sim, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0, "bad batch sizes %d x %d", Bi, Bc);
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S), lpad = gloria_b200_tc_lpad(Lcap);
  CUtensorMap rt, wt, rn;
  int rc;
  if ((rc = make_map(&rt, ctx_h, (uint64_t)D, (uint64_t)Bi * Spad, TILE))) return rc;      // 2-byte elements: the map
  if ((rc = make_map(&wt, words_h, (uint64_t)D, (uint64_t)Bc * lpad, (uint32_t)lpad))) return rc;   // is dtype-agnostic
  if ((rc = make_map(&rn, ctx_n, (uint64_t)Spad, (uint64_t)Bi * D, TILE))) return rc;
  FwdParams p;
  p.wt = (const __half*)words_h; p.wnorm = wnorm; p.cap_lens = cap_lens; p.sim = sim; p.stats = stats;
  p.Bi = Bi; p.Bc = Bc; p.D = D; p.S = S; p.NT = Spad / TILE; p.n_caps = Bc; p.per = 1;
  p.t1_log2e = temp1 * 1.4426950408889634f; p.temp2 = temp2; p.agg = agg; p.eps_s = eps * (float)S;
  p.dbg = (long long*)g_phase_clock_buffer;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms;
  switch (lpad) {
    case 16: return launch_fwd<16>(rt, wt, rn, p, grid, st);
    case 32: return launch_fwd<32>(rt, wt, rn, p, grid, st);
    case 48: return launch_fwd<48>(rt, wt, rn, p, grid, st);
    case 64: return launch_fwd<64>(rt, wt, rn, p, grid, st);
    case 80: return launch_fwd<80>(rt, wt, rn, p, grid, st);
    case 96: return launch_fwd<96>(rt, wt, rn, p, grid, st);
    case 112: return launch_fwd<112>(rt, wt, rn, p, grid, st);
    case 128: return launch_fwd<128>(rt, wt, rn, p, grid, st);
  }
  return fail(GLORIA_ERR_UNSUPPORTED, "lpad %d", lpad);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed prompts (zero-shot scoring, gloria_model.py:171-207 driven by gloria.py:278-306: thousands of images against a
// few short class prompts).  `per` = 1..8 captions of at most 16 words share one word tile of lpad = 16 * per words.
// ---------------------------------------------------------------------------------------------------------------
namespace gloria {
namespace tc {
// one block per (caption slot, 64 channels): Wh[g, (slot % per) * 16 + l, d] = words[c, d, off + l] (0 beyond the caption /
// for empty slots)
__global__ void pack_words_packed(const float* __restrict__ words, const int* __restrict__ cap_lens,
                                  __half* __restrict__ Wh, int n_caps, int per, int D, int Lw, int lpad, int off) {
  const int slot = blockIdx.x, g = slot / per, sg = slot % per;
  const int d = blockIdx.y * 64 + threadIdx.x;
  const int L = slot < n_caps ? min(max(cap_lens[slot], 0), 16) : 0;
  if (d >= D) return;
  for (int l = 0; l < 16; ++l) {
    const float v = l < L ? words[((size_t)slot * D + d) * Lw + off + l] : 0.f;
    Wh[((size_t)g * lpad + sg * 16 + l) * D + d] = __float2half_rn(v);
  }
}
// wnorm[g, (slot % per) * 16 + l] = |words[c, :, off + l]|;  one warp per (slot, l)
__global__ void word_norms_packed(const float* __restrict__ words, const int* __restrict__ cap_lens,
                                  float* __restrict__ wnorm, int n_caps, int per, int D, int Lw, int lpad, int off) {
  const int slot = blockIdx.x, l = threadIdx.y, lane = threadIdx.x;
  const int L = slot < n_caps ? min(max(cap_lens[slot], 0), 16) : 0;
  float ss = 0.f;
  if (l < L)
    for (int d = lane; d < D; d += 32) {
      const float v = words[((size_t)slot * D + d) * Lw + off + l];
      ss = fmaf(v, v, ss);
    }
  ss = warp_sum(ss);
  if (lane == 0) wnorm[(size_t)(slot / per) * lpad + (slot % per) * 16 + l] = sqrtf(ss);
}
}  // namespace tc
}  // namespace gloria

extern "C" int gloria_b200_tc_packed_groups(int Bc) { return Bc <= 0 ? 0 : (Bc + 7) / 8; }
extern "C" int gloria_b200_tc_packed_per(int Bc) {
  const int g = gloria_b200_tc_packed_groups(Bc);
  return g == 0 ? 0 : (Bc + g - 1) / g;
}

extern "C" int gloria_b200_tc_prepack_words_packed(const float* words, const int32_t* cap_lens, int Bc, int D, int Lw,
                                                   int word_off, void* words_h, float* wnorm, void* stream) {
  GLORIA_CHECK_ARG(words && cap_lens && words_h && wnorm, "null pointer");
  GLORIA_CHECK_ARG(Bc > 0 && word_off >= 0 && word_off + 1 <= Lw, "bad sizes");
  if (gloria_b200_tc_supported(D, 1, 16)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d", D);
  cudaStream_t st = (cudaStream_t)stream;
  const int per = gloria_b200_tc_packed_per(Bc), G = gloria_b200_tc_packed_groups(Bc), lpad = 16 * per;
  pack_words_packed<<<dim3(G * per, (D + 63) / 64), 64, 0, st>>>(words, cap_lens, (__half*)words_h, Bc, per, D, Lw, lpad,
                                                               word_off);
  GLORIA_LAUNCHED("pack_words_packed");
  word_norms_packed<<<G * per, dim3(32, 16), 0, st>>>(words, cap_lens, wnorm, Bc, per, D, Lw, lpad, word_off);
  GLORIA_LAUNCHED("word_norms_packed");
  return GLORIA_OK;
}

extern "C" int gloria_b200_tc_local_sim_fwd_packed(const void* ctx_h, const void* ctx_n, const void* words_h,
                                                   const float* wnorm, const int32_t* cap_lens, int Bi, int Bc, int D,
                                                   int S, float temp1, float temp2, int agg, float eps, float* sim,
                                                   void* stream) {
  GLORIA_CHECK_ARG(ctx_h && ctx_n && words_h && wnorm && cap_lens && sim, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0, "bad batch sizes %d x %d", Bi, Bc);
  if (gloria_b200_tc_supported(D, S, 16)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d", D, S);
  cudaStream_t st = (cudaStream_t)stream;
  const int per = gloria_b200_tc_packed_per(Bc), G = gloria_b200_tc_packed_groups(Bc), lpad = 16 * per;
  const int Spad = gloria_b200_tc_spad(S);
  CUtensorMap rt, wt, rn;
  int rc;
  if ((rc = make_map(&rt, ctx_h, (uint64_t)D, (uint64_t)Bi * Spad, TILE))) return rc;
  if ((rc = make_map(&wt, words_h, (uint64_t)D, (uint64_t)G * lpad, (uint32_t)lpad))) return rc;
  if ((rc = make_map(&rn, ctx_n, (uint64_t)Spad, (uint64_t)Bi * D, TILE))) return rc;
  FwdParams p;
  p.wt = (const __half*)words_h; p.wnorm = wnorm; p.cap_lens = cap_lens; p.sim = sim; p.stats = nullptr;
  p.Bi = Bi; p.Bc = G; p.D = D; p.S = S; p.NT = Spad / TILE; p.n_caps = Bc; p.per = per;
  p.t1_log2e = temp1 * 1.4426950408889634f; p.temp2 = temp2; p.agg = agg; p.eps_s = eps * (float)S;
  p.dbg = (long long*)g_phase_clock_buffer;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  switch (lpad) {
    case 16: return launch_fwd<16, 16>(rt, wt, rn, p, sms, st);
    case 32: return launch_fwd<32, 16>(rt, wt, rn, p, sms, st);
    case 48: return launch_fwd<48, 16>(rt, wt, rn, p, sms, st);
    case 64: return launch_fwd<64, 16>(rt, wt, rn, p, sms, st);
    case 80: return launch_fwd<80, 16>(rt, wt, rn, p, sms, st);
    case 96: return launch_fwd<96, 16>(rt, wt, rn, p, sms, st);
    case 112: return launch_fwd<112, 16>(rt, wt, rn, p, sms, st);
    case 128: return launch_fwd<128, 16>(rt, wt, rn, p, sms, st);
  }
  return fail(GLORIA_ERR_UNSUPPORTED, "lpad %d", lpad);
}

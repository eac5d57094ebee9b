// bf16 tensor-core backward of the GLoRIA local similarity (sm_100a: TMA + tcgen05 + TMEM, plus plain GEMMs).
//
// Closed-form backward of gloria_loss.py:19-63 + :144-158, restated so that nothing of size
// D x S per pair is ever accumulated with atomics.  Per pair (image j, caption i), with E = exp(temp1 * P) the
// un-normalised softmax-#2 numerator, Z_l = sum_s E[s,l], A = E / Z, and G_j = R_j^T R_j the image's Gram matrix:
//     T'[s,l]  = sum_s' G_j[s,s'] E[s',l]              (= Z_l * <C_l, R_s>;  replaces the D-deep GEMM  dC^T R)
//     dP[s,l]  = E * (a_l S_[s,l] + b_l T'[s,l] - c_l)  a = t1 ddot/Z, b = t1 beta/Z^2, c = t1 (ddot dot + beta nc^2)/Z
//     u_s      = sum_l P dP ;   dS = P (dP - u_s)
//     X[s,l]   = (ddot_l / Z_l) E + dS                  -> dW_i += R_j X ,  dR_j += W_i X^T
//     Eo = E,  Bo[s,l] = (beta_l / Z_l^2) E             -> dR_j += R_j (sum_i Eo Bo^T)      (the |C| term)
// ddot / beta / gamma come from the forward's saved per-word <W, C'> and |C'|^2 (C' = Z C) and from Z, which the
// score-shaped GEMM  G E  delivers for free through a row of ones kept in G's first padded row.
//
// One fused kernel per chunk of captions ("pair kernel", caption-stationary persistent CTAs like the forward):
//   GEMM1  S_[s,l] (3 region tiles, TMEM lanes = regions, stays resident)  ->  softmax warps: P, E -> smem (bf16)
//   GEMM-T T'[s,l] = G tile (TMA) x E (smem, N-major)                      ->  same lanes/columns as S_
//   12 SIMT warps (3 word-column groups x 4 lane quarters) then do everything else with one region row per thread
//   (row reductions exchanged between the column groups through shared memory) and write the rows of
//   X^T, Eo^T, Bo^T as [(j,s), (i,l)] bf16 matrices.
// The sums over images / captions are then three large plain GEMMs over those matrices (cuBLAS, the one place a
// library GEMM is used) and a per-image [S x S] x [S x D] product; see DESIGN.md for the byte / FLOP accounting.
#include <cublas_v2.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gloria {
namespace tc {
namespace bw {

constexpr int SLOT = 16384;
constexpr int NSLOT = 7;
constexpr int NGROUP = 3;                                // column groups of SIMT warps
constexpr int OFF_E = NSLOT * SLOT;                      // [2 word blocks of 64][Spad regions][128 B]
constexpr int E_BYTES = 2 * MAX_NT * TILE * 128;
constexpr int NCOL = 144;                                // per-column arrays: 3 groups x 6 chunks x 8 (phantom chunks included)
constexpr int OFF_COEFA = OFF_E + E_BYTES;               // float4 (a, b, c', -) per word: dP = E (a S_ + b T' + c')
constexpr int OFF_COEFX = OFF_COEFA + NCOL * 16;         // float e per word (X = e E + dS)
constexpr int OFF_COEFAX = OFF_COEFX + NCOL * 4;         // float a per word again, dense (the S_ half of the element pass)
constexpr int OFF_NEGM = OFF_COEFAX + NCOL * 4;          // float 0 / -inf per word (beyond the caption)
constexpr int OFF_CMASK = OFF_NEGM + NCOL * 4;           // uint16 0xFFFF / 0 per word
constexpr int OFF_ZBUF = OFF_CMASK + NCOL * 2 + 32;      // Z per word
constexpr int OFF_RED = OFF_ZBUF + 128 * 4;              // 32 floats of reduction scratch
constexpr int OFF_ACCS = OFF_RED + 128;                  // FUSED: float [2][NCOL] column sums (dot', |C'|^2) of the pair
constexpr int OFF_XCH = OFF_ACCS + 2 * NCOL * 4;         // float2 [2][NGROUP][128] row-reduction exchange
constexpr int OFF_BAR = OFF_XCH + 2 * NGROUP * 128 * 8;
enum { B_FULL = 0, B_EMPTY = NSLOT, B_D1F = 2 * NSLOT, B_D1E = B_D1F + MAX_NT, B_EF = B_D1E + MAX_NT, B_EE, B_TTF, B_TTE, B_COUNT };
constexpr int SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16 + 1024;
constexpr int NTHREADS = 512;                            // warp 0 TMA, warp 1 MMA, warps 4-15 SIMT (lanes = regions)
constexpr int NSIMT = NTHREADS - 128;

struct PairParams {
  const float* wnorm;       // [Bc, LPAD]
  const int* cap_lens;      // [Bc]
  const float* stats;       // [Bi, Bc, 2, LPAD]
  const float* dsim;        // [Bi, Bc]
  __nv_bfloat16* xt;        // [Bi*Spad, nc*LPAD]   X^T
  __nv_bfloat16* et;        //                      E^T (un-normalised attention numerators)
  float* fo;                // [Bi, nc*LPAD]  f = beta / Z^2 per pair and word (Bo = f * Eo is formed by scale_rows)
  float* go;                // FUSED: [Bi, nc*LPAD]  gamma per pair and word for g = 1
  float* sim;               // FUSED: [Bi, Bc]  the forward's similarities
  int agg;                  // FUSED: GLORIA_AGG_SUM / _MEAN
  float* gamma;             // [Bc, LPAD]  sum_j (dL/d|W_l|) / |W_l|   (atomicAdd)
  // word-mean attention (regularisers of gloria_loss.py:108-139): mean[j,i,s] = (1/L_i) sum_l E[s,l] / Z_l
  float* mean_out;          // FUSED with xt == nullptr ("lean" forward): [Bi, Bc, S]
  float* stats_out;         // FUSED: [Bi, Bc, 2, LPAD]  <W,C'>, |C'|^2 per word for a later recompute backward (or null)
  const float* dmean;       // !FUSED: [Bi, Bc, S]  dL/d mean (or null)
  int Bi, Bc, i0, nc, D, S, NT;
  int lp;                   // column pitch per caption of X^T / E^T / fo / go: round_up(Lcap, 8) <= LPAD
  int sp;                   // region rows per image in X^T / E^T: round_up(S, 16) <= Spad
  float t1, t1_log2e, t2, eps;
  long long* dbg;           // phase clocks (only read when built with -DGLORIA_PHASE_CLOCKS)
  int timer_first = 1, timer_last = 1;   // host side only: which ends of the launch the bench timer slot records
  int l2_hints = 0;         // bit 0: X^T / E^T stores evict-first; bit 1: operand tile loads evict-last
  // FUSED: attention maps of the diagonal pairs (att_maps of local_loss, gloria_loss.py:141-143).  The softmax warps of the
  // pair (image diag_j0 + j == caption i0 + i) store their fp32 numerators E[s,l] = exp(temp1 P) here, [Bc, diag_lcap, S];
  // `normalise_diag` then divides each row by its sum.  null = not wanted.
  float* diag_raw = nullptr;
  int diag_lcap = 0, diag_j0 = 0;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Column sums over the 32 lanes of a warp for N per-lane values (N a multiple of 8): halving butterfly -- at each
// step a lane keeps one half of its remaining columns and trades the other half with its partner (N + N/2 + ... ~ 2N
// shuffles instead of 5N); the survivors are added into dst[] (shared memory, one atomic per surviving value).
template <int N, int STEP>
__device__ __forceinline__ void warp_colsum_step(float* v, int lane, int lo, float* dst, int total) {
  if constexpr (STEP == 5) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (lo + i < total) atomicAdd(dst + lo + i, v[i]);
  } else {
    constexpr int H = (N + 1) / 2;                  // columns kept
    constexpr int MASK = 16 >> STEP;
    const bool up = (lane & MASK) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float hi = (i + H < N) ? v[i + H] : 0.f;
      const float send = up ? v[i] : hi;
      const float keep = up ? hi : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
    warp_colsum_step<H, STEP + 1>(v, lane, lo + (up ? H : 0), dst, total);
  }
}
template <int N>
__device__ __forceinline__ void warp_colsum(float* v, int lane, float* dst) {
  warp_colsum_step<N, 0>(v, lane, 0, dst, N);
}

// region tile processed at position idx of a pair: the tile holding the ones row of G (the last one) goes first
__device__ __forceinline__ int tile_at(int idx, int NT) { return idx == 0 ? NT - 1 : idx - 1; }
// FUSED: pass A runs in tile_at order; the softmax, pass B and the next pair's GEMM1 run in an order that STARTS with
// pass A's last tile, whose T' is still in TMEM when the coefficients are ready (one GEMM-T per pair saved)
template <bool FUSED>
__device__ __forceinline__ int tile_ord(int idx, int NT) {
  if (!FUSED) return tile_at(idx, NT);
  int k = idx + NT - 1;
  if (k >= NT) k -= NT;
  return tile_at(k, NT);
}

// FUSED = false: backward by recomputation (needs the forward's stats and dsim).
// FUSED = true : training forward.  Every backward quantity is linear in g = dsim[j,i], so this mode computes sim AND
//   the backward operand rows for g = 1 in one pass; the backward proper is then a scale by g plus the GEMMs.  The
//   per-word <W,C'> and |C'|^2 are reduced in-kernel (Gram form: sum_s E S_ and sum_s E T'), which takes a first
//   GEMM-T pass (pass A) before the coefficients exist; GEMM-T is then issued again for the element pass (pass B).
template <int LPAD, bool FUSED>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_bwd_pair_kernel(const __grid_constant__ CUtensorMap tm_rt, const __grid_constant__ CUtensorMap tm_wt,
                   const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_e,
                   const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
#ifdef GLORIA_PHASE_CLOCKS
  long long* g_dbg = p.dbg;
#endif
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* zbuf = reinterpret_cast<float*>(smem + OFF_ZBUF);
  float* accs = reinterpret_cast<float*>(smem + OFF_ACCS);
  float4* coefA = reinterpret_cast<float4*>(smem + OFF_COEFA);
  float* coefX = reinterpret_cast<float*>(smem + OFF_COEFX);
  float* coefAx = reinterpret_cast<float*>(smem + OFF_COEFAX);
  float* negm = reinterpret_cast<float*>(smem + OFF_NEGM);
  uint32_t* cmask = reinterpret_cast<uint32_t*>(smem + OFF_CMASK);
  auto bar = [&](int idx) { return bars + 8u * (uint32_t)idx; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT;
  const int Spad = NT * TILE;
  const int nkb1 = p.D / KBLK;        // k-blocks of GEMM1 (over channels)
  const int nkb2 = Spad / KBLK;       // k-blocks of GEMM-T (over regions)
  const uint32_t TT_COL = (uint32_t)(MAX_NT * LPAD);
  // "lean" forward: sim (+ word-mean attention, per-word stats) only -- no operand rows, so no element pass and no
  // pass-B GEMM-T
  const bool lean = FUSED && p.xt == nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); }
    for (int t = 0; t < MAX_NT; ++t) { mbar_init(bar(B_D1F + t), 1); mbar_init(bar(B_D1E + t), NSIMT); }
    mbar_init(bar(B_EF), NSIMT * NT);
    mbar_init(bar(B_EE), 2);            // GEMM-T done reading E (tcgen05.commit) + Eo bulk stores done reading E
    mbar_init(bar(B_TTF), 1);
    mbar_init(bar(B_TTE), NSIMT);
    fence_barrier_init();
    tma_prefetch_desc(&tm_rt); tma_prefetch_desc(&tm_wt); tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_e);
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
#ifdef GLORIA_PHASE_CLOCKS
      long long pw_empty = 0, pt_all = clock64();
#define PTIMED(acc, ...) do { long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; } while (0)
#else
#define PTIMED(acc, ...) do { __VA_ARGS__; } while (0)
#endif
      const bool keep = (p.l2_hints & 2) != 0;
      const uint64_t pol_keep = l2_policy_evict_last();
      auto load = [&](const CUtensorMap* tm, int x, int y, uint32_t bytes) {
        PTIMED(pw_empty, mbar_wait(bar(B_EMPTY + slot), ph ^ 1));
        mbar_expect_tx(bar(B_FULL + slot), bytes);
        if (keep) tma_load_2d_hint(base + slot * SLOT, tm, x, y, bar(B_FULL + slot), pol_keep);
        else tma_load_2d(base + slot * SLOT, tm, x, y, bar(B_FULL + slot));
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
      };
      // load order mirrors the MMA issuer: GEMM1 of the first pair, then per pair: G tiles of GEMM-T tile k followed
      // by the GEMM1 operands of the NEXT pair's tile k-1 (see the issuer for why)
      auto load_g1 = [&](int i_loc, int j, int t) {
        for (int kb = 0; kb < nkb1; ++kb) {
          load(&tm_rt, kb * KBLK, j * Spad + t * TILE, TILE * 128);
          load(&tm_wt, kb * KBLK, (p.i0 + i_loc) * LPAD, LPAD * 128);
        }
      };
      auto load_tt = [&](int j, int t) {
        for (int kb = 0; kb < nkb2; ++kb) load(&tm_g, kb * KBLK, j * Spad + t * TILE, TILE * 128);
      };
      UnitIter it(p.Bi, p.nc);
      bool has = it.next();
      int ci = it.cap(), cj = it.j;
      if (has)
        for (int idx = 0; idx < NT; ++idx) load_g1(ci, cj, tile_ord<FUSED>(idx, NT));
      while (has) {
        const bool hasn = it.next();
        const int ni = it.cap(), nj = it.j;
        if (FUSED)
          for (int k = 0; k < NT; ++k) load_tt(cj, tile_at(k, NT));      // pass A (its last T' tile is pass B's first)
        else
          load_tt(cj, tile_at(0, NT));
        for (int k = 1; k < NT; ++k) {
          if (!lean) load_tt(cj, tile_ord<FUSED>(k, NT));
          if (hasn) load_g1(ni, nj, tile_ord<FUSED>(k - 1, NT));
        }
        if (hasn) load_g1(ni, nj, tile_ord<FUSED>(NT - 1, NT));
        has = hasn; ci = ni; cj = nj;
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (g_dbg) { long long* d = g_dbg + (size_t)blockIdx.x * 32 + 8; d[0] = clock64() - pt_all; d[1] = pw_empty; }
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(TILE, LPAD, 0, 0, 0);   // fp16: A = Rh tile (K-major), B = Wh tile (K-major)
      constexpr uint32_t idesct = make_idesc(TILE, LPAD, 0, 1);   // A = G tile (K-major),  B = E (N-major)
      const uint32_t e_lbo = (uint32_t)Spad * 128u;               // between the two 64-word blocks of E
      int slot = 0; uint32_t ph = 0;
      uint32_t n = 0, ttc = 0;
#ifdef GLORIA_PHASE_CLOCKS
      long long wt_full = 0, wt_d1e = 0, wt_ef = 0, wt_tte = 0, t_all = clock64();
#define TIMED_WAIT(acc, ...) do { long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; } while (0)
#else
#define TIMED_WAIT(acc, ...) do { __VA_ARGS__; } while (0)
#endif
      auto take = [&]() {
        TIMED_WAIT(wt_full, mbar_wait(bar(B_FULL + slot), ph));
        const int s = slot;
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
        return s;
      };
      const uint64_t d_slot0 = make_smem_desc(base, 16, 1024);            // K-major tile in slot 0
      const uint64_t d_e0 = make_smem_desc(base + OFF_E, e_lbo, 1024);    // E, N-major
      const uint64_t pol_stream = l2_policy_evict_first();
      // GEMM1 of one S_ tile of pair number `pn` (waits until the SIMT warps are done with that tile of pair pn-1)
      auto gemm1 = [&](uint32_t pn, int t) {
        TIMED_WAIT(wt_d1e, mbar_wait(bar(B_D1E + t), (pn & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < nkb1; ++kb) {
          const int sa = take();
          const int sb = take();
          tc_fence_after();
          const uint64_t ad = d_slot0 + (uint64_t)(sa * (SLOT >> 4)), bd = d_slot0 + (uint64_t)(sb * (SLOT >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + (uint32_t)(t * LPAD), ad + 2 * k, bd + 2 * k, idesc1, (uint32_t)((kb | k) != 0));
          umma_commit(bar(B_EMPTY + sa));
          umma_commit(bar(B_EMPTY + sb));
        }
        umma_commit(bar(B_D1F + t));
      };
      // GEMM-T of one tile (T' buffer must have been read by the SIMT warps)
      auto gemmt = [&]() {
        TIMED_WAIT(wt_tte, mbar_wait(bar(B_TTE), (ttc & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < nkb2; ++kb) {
          const int sa = take();
          tc_fence_after();
          const uint64_t ad = d_slot0 + (uint64_t)(sa * (SLOT >> 4));
          const uint64_t ed = d_e0 + (uint64_t)(kb * (KBLK / 8) * 64);          // 8 regions per 1024-byte atom
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TT_COL, ad + 2 * k, ed + 128 * k, idesct, (uint32_t)((kb | k) != 0));
          umma_commit(bar(B_EMPTY + sa));
        }
        umma_commit(bar(B_TTF));
        ++ttc;
      };
      // Issue order.  The tensor pipe runs MMAs in issue order, so the next pair's GEMM1 is interleaved with this
      // pair's GEMM-T tiles: the SIMT event that frees the T' buffer for tile k (pass 1 of tile k-1 done) also frees
      // S_ tile k-1, whose GEMM1 for the next pair is issued right behind GEMM-T tile k.  By the time the SIMT warps
      // finish this pair, the next pair's scores are already in TMEM.
      UnitIter it(p.Bi, p.nc);
      bool has = it.next();
      int ci = it.cap(), cj = it.j;
      if (has)
        for (int idx = 0; idx < NT; ++idx) gemm1(0, tile_ord<FUSED>(idx, NT));
      while (has) {
        const bool hasn = it.next();
        const int ni = it.cap(), nj = it.j;
        TIMED_WAIT(wt_ef, mbar_wait(bar(B_EF), n & 1));      // E of this pair is complete in shared memory
        tc_fence_after();
        // Eo rows: straight from the E buffer by bulk tensor stores (columns >= LPAD are clipped by the map)
        if (p.et != nullptr) {
          for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int wb = 0; wb < (LPAD + 63) / 64; ++wb) {
              const uint32_t src = base + OFF_E + (uint32_t)wb * e_lbo + (uint32_t)t * (TILE * 128);
              if (p.l2_hints & 1) tma_store_4d_hint(&tm_e, src, wb * 64, ci, t * TILE, cj, pol_stream);
              else tma_store_4d(&tm_e, src, wb * 64, ci, t * TILE, cj);
            }
          tma_store_commit();
        }
        if (FUSED)
          for (int k = 0; k < NT; ++k) gemmt();              // pass A: T' for the in-kernel |C'|^2 reduction; the last
        else                                                 // tile stays in TMEM as pass B's first
          gemmt();
        for (int k = 1; k < NT; ++k) {
          if (!lean) gemmt();
          if (hasn) gemm1(n + 1, tile_ord<FUSED>(k - 1, NT));
        }
        umma_commit(bar(B_EE));                              // GEMM-T has finished reading E
        tma_store_wait_read();                               // ... and so have the Eo stores (issued long ago)
        mbar_arrive(bar(B_EE));
        if (hasn) gemm1(n + 1, tile_ord<FUSED>(NT - 1, NT));
        ++n;
        has = hasn; ci = ni; cj = nj;
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (g_dbg) {
        long long* d = g_dbg + (size_t)blockIdx.x * 32;
        d[0] = clock64() - t_all; d[1] = wt_full; d[2] = wt_d1e; d[3] = wt_ef; d[4] = wt_tte; d[5] = n;
      }
#endif
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ SIMT warps (TMEM lanes = regions)
    // 12 warps = 3 column groups x 4 lane quarters: thread (g, row) owns region row `row` of every tile and CW
    // consecutive 8-word chunks.  Row-wise reductions (softmax max / sum, u) are exchanged between the 3 groups
    // through shared memory.  There are no per-element predicates: words beyond the caption are handled by data
    // (a -inf bias per column, zero coefficients, halfword masks), chunks beyond LPAD ("phantom" chunks of the last
    // group) compute on masked garbage and only their stores are skipped.
    constexpr int NCH = LPAD / 8;                  // 16-byte chunks (8 words) per row
    constexpr int CW = (NCH + NGROUP - 1) / NGROUP;   // chunks per column group
    constexpr int WG = CW * 8;
    static_assert(NGROUP * WG <= NCOL, "per-column arrays too small");
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // region row inside a tile == TMEM lane
    const int wl = (warp - 4) * 32 + lane;         // 0..383: column index for the per-column arrays
    const int c_lo = g * CW;
    const int col0 = c_lo * 8;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    const int zrow = p.S - (NT - 1) * TILE;        // row of the ones row of G inside the last tile
    float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH);     // [2][NGROUP][128] exchange buffer
    const size_t e_blk = (size_t)Spad * 128u;      // bytes between the two 64-word blocks of E
    const float4* cA = coefA + col0;
    const float* cX = coefX + col0;
    const float* cAx = coefAx + col0;
    const float* nM = negm + col0;
    const uint32_t* cM = cmask + (col0 >> 1);
    uint32_t n = 0, ttc = 0, xc = 0;
    const bool stream_x = (p.l2_hints & 1) != 0;
#ifdef GLORIA_PHASE_CLOCKS
    long long sw_d1f = 0, sw_ttf = 0, sw_ee = 0, sw_bar = 0, st_all = clock64();
#define STIMED(acc, ...) do { long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; } while (0)
#else
#define STIMED(acc, ...) do { __VA_ARGS__; } while (0)
#endif
    Units u(p.Bi, p.nc);
    while (u.next_caption()) {
      const int i = p.i0 + u.i;
      const int L = min(max(p.cap_lens[i], 0), LPAD);
      const float nw = (wl < LPAD) ? p.wnorm[(size_t)i * LPAD + wl] : 0.f;
      // per-column masks of this caption (all SIMT threads are past the previous caption's last use: the barriers
      // of its last tile precede this point, and the barrier of the first softmax tile follows it)
      if (wl < NCOL) {
        negm[wl] = (wl < L) ? 0.f : -INFINITY;
        reinterpret_cast<uint16_t*>(cmask)[wl] = (wl < L) ? (uint16_t)0xFFFFu : (uint16_t)0u;
      }
      asm volatile("bar.sync 1, 384;" ::: "memory");
      float gacc = 0.f;
      const size_t pitch = (size_t)p.nc * p.lp;
      const int nchs = p.lp >> 3;                      // chunks that exist in the output matrices
      for (int j = u.j; j < u.j_end; ++j) {
        // per-word inputs of the coefficient step (thread wl owns word l = wl); latency hidden by the softmax
        float dotp = 0.f, c2p = 0.f;
        float gsim = 1.f;
        if (!FUSED) {
          if (wl < LPAD) {
            const float* sp = p.stats + ((size_t)j * p.Bc + i) * 2 * LPAD;
            dotp = __ldg(sp + wl);
            c2p = __ldg(sp + LPAD + wl);
          }
          gsim = __ldg(p.dsim + (size_t)j * p.Bc + i);
        } else if (wl < NCOL) {
          accs[wl] = 0.f;                                    // column sums of this pair (ordered by the softmax barriers)
          accs[NCOL + wl] = 0.f;
        }
        // gradient of the word-mean attention (recompute backward only): t1 * dL/dmean[j, i, s] of this thread's rows
        float ms[MAX_NT];
#pragma unroll
        for (int idx = 0; idx < MAX_NT; ++idx) ms[idx] = 0.f;
        if (!FUSED && p.dmean != nullptr) {
          if (wl < NCOL) accs[wl] = 0.f;
          const float* dm = p.dmean + ((size_t)j * p.Bc + i) * p.S;
#pragma unroll
          for (int idx = 0; idx < MAX_NT; ++idx) {
            const int s_glob = tile_ord<FUSED>(idx, NT) * TILE + row;
            if (idx < NT && s_glob < p.S) ms[idx] = p.t1 * __ldg(dm + s_glob);
          }
        }
        float nmb[MAX_NT], inv[MAX_NT];
        // ---------------- phase 1: word softmax, E -> shared memory (S_ stays in TMEM)
#pragma unroll
        for (int idx = 0; idx < MAX_NT; ++idx) {
          if (idx < NT) {
            const int t = tile_ord<FUSED>(idx, NT);
            const int s_glob = t * TILE + row;
            const uint32_t sw = (uint32_t)(s_glob & 7);
            uint8_t* erow = smem + OFF_E + (size_t)(s_glob >> 3) * 1024u + (size_t)sw * 128u;
            const uint32_t rowmask = (s_glob < p.S) ? 0xFFFFFFFFu : 0u;
            STIMED(sw_d1f, mbar_wait(bar(B_D1F + t), n & 1));
            tc_fence_after();
            float x[WG];
            const uint32_t ts = tmem + lane_addr + (uint32_t)(t * LPAD + col0);
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              if (c_lo + c < NCH) {
                tmem_ld8(ts + c * 8, x + c * 8);
              } else {                                   // phantom chunk: never touch (possibly uninitialised) TMEM
#pragma unroll
                for (int k = 0; k < 8; ++k) x[c * 8 + k] = -INFINITY;
              }
            }
            tmem_ld_wait();
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              const float4 b0 = *reinterpret_cast<const float4*>(nM + c * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(nM + c * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                x[c * 8 + k] += bb[k];                     // words beyond the caption: -inf
                m = fmaxf(m, x[c * 8 + k]);
              }
            }
            const bool dead = (m == -INFINITY);                // this group holds no live word
            const float mg = dead ? 0.f : m * LOG2E;
            float sg = 0.f;
#pragma unroll
            for (int k = 0; k < WG; ++k) {
              const float e = ex2(fmaf(x[k], LOG2E, -mg));     // -inf -> 0
              x[k] = e;
              sg += e;
            }
            float2* xb = xch + (xc & 1) * (NGROUP * 128);
            ++xc;
            xb[g * 128 + row] = make_float2(dead ? -INFINITY : mg, sg);
            STIMED(sw_bar, asm volatile("bar.sync 1, 384;" ::: "memory"));
            const float2 v0 = xb[row], v1 = xb[128 + row], v2 = xb[256 + row];
            const float mbv = fmaxf(v0.x, fmaxf(v1.x, v2.x));
            const float tot = v0.y * ex2(v0.x - mbv) + v1.y * ex2(v1.x - mbv) + v2.y * ex2(v2.x - mbv);
            const float rinv = 1.f / tot;
            nmb[idx] = -mbv;
            inv[idx] = rinv;
            const float sc = p.t1_log2e * rinv * ex2(mg - mbv);       // P = e * ex2(mg - mb) / tot
            if (idx == 0) STIMED(sw_ee, mbar_wait(bar(B_EE), (n & 1) ^ 1));   // previous pair's GEMM-T done with E
            // diagonal pair whose attention map is wanted: this row's fp32 numerators go out as they are computed
            float* drow = nullptr;
            if (FUSED && p.diag_raw != nullptr && p.diag_j0 + j == i && s_glob < p.S)
              drow = p.diag_raw + (size_t)i * p.diag_lcap * p.S + s_glob;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              const uint4 cm = *reinterpret_cast<const uint4*>(cM + c * 4);
              const uint32_t mk[4] = {cm.x & rowmask, cm.y & rowmask, cm.z & rowmask, cm.w & rowmask};
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                const float e0 = ex2(x[c * 8 + k] * sc), e1 = ex2(x[c * 8 + k + 1] * sc);
                w[k >> 1] = pack_bf16(e0, e1) & mk[k >> 1];
                if (drow != nullptr) {                       // (lanes = consecutive regions: 128-byte rows per word)
                  const int l = col0 + c * 8 + k;
                  if (l < L && l < p.diag_lcap) drow[(size_t)l * p.S] = e0;
                  if (l + 1 < L && l + 1 < p.diag_lcap) drow[(size_t)(l + 1) * p.S] = e1;
                }
              }
              const int ch = c_lo + c;
              if (ch < NCH)
                *reinterpret_cast<uint4*>(erow + (size_t)(ch >> 3) * e_blk + (size_t)((((uint32_t)ch & 7u) ^ sw) << 4)) =
                    make_uint4(w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar(B_EF));
          }
        }
        if (!FUSED && p.dmean != nullptr) {
          // ---------------- q'_l = sum_s E[s,l] t1 m_s  (A = softmax_s(t1 P): d(t1 P) = A (dA - sum_s A dA), dA = m_s / L)
          float a0[WG];
#pragma unroll
          for (int k = 0; k < WG; ++k) a0[k] = 0.f;
#pragma unroll
          for (int idx = 0; idx < MAX_NT; ++idx) {
            if (idx < NT) {
              const int s_glob = tile_ord<FUSED>(idx, NT) * TILE + row;
              const uint32_t sw = (uint32_t)(s_glob & 7);
              const uint8_t* erow = smem + OFF_E + (size_t)(s_glob >> 3) * 1024u + (size_t)sw * 128u;
#pragma unroll
              for (int c = 0; c < CW; ++c) {
                if (c_lo + c < NCH) {
                  const int ch = c_lo + c;
                  const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                                   (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
                  const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                  for (int k = 0; k < 8; ++k)
                    a0[c * 8 + k] = fmaf((k & 1) ? bf_hi(ew[k >> 1]) : bf_lo(ew[k >> 1]), ms[idx], a0[c * 8 + k]);
                }
              }
            }
          }
          warp_colsum<WG>(a0, lane, accs + col0);      // read after the barriers of the coefficient step
        }
        if (FUSED) {
          // ---------------- pass A: dot'_l = sum_s E S_,  |C'_l|^2 = sum_s E T'  (own columns, all tiles), Z row
          float a1[WG], a2[WG];
#pragma unroll
          for (int k = 0; k < WG; ++k) a1[k] = a2[k] = 0.f;
#pragma unroll
          for (int idx = 0; idx < MAX_NT; ++idx) {
            if (idx < NT) {
              const int t = tile_at(idx, NT);
              STIMED(sw_ttf, mbar_wait(bar(B_TTF), ttc & 1));
              ++ttc;
              tc_fence_after();
              const int s_glob = t * TILE + row;
              const uint32_t sw = (uint32_t)(s_glob & 7);
              const uint8_t* erow = smem + OFF_E + (size_t)(s_glob >> 3) * 1024u + (size_t)sw * 128u;
              const uint32_t ts = tmem + lane_addr + (uint32_t)(t * LPAD + col0);
              const uint32_t tt = tmem + lane_addr + TT_COL + (uint32_t)col0;
              // T' half first, so that its single TMEM buffer goes back to the tensor pipe as early as possible (the next
              // GEMM-T tile then runs under the S_ half below)
#pragma unroll
              for (int c = 0; c < CW; ++c) {
                if (c_lo + c < NCH) {
                  float tv[8];
                  tmem_ld8(tt + c * 8, tv);
                  const int ch = c_lo + c;
                  const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                                   (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
                  tmem_ld_wait();
                  const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
                  if (idx == 0 && q == (zrow >> 5) && lane == (zrow & 31)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) zbuf[col0 + c * 8 + k] = tv[k];     // the ones row: Z_l
                  }
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float e = (k & 1) ? bf_hi(ew[k >> 1]) : bf_lo(ew[k >> 1]);   // 0 in padded rows / words
                    a2[c * 8 + k] = fmaf(e, tv[k], a2[c * 8 + k]);
                  }
                }
              }
              if (idx < NT - 1) {                    // (the last tile's T' is kept: pass B starts with it)
                tc_fence_before();
                mbar_arrive(bar(B_TTE));             // T' buffer may be overwritten by the next GEMM-T tile
              }
#pragma unroll
              for (int c = 0; c < CW; ++c) {
                if (c_lo + c < NCH) {
                  float sv[8];
                  tmem_ld8(ts + c * 8, sv);
                  const int ch = c_lo + c;
                  const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                                   (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
                  tmem_ld_wait();
                  const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float e = (k & 1) ? bf_hi(ew[k >> 1]) : bf_lo(ew[k >> 1]);
                    a1[c * 8 + k] = fmaf(e, sv[k], a1[c * 8 + k]);
                  }
                }
              }
            }
          }
          // column sums over the 32 rows of this warp (halving butterfly), then over warps through shared memory
          warp_colsum<WG>(a1, lane, accs + col0);
          warp_colsum<WG>(a2, lane, accs + NCOL + col0);
          asm volatile("bar.sync 1, 384;" ::: "memory");
          if (wl < NCOL) { dotp = accs[wl]; c2p = accs[NCOL + wl]; }
        }
        // ---------------- phase 2: per region tile, T' arrives; coefficients; dP, u, X / Eo / Bo rows
#pragma unroll
        for (int idx = 0; idx < MAX_NT; ++idx) {
          if (idx < NT) {
            const int t = tile_ord<FUSED>(idx, NT);
            if ((!FUSED || idx > 0) && !lean) {
              STIMED(sw_ttf, mbar_wait(bar(B_TTF), ttc & 1));
              ++ttc;
              tc_fence_after();
            }
            if (idx == 0) {
              // Z_l = sum_s E[s,l] sits in the ones row of this tile; the owning lane quarter reads it
              if (!FUSED && q == (zrow >> 5)) {
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                  if (c_lo + c < NCH) {
                    float z[8];
                    tmem_ld8(tmem + lane_addr + TT_COL + (uint32_t)(col0 + c * 8), z);
                    tmem_ld_wait();
                    if (lane == (zrow & 31)) {
#pragma unroll
                      for (int k = 0; k < 8; ++k) zbuf[col0 + c * 8 + k] = z[k];
                    }
                  }
                }
              }
              asm volatile("bar.sync 1, 384;" ::: "memory");
              // closed-form coefficients of word l = wl (SURVEY.md section 0 / oracle local_sim_pair_bwd)
              const bool live = wl < L;
              const float Zl = live ? zbuf[wl] : 1.f;
              const float iz = 1.f / Zl;
              const float nc = sqrtf(c2p) * iz;
              const float dot = dotp * iz;
              const float prod = nw * nc;
              const float den = fmaxf(prod, p.eps);
              const float cosv = dot / den;
              const float v = live ? p.t2 * cosv : -INFINITY;
              float mx = warp_max(v);
              if (lane == 0) red[warp - 4] = mx;
              asm volatile("bar.sync 1, 384;" ::: "memory");
              mx = red[0];
#pragma unroll
              for (int k = 1; k < 12; ++k) mx = fmaxf(mx, red[k]);
              float ex = live ? __expf(v - mx) : 0.f;
              float tot = warp_sum(ex);
              if (lane == 0) red[12 + (warp - 4)] = tot;
              asm volatile("bar.sync 1, 384;" ::: "memory");
              tot = 0.f;
#pragma unroll
              for (int k = 0; k < 12; ++k) tot += red[12 + k];
              const float dr = gsim * p.t2 * ex / tot;
              const float ddot = dr / den;
              const float dden = (prod >= p.eps) ? -dr * dot / (den * den) : 0.f;
              const float beta = nc > 0.f ? dden * nw / nc : 0.f;
              const float gamma = nw > 0.f ? dden * nc / nw : 0.f;
              const float rs = ddot * dot + beta * nc * nc;
              if (wl < NCOL) {
                // dP = E (a S_ + b T' + c' + w t1 m_s),  X = e E + dS,  Bo = f E;  zero beyond the caption.
                // w = 1 / (Z_l L): weight of E[s,l] in the word-mean attention; its gradient adds w (t1 m_s - t1 q_l)
                const float wm = iz / (float)max(L, 1);
                const float qt = (!FUSED && p.dmean != nullptr) ? accs[wl] * iz : 0.f;      // t1 sum_s A[s,l] m_s
                coefA[wl] = live ? make_float4(p.t1 * ddot * iz, p.t1 * beta * iz * iz, -p.t1 * rs * iz - wm * qt, wm)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                coefX[wl] = live ? ddot * iz : 0.f;
                coefAx[wl] = live ? p.t1 * ddot * iz : 0.f;
              }
              if (p.fo != nullptr && wl < p.lp) p.fo[((size_t)j * p.nc + u.i) * p.lp + wl] = live ? beta * iz * iz : 0.f;
              if (FUSED) {
                if (p.stats_out != nullptr && wl < LPAD) {
                  float* so = p.stats_out + ((size_t)j * p.Bc + i) * 2 * LPAD;
                  so[wl] = dotp;
                  so[LPAD + wl] = c2p;
                }
                if (p.go != nullptr && wl < p.lp) p.go[((size_t)j * p.nc + u.i) * p.lp + wl] = live ? gamma : 0.f;
                if (wl == 0) {      // sim = log sum_l exp(t2 cos_l)  (mean: minus log L), gloria_loss.py:153-158
                  float r = mx + logf(tot);
                  if (p.agg == GLORIA_AGG_MEAN) r -= logf((float)L);
                  p.sim[(size_t)j * p.Bc + i] = r;
                }
              } else if (live) {
                gacc += gamma;
              }
              asm volatile("bar.sync 1, 384;" ::: "memory");
            }
            const int s_glob = t * TILE + row;
            const uint32_t sw = (uint32_t)(s_glob & 7);
            const uint8_t* erow = smem + OFF_E + (size_t)(s_glob >> 3) * 1024u + (size_t)sw * 128u;
            if (lean) {
              // word-mean attention row: sum_l E[s,l] / (Z_l L) over this group's words, then over the 3 groups
              float macc = 0.f;
#pragma unroll
              for (int c = 0; c < CW; ++c) {
                if (c_lo + c < NCH) {
                  const int ch = c_lo + c;
                  const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                                   (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
                  const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                  for (int k = 0; k < 8; ++k)
                    macc = fmaf((k & 1) ? bf_hi(ew[k >> 1]) : bf_lo(ew[k >> 1]), cA[c * 8 + k].w, macc);
                }
              }
              if (idx == 0) {
                tc_fence_before();
                mbar_arrive(bar(B_TTE));               // the T' tile pass A left behind is released
              }
              mbar_arrive(bar(B_D1E + t));             // S_ tile may be overwritten by the next pair's GEMM1
              float2* xb = xch + (xc & 1) * (NGROUP * 128);
              ++xc;
              xb[g * 128 + row].x = macc;
              asm volatile("bar.sync 1, 384;" ::: "memory");
              if (g == 0 && p.mean_out != nullptr && s_glob < p.S)
                p.mean_out[((size_t)j * p.Bc + i) * p.S + s_glob] = xb[row].x + xb[128 + row].x + xb[256 + row].x;
              continue;
            }
            const float nmbv = nmb[idx], rinv = inv[idx];
            const float msv = ms[idx];
            const uint32_t ts = tmem + lane_addr + (uint32_t)(t * LPAD + col0);
            const uint32_t tt = tmem + lane_addr + TT_COL + (uint32_t)col0;
            float dp[WG];
            uint32_t pp[WG / 2];
            float ug = 0.f;
            // pass 1a (T' half): dp = b T' + c'.  The T' buffer goes back to the tensor pipe right after it, so the next
            // tile's GEMM-T runs under pass 1b and pass 2 of this tile.
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              float tv[8];
              if (c_lo + c < NCH) {
                tmem_ld8(tt + c * 8, tv);
              } else {                                         // phantom chunk: finite inputs, zero coefficients
#pragma unroll
                for (int k = 0; k < 8; ++k) tv[k] = 0.f;
              }
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int o = c * 8 + k;
                const float4 f0 = cA[o];
                const float c0 = FUSED ? f0.z : fmaf(f0.w, msv, f0.z);
                dp[o] = fmaf(f0.y, tv[k], c0);
              }
            }
            tc_fence_before();
            mbar_arrive(bar(B_TTE));                 // T' buffer may be overwritten by the next tile's GEMM-T
            // pass 1b (S_ half, one chunk at a time): dP = E (a S_ + dp), partial u = sum_l P dP
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              float sv[8];
              if (c_lo + c < NCH) {
                tmem_ld8(ts + c * 8, sv);
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) sv[k] = 0.f;
              }
              const int ch = min(c_lo + c, NCH - 1);           // phantom chunk: any valid address
              const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                               (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
              const uint4 cm = *reinterpret_cast<const uint4*>(cM + c * 4);
              const float4 a0 = *reinterpret_cast<const float4*>(cAx + c * 8);
              const float4 a1v = *reinterpret_cast<const float4*>(cAx + c * 8 + 4);
              const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1v.x, a1v.y, a1v.z, a1v.w};
              tmem_ld_wait();
              const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
              const uint32_t mk[4] = {cm.x, cm.y, cm.z, cm.w};
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                const int o = c * 8 + k;
                const float e0 = bf_lo(ew[k >> 1]), e1 = bf_hi(ew[k >> 1]);
                const float P0 = ex2(fmaf(sv[k], LOG2E, nmbv)) * rinv, P1 = ex2(fmaf(sv[k + 1], LOG2E, nmbv)) * rinv;
                const float d0 = e0 * fmaf(av[k], sv[k], dp[o]);
                const float d1 = e1 * fmaf(av[k + 1], sv[k + 1], dp[o + 1]);
                const uint32_t pk = pack_bf16(P0, P1) & mk[k >> 1];   // P = 0 beyond the caption (so dS = 0 there)
                dp[o] = d0;
                dp[o + 1] = d1;
                pp[o >> 1] = pk;
                ug = fmaf(P0, d0, ug);                         // fp32 P here: P ~ one-hot makes dP - u a cancellation
                ug = fmaf(P1, d1, ug);                         // (dP is 0 beyond the caption, so no mask is needed)
              }
            }
            tc_fence_before();
            mbar_arrive(bar(B_D1E + t));             // S_ tile may be overwritten by the next pair's GEMM1
            float2* xb = xch + (xc & 1) * (NGROUP * 128);
            ++xc;
            xb[g * 128 + row].x = ug;
            STIMED(sw_bar, asm volatile("bar.sync 1, 384;" ::: "memory"));
            const float uacc = xb[row].x + xb[128 + row].x + xb[256 + row].x;
            // pass 2: rows of X^T  ([(j, s), (i, l)] bf16, 16-byte stores along l)
            const size_t goff = ((size_t)j * p.sp + s_glob) * pitch + (size_t)u.i * p.lp;
            uint4* xo = reinterpret_cast<uint4*>(p.xt + goff);
            const bool row_out = s_glob < p.sp && p.xt != nullptr;   // padded region rows beyond sp do not exist in X^T
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              const int ch = c_lo + c;
              if (ch < nchs && row_out) {
                const uint4 ev = *reinterpret_cast<const uint4*>(erow + (size_t)(ch >> 3) * e_blk +
                                                                 (size_t)((((uint32_t)ch & 7u) ^ sw) << 4));
                const float4 x0 = *reinterpret_cast<const float4*>(cX + c * 8);
                const float4 x1 = *reinterpret_cast<const float4*>(cX + c * 8 + 4);
                const float ce[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
                uint32_t xw[4];
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                  const int o = c * 8 + k;
                  const float e0 = bf_lo(ew[k >> 1]), e1 = bf_hi(ew[k >> 1]);
                  const float P0 = bf_lo(pp[o >> 1]), P1 = bf_hi(pp[o >> 1]);
                  xw[k >> 1] = pack_bf16(fmaf(ce[k], e0, P0 * (dp[o] - uacc)), fmaf(ce[k + 1], e1, P1 * (dp[o + 1] - uacc)));
                }
#ifdef GLORIA_EXP_NOSTORE
                if (xw[0] == 0x12345678u && xw[3] == 0x9abcdef0u) xo[ch] = make_uint4(xw[0], xw[1], xw[2], xw[3]);
#else
                if (stream_x) __stcs(xo + ch, make_uint4(xw[0], xw[1], xw[2], xw[3]));
                else xo[ch] = make_uint4(xw[0], xw[1], xw[2], xw[3]);
#endif
              }
            }
          }
        }
        ++n;
      }
      if (!FUSED && wl < p.lp && gacc != 0.f) atomicAdd(p.gamma + (size_t)i * p.lp + wl, gacc);
    }
#ifdef GLORIA_PHASE_CLOCKS
    if (g_dbg && (lane == 0) && (q == 0)) {
      long long* d = g_dbg + (size_t)blockIdx.x * 32 + 16 + 5 * g;
      d[0] = clock64() - st_all; d[1] = sw_d1f; d[2] = sw_ttf; d[3] = sw_ee; d[4] = sw_bar;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// small SIMT helpers of the backward
// ---------------------------------------------------------------------------------------------------------------
// Gram matrices come out of the GEMM as bf16 [Bi, Spad, Spad]; row S (the first padded region) becomes the ones row
__global__ void gram_ones_row(__nv_bfloat16* __restrict__ G, int S, int Spad) {
  __nv_bfloat16* r = G + ((size_t)blockIdx.x * Spad + S) * Spad;
  for (int x = threadIdx.x; x < Spad; x += blockDim.x) r[x] = __float2bfloat16_rn(x < S ? 1.f : 0.f);
}

constexpr int SCALE_ROWS = 16;
// X[(j,s), (i,l)] *= g[j, i]  in place (fused training path: X was produced for g = 1).
// grid (ceil(R1/8/256), sp/SCALE_ROWS, Bi): each thread owns 8 columns of one image and streams SCALE_ROWS rows
// `consumed` (device flag in the training state, may be null): set once a backward has scaled X in place; a second
// backward over the same state would scale twice, so it poisons its result with NaN instead of returning wrong numbers
__global__ void __launch_bounds__(256) scale_x(__nv_bfloat16* __restrict__ X, const float* __restrict__ g, int R1,
                                               int Spad, int Bc, int i0, int lpad, const int* __restrict__ consumed) {
  const int c8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c8 >= R1) return;
  const int j = blockIdx.z;
  float gs = g[(size_t)j * Bc + i0 + c8 / lpad];
  if (consumed != nullptr && *consumed != 0) gs = __int_as_float(0x7fc00000);
  const size_t row0 = (size_t)j * Spad + (size_t)blockIdx.y * SCALE_ROWS;
#pragma unroll 4
  for (int r = 0; r < SCALE_ROWS; ++r) {
    uint4* ptr = reinterpret_cast<uint4*>(X + (row0 + r) * R1 + c8);
    const uint4 ev = *ptr;
    const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = pack_bf16(bf_lo(ew[k]) * gs, bf_hi(ew[k]) * gs);
    *ptr = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void mark_consumed(int* flag) { *flag = 1; }

// attn[i, l, :] = raw[i, l, :] / sum_s raw[i, l, s] for l < cap_len (rows beyond: zero); one warp per (caption, word)
__global__ void normalise_diag(float* __restrict__ a, const int* __restrict__ cap_lens, int Bc, int lcap, int S) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= Bc * lcap) return;
  const int lane = threadIdx.x & 31;
  const int i = row / lcap, l = row - i * lcap;
  float* r = a + (size_t)row * S;
  if (l >= min(max(cap_lens[i], 0), lcap)) {
    for (int s = lane; s < S; s += 32) r[s] = 0.f;
    return;
  }
  float z = 0.f;
  for (int s = lane; s < S; s += 32) z += r[s];
  z = 1.f / warp_sum(z);
  for (int s = lane; s < S; s += 32) r[s] *= z;
}

// gamma[i, l] = sum_j g[j, i] * go[j, (i,l)]   (fused training path).  Block = 32 columns x 8 slices of the image axis
// (one 128-byte row segment per warp-load, 8 loads in flight per thread), slices summed in a fixed order through shared
// memory: deterministic, and R1 / 32 blocks instead of R1 / 256 (a caption shard of 64 captions used 26 SMs for 140 us).
__global__ void __launch_bounds__(256) gamma_sum(const float* __restrict__ go, const float* __restrict__ g,
                                                 float* __restrict__ gamma, int Bi, int R1, int Bc, int i0, int lpad) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < R1) {
    const int i = i0 + c / lpad;
    const int per = (Bi + 7) / 8;
    const int j0 = slice * per, j1 = min(Bi, j0 + per);
    const float* gp = g + i;
    const float* op = go + c;
    int j = j0;
    for (; j + 8 <= j1; j += 8) {
      float a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] = __ldg(gp + (size_t)(j + k) * Bc);
        b[k] = __ldg(op + (size_t)(j + k) * R1);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(a[k], b[k], acc);
    }
    for (; j < j1; ++j) acc = fmaf(__ldg(gp + (size_t)j * Bc), __ldg(op + (size_t)j * R1), acc);
  }
  part[slice][lane] = acc;
  __syncthreads();
  if (slice == 0 && c < R1) {
    float tot = part[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) tot += part[k][lane];
    gamma[(size_t)i0 * lpad + c] = tot;
  }
}

__global__ void f32_to_bf16(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (size_t k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
  }
}

// dRt [Bi, Spad, D] -> d_ctx [Bi, D, S];  grid (ceil(S/32), D/32, Bi), block (32, 8)
__global__ void unpack_dctx(const float* __restrict__ dRt, float* __restrict__ dctx, int D, int S, int Spad) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int s = s0 + r, d = d0 + threadIdx.x;
    t[r][threadIdx.x] = (s < S) ? dRt[((size_t)b * Spad + s) * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int d = d0 + r, s = s0 + threadIdx.x;
    if (s < S) dctx[((size_t)b * D + d) * S + s] = t[threadIdx.x][r];
  }
}

// dWt [Bc, LPAD, D] (+ gamma * W) -> d_words [Bc, D, Lw], zero outside [off, off + cap_len)
// grid (ceil(Lw/32), D/32, Bc), block (32, 8)
__global__ void unpack_dwords_tc(const float* __restrict__ dWt, const float* __restrict__ gamma,
                                 const __nv_bfloat16* __restrict__ Wt, const int* __restrict__ cap_lens,
                                 float* __restrict__ dwords, int D, int Lw, int lpad, int lcap, int off) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int L = min(max(cap_lens[b], 0), lcap);
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int lw = l0 + r, d = d0 + threadIdx.x;      // lw: position on the caller's word axis
    const int l = lw - off;
    float v = 0.f;
    if (lw < Lw && l >= 0 && l < L) {
      const size_t o = ((size_t)b * lpad + l) * D + d;
      v = dWt[o] + gamma[(size_t)b * lpad + l] * __bfloat162float(Wt[o]);
    }
    t[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int d = d0 + r, lw = l0 + threadIdx.x;
    if (lw < Lw) dwords[((size_t)b * D + d) * Lw + lw] = t[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------
struct Plan {
  int nc;   // captions per chunk
  size_t off_gram, off_dwt, off_drt, off_m, off_mb, off_gamma, off_stats, off_sim, off_cublas, off_x, off_e, off_f, total;
};
constexpr size_t CUBLAS_WS = 64u << 20;

// lp = column pitch of the operand matrices (round_up(Lcap, 8)), lpad = the kernels' word padding (stats pitch)
// sp = region rows per image of the operand matrices (round_up(S, 16)), Spad = the kernels' tile padding (Gram buffer)
size_t fixed_bytes(int Bi, int Bc, int D, int Spad, int sp, int lp, int lpad, bool own_stats, Plan* pl) {
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 1024); return r; };
  Plan t{};
  t.off_gram = take((size_t)Bi * Spad * Spad * 2);
  t.off_dwt = take((size_t)Bc * lp * D * 4);
  t.off_drt = take((size_t)Bi * sp * D * 4);
  t.off_m = take((size_t)Bi * sp * sp * 4);
  t.off_mb = take((size_t)Bi * sp * sp * 2);
  t.off_gamma = take((size_t)Bc * lp * 4);
  t.off_stats = take(own_stats ? (size_t)Bi * Bc * 2 * lpad * 4 : 0);
  t.off_sim = take(own_stats ? (size_t)Bi * Bc * 4 : 0);
  t.off_cublas = take(CUBLAS_WS);
  if (pl) *pl = t;
  return o;
}
size_t per_caption_bytes(int Bi, int sp, int lp) {
  return 2 * align_up((size_t)Bi * sp * lp * 2, 1024) + align_up((size_t)Bi * lp * 4, 1024) + 4096;
}

Plan make_plan(int Bi, int Bc, int D, int Spad, int sp, int lp, int lpad, bool own_stats, size_t bytes) {
  Plan pl{};
  const size_t fixed = fixed_bytes(Bi, Bc, D, Spad, sp, lp, lpad, own_stats, &pl);
  const size_t per = per_caption_bytes(Bi, sp, lp);
  if (bytes < fixed + per) { pl.nc = 0; return pl; }
  size_t nc = (bytes - fixed) / per;
  if (nc > (size_t)Bc) nc = Bc;
  pl.nc = (int)nc;
  size_t o = fixed;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 1024); return r; };
  const size_t arr = (size_t)Bi * sp * nc * lp * 2;
  pl.off_x = take(arr);
  pl.off_e = take(arr);
  pl.off_f = take((size_t)Bi * nc * lp * 4);
  pl.total = o;
  if (pl.total > bytes) pl.nc = 0;
  return pl;
}

cublasHandle_t cublas_handle() { return (cublasHandle_t)cublas_handle_opaque(); }

// L2 eviction hints of the fused training kernel (see PairParams::l2_hints); GLORIA_B200_L2_HINTS overrides (0..3).
// Measured at B = 512 on one box: none 64.7 ms, stores evict-first 62.9, loads evict-last 64.9, both 63.2 -> default 1.
int l2_hints_mode() {
  static const int mode = [] {
    const char* e = getenv("GLORIA_B200_L2_HINTS");
    return e ? atoi(e) & 3 : 1;
  }();
  return mode;
}

// How the two accumulation GEMMs of the training backward run (GLORIA_B200_BWD_GEMM):
//   0 "inflight": own CTA-pair tcgen05 GEMMs (tc_gemm.cu) that apply g = dsim[j,i] to the A operand on the fly
//   1 "scale"   : one streaming pass X *= g, then the own GEMMs in their plain mode
//   2 "cublas"  : the streaming pass, then cuBLAS -- the default: a plain library GEMM that runs the same tile at the same
//                 tensor-pipe occupancy but ~25% fewer L2 sector reads and therefore higher clocks under the 1000 W cap
//                 (measured: DESIGN.md section 5b); the own kernels stay under test and one environment variable away
int bwd_gemm_mode() {
  static const int mode = [] {
    const char* e = getenv("GLORIA_B200_BWD_GEMM");
    if (!e) return 2;
    if (!strcmp(e, "inflight") || !strcmp(e, "0")) return 0;
    if (!strcmp(e, "scale") || !strcmp(e, "1")) return 1;
    return 2;
  }();
  return mode;
}

// GLORIA_B200_BWD_OVERLAP=1: the M-term kernel (reads E^T, never X) runs on a library-owned side stream beside the
// streaming pass X *= g and the dR GEMM instead of between the GEMMs.  Per device and thread: one non-blocking stream and
// two events (fork / join), created on first use; the fork-join is expressed with events only, so it is capturable.
int bwd_overlap_mode() {
  static const int mode = [] {
    const char* e = getenv("GLORIA_B200_BWD_OVERLAP");
    return e ? atoi(e) : 0;
  }();
  return mode;
}
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
SideStream* side_stream() {
  static thread_local SideStream s[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!s[dev].st) {
    if (cudaStreamCreateWithFlags(&s[dev].st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s[dev].fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s[dev].join, cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
  }
  return &s[dev];
}

#define GLORIA_CUBLAS(expr)                                                                       \
  do {                                                                                            \
    cublasStatus_t _s = (expr);                                                                   \
    if (_s != CUBLAS_STATUS_SUCCESS) return fail(GLORIA_ERR_DRIVER, "%s -> cublas status %d", #expr, (int)_s); \
    ++launch_counter();                                                                           \
  } while (0)

template <int LPAD, bool FUSED>
int launch_pair(const CUtensorMap& rt, const CUtensorMap& wt, const CUtensorMap& g, const CUtensorMap& e,
                const PairParams& p, int grid, cudaStream_t st) {
  GLORIA_CUDA(cudaFuncSetAttribute(tc_bwd_pair_kernel<LPAD, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SMEM_BYTES));
  const int slot = FUSED ? GLORIA_TIMER_TC_FWD : GLORIA_TIMER_TC_BWD_PAIR;
  if (p.timer_first) timer_record(slot, 0, st);          // a forward launched in image parts is timed first to last
  tc_bwd_pair_kernel<LPAD, FUSED><<<grid, NTHREADS, SMEM_BYTES, st>>>(rt, wt, g, e, p);
  if (p.timer_last) timer_record(slot, 1, st);
  GLORIA_LAUNCHED("tc_bwd_pair_kernel");
  return GLORIA_OK;
}

template <bool FUSED>
int launch_pair_lpad(int lpad, const CUtensorMap& rt, const CUtensorMap& wt, const CUtensorMap& g, const CUtensorMap& e,
                     const PairParams& p, int grid, cudaStream_t st) {
  switch (lpad) {
    case 16: return launch_pair<16, FUSED>(rt, wt, g, e, p, grid, st);
    case 32: return launch_pair<32, FUSED>(rt, wt, g, e, p, grid, st);
    case 48: return launch_pair<48, FUSED>(rt, wt, g, e, p, grid, st);
    case 64: return launch_pair<64, FUSED>(rt, wt, g, e, p, grid, st);
    case 80: return launch_pair<80, FUSED>(rt, wt, g, e, p, grid, st);
    case 96: return launch_pair<96, FUSED>(rt, wt, g, e, p, grid, st);
    case 112: return launch_pair<112, FUSED>(rt, wt, g, e, p, grid, st);
    case 128: return launch_pair<128, FUSED>(rt, wt, g, e, p, grid, st);
  }
  return fail(GLORIA_ERR_UNSUPPORTED, "lpad %d", lpad);
}

// ---- fused training path: one workspace shared by the forward (gram, X, E, fo, go) and the backward (the rest)
struct TrainPlan {
  size_t off_gram, off_x, off_e, off_fo, off_go, off_dwt, off_drt, off_m, off_mb, off_gamma, off_cublas, off_flag, total;
};
TrainPlan train_plan(int Bi, int Bc, int D, int Spad, int sp, int lpad) {
  TrainPlan t{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 1024); return r; };
  const size_t arr = (size_t)Bi * sp * Bc * lpad * 2;
  t.off_gram = take((size_t)Bi * Spad * Spad * 2);
  t.off_x = take(arr);
  t.off_e = take(arr);
  t.off_fo = take((size_t)Bi * Bc * lpad * 4);
  t.off_go = take((size_t)Bi * Bc * lpad * 4);
  t.off_dwt = take((size_t)Bc * lpad * D * 4);
  t.off_drt = take((size_t)Bi * sp * D * 4);
  t.off_m = take((size_t)Bi * sp * sp * 4);
  t.off_mb = take((size_t)Bi * sp * sp * 2);
  t.off_gamma = take((size_t)Bc * lpad * 4);
  t.off_cublas = take(CUBLAS_WS);
  t.off_flag = take(1024);
  t.total = o;
  return t;
}

int gram_matrices(cublasHandle_t h, const __nv_bfloat16* Rt, __nv_bfloat16* gram, int Bi, int D, int S, int Spad,
                  int sp, cudaStream_t st) {
  const float one = 1.f, zero = 0.f;
  // G_j = Rt_j Rt_j^T (Rt_j row-major [sp, D] == column-major [D, sp]) into the [Spad, Spad] tile buffer (rows / columns
  // >= sp stay zero); row S becomes the row of ones
  GLORIA_CUDA(cudaMemsetAsync(gram, 0, (size_t)Bi * Spad * Spad * 2, st));
  GLORIA_CUBLAS(cublasGemmStridedBatchedEx(h, CUBLAS_OP_T, CUBLAS_OP_N, sp, sp, D, &one, Rt, CUDA_R_16BF, D,
                                           (long long)sp * D, Rt, CUDA_R_16BF, D, (long long)sp * D, &zero, gram,
                                           CUDA_R_16BF, Spad, (long long)Spad * Spad, Bi, CUBLAS_COMPUTE_32F,
                                           CUBLAS_GEMM_DEFAULT));
  gram_ones_row<<<Bi, 128, 0, st>>>(gram, S, Spad);
  GLORIA_LAUNCHED("gram_ones_row");
  return GLORIA_OK;
}

// the accumulation GEMMs shared by both backward flavours (X, E hold nc captions starting at i0; f, g: see launch_mterm)
int accumulate_chunk(cublasHandle_t h, const __nv_bfloat16* Rt, const __nv_bfloat16* Wt, const __nv_bfloat16* X,
                     const __nv_bfloat16* E, const float* f, const float* g, float* dWt, float* dRt, float* Mf, int Bi,
                     int Bc, int D, int sp, int lp, int i0, int nc, bool first, cudaStream_t st) {
  const int K1 = Bi * sp, R1 = nc * lp;
  int rc;
  if (bwd_gemm_mode() != 2) {
    // dWt[(i,l), d] = sum_(j,s) X^T[(j,s),(i,l)] Rt[(j,s), d]      (A^T = X^T in memory)
    if ((rc = acc_gemm(X, Rt, dWt + (size_t)i0 * lp * D, R1, D, K1, D, false, 1, false, nullptr, 0, 0, 1, 1, false, st))) return rc;
    // dRt[(j,s), d] (+)= sum_(i,l) X^T[(j,s),(i,l)] Wt[(i,l), d]
    if ((rc = acc_gemm(X, Wt + (size_t)i0 * lp * D, dRt, K1, D, R1, D, true, 1, !first, nullptr, 0, 0, 1, 1, false, st))) return rc;
  } else {
    const float one = 1.f, zero = 0.f;
    const float beta = first ? 0.f : 1.f;
    GLORIA_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, D, R1, K1, &one, Rt, CUDA_R_16BF, D, X, CUDA_R_16BF, R1, &zero,
                               dWt + (size_t)i0 * lp * D, CUDA_R_32F, D, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
    GLORIA_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, D, K1, R1, &one, Wt + (size_t)i0 * lp * D, CUDA_R_16BF, D, X,
                               CUDA_R_16BF, R1, &beta, dRt, CUDA_R_32F, D, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
  }
  // M_j[a, b] (+)= sum_(i,l) E^T[(j,a),(i,l)] g[j,i] f[j,(i,l)] E^T[(j,b),(i,l)]   (own tcgen05 kernel, tc_mterm.cu)
  return launch_mterm(E, f, g, Mf, nullptr, Bi, Bc, i0, R1, lp, sp, !first, st);
}

int finish_backward(cublasHandle_t h, const __nv_bfloat16* Rt, const __nv_bfloat16* Wt, const int32_t* cap_lens,
                    float* dWt, float* dRt, float* Mf, __nv_bfloat16* Mb, const float* gamma, float* d_ctx,
                    float* d_words, int Bi, int Bc, int D, int S, int Spad, int Lw, int lpad, int Lcap, int word_off,
                    cudaStream_t st) {
  const float one = 1.f;
  // dRt_j += M_j Rt_j   (M_j symmetric up to rounding; column-major: [D, Spad] = Rt_j^T . M_j)
  const size_t nm = (size_t)Bi * Spad * Spad;
  f32_to_bf16<<<(unsigned)((nm / 4 + 255) / 256), 256, 0, st>>>(Mf, Mb, nm);
  GLORIA_LAUNCHED("f32_to_bf16");
  GLORIA_CUBLAS(cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_N, D, Spad, Spad, &one, Rt, CUDA_R_16BF, D,
                                           (long long)Spad * D, Mb, CUDA_R_16BF, Spad, (long long)Spad * Spad, &one, dRt,
                                           CUDA_R_32F, D, (long long)Spad * D, Bi, CUBLAS_COMPUTE_32F,
                                           CUBLAS_GEMM_DEFAULT));
  timer_record(GLORIA_TIMER_TC_BWD_GEMM, 1, st);
  unpack_dctx<<<dim3((S + 31) / 32, D / 32, Bi), dim3(32, 8), 0, st>>>(dRt, d_ctx, D, S, Spad);
  GLORIA_LAUNCHED("unpack_dctx");
  unpack_dwords_tc<<<dim3((Lw + 31) / 32, D / 32, Bc), dim3(32, 8), 0, st>>>(dWt, gamma, Wt, cap_lens, d_words, D, Lw,
                                                                           lpad, Lcap, word_off);
  GLORIA_LAUNCHED("unpack_dwords_tc");
  return GLORIA_OK;
}

}  // namespace bw
}  // namespace tc
}  // namespace gloria

using namespace gloria;
using namespace gloria::tc;

extern "C" size_t gloria_b200_tc_bwd_workspace(int Bi, int Bc, int D, int S, int Lcap, int have_stats, size_t budget) {
  if (Bi <= 0 || Bc <= 0 || gloria_b200_tc_supported(D, S, Lcap)) return 0;
  const int Spad = gloria_b200_tc_spad(S), lpad = gloria_b200_tc_lpad(Lcap), lp = gloria_b200_tc_lp(Lcap);
  const int sp = gloria_b200_tc_sp(S);
  const size_t fixed = bw::fixed_bytes(Bi, Bc, D, Spad, sp, lp, lpad, !have_stats, nullptr);
  const size_t per = bw::per_caption_bytes(Bi, sp, lp);
  size_t want = fixed + per * (size_t)Bc;
  if (budget != 0 && want > budget) {
    size_t nc = budget > fixed + per ? (budget - fixed) / per : 1;
    if (nc < 1) nc = 1;
    want = fixed + per * nc;
  }
  return want;
}

extern "C" int gloria_b200_tc_local_sim_bwd(const void* ctx_h, const void* ctx_t, const void* ctx_n,
                                            const void* words_h, const void* words_t, const float* wnorm,
                                            const int32_t* cap_lens, const float* stats, int Bi,
                                            int Bc, int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                            float temp2, int agg, float eps, const float* dsim,
                                            const float* d_attn_mean, float* d_ctx, float* d_words, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  GLORIA_CHECK_ARG(ctx_h && ctx_t && ctx_n && words_h && words_t && wnorm && cap_lens && dsim && d_ctx && d_words &&
                       workspace,
                   "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && word_off >= 0 && word_off + Lcap <= Lw, "bad sizes");
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  if (agg == GLORIA_AGG_MAX) return fail(GLORIA_ERR_UNSUPPORTED, "backward of agg=max is not part of the path");
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S), lpad = gloria_b200_tc_lpad(Lcap), lp = gloria_b200_tc_lp(Lcap);
  const int sp = gloria_b200_tc_sp(S);
  const bw::Plan pl = bw::make_plan(Bi, Bc, D, Spad, sp, lp, lpad, stats == nullptr, workspace_bytes);
  if (pl.nc < 1) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B too small", workspace_bytes);
  char* ws = (char*)workspace;
  __nv_bfloat16* gram = (__nv_bfloat16*)(ws + pl.off_gram);
  float* dWt = (float*)(ws + pl.off_dwt);
  float* dRt = (float*)(ws + pl.off_drt);
  float* Mf = (float*)(ws + pl.off_m);
  __nv_bfloat16* Mb = (__nv_bfloat16*)(ws + pl.off_mb);
  float* gamma = (float*)(ws + pl.off_gamma);
  __nv_bfloat16* X = (__nv_bfloat16*)(ws + pl.off_x);
  __nv_bfloat16* E = (__nv_bfloat16*)(ws + pl.off_e);
  float* Fo = (float*)(ws + pl.off_f);
  int rc;
  if (stats == nullptr) {   // stand-alone use: one forward pass regenerates the per-word statistics
    float* own = (float*)(ws + pl.off_stats);
    if ((rc = gloria_b200_tc_local_sim_fwd(ctx_h, ctx_n, words_h, wnorm, cap_lens, Bi, Bc, D, S, Lcap, temp1, temp2,
                                           agg, eps, (float*)(ws + pl.off_sim), own, stream)))
      return rc;
    stats = own;
  }
  cublasHandle_t h = bw::cublas_handle();
  if (!h) return fail(GLORIA_ERR_DRIVER, "cublasCreate failed");
  GLORIA_CUBLAS(cublasSetStream(h, st));
  GLORIA_CUBLAS(cublasSetWorkspace(h, ws + pl.off_cublas, bw::CUBLAS_WS));
  const __nv_bfloat16* Rt = (const __nv_bfloat16*)ctx_t;
  const __nv_bfloat16* Wt = (const __nv_bfloat16*)words_t;
  if ((rc = bw::gram_matrices(h, Rt, gram, Bi, D, S, Spad, sp, st))) return rc;
  GLORIA_CUDA(cudaMemsetAsync(gamma, 0, (size_t)Bc * lp * sizeof(float), st));

  CUtensorMap rt, wt, gm;
  if ((rc = make_map(&rt, ctx_h, (uint64_t)D, (uint64_t)Bi * Spad, TILE))) return rc;
  if ((rc = make_map(&wt, words_h, (uint64_t)D, (uint64_t)Bc * lpad, (uint32_t)lpad))) return rc;
  if ((rc = make_map(&gm, gram, (uint64_t)Spad, (uint64_t)Bi * Spad, TILE))) return rc;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  for (int i0 = 0; i0 < Bc; i0 += pl.nc) {
    const int nc = min(pl.nc, Bc - i0);
    const int R1 = nc * lp;
    bw::PairParams p{};
    p.wnorm = wnorm; p.cap_lens = cap_lens; p.stats = stats; p.dsim = dsim; p.dmean = d_attn_mean;
    p.xt = X; p.et = E; p.fo = Fo; p.gamma = gamma;
    CUtensorMap em;
    if ((rc = make_map4(&em, E, (uint64_t)lp, (uint64_t)nc, (uint64_t)sp, (uint64_t)Bi, (uint64_t)lp, (uint64_t)R1,
                        (uint64_t)sp * R1, TILE)))
      return rc;
    p.Bi = Bi; p.Bc = Bc; p.i0 = i0; p.nc = nc; p.D = D; p.S = S; p.NT = Spad / TILE; p.lp = lp; p.sp = sp;
    p.t1 = temp1; p.t1_log2e = temp1 * 1.4426950408889634f; p.t2 = temp2; p.eps = eps; p.agg = agg;
    p.dbg = (long long*)g_phase_clock_buffer;
    if ((rc = bw::launch_pair_lpad<false>(lpad, rt, wt, gm, em, p, sms, st))) return rc;
    timer_record(GLORIA_TIMER_TC_BWD_GEMM, 0, st);
    if ((rc = bw::accumulate_chunk(h, Rt, Wt, X, E, Fo, nullptr, dWt, dRt, Mf, Bi, Bc, D, sp, lp, i0, nc, i0 == 0, st)))
      return rc;
    if (i0 + nc < Bc) timer_record(GLORIA_TIMER_TC_BWD_GEMM, 1, st);
  }
  return bw::finish_backward(h, Rt, Wt, cap_lens, dWt, dRt, Mf, Mb, gamma, d_ctx, d_words, Bi, Bc, D, S, sp, Lw, lp,
                             Lcap, word_off, st);
}

// ---------------------------------------------------------------------------------------------------------------
// fused training path
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t gloria_b200_tc_train_workspace(int Bi, int Bc, int D, int S, int Lcap) {
  if (Bi <= 0 || Bc <= 0 || gloria_b200_tc_supported(D, S, Lcap)) return 0;
  return bw::train_plan(Bi, Bc, D, gloria_b200_tc_spad(S), gloria_b200_tc_sp(S), gloria_b200_tc_lp(Lcap)).total;
}

// Image range [j0, j0 + nj) of a (Bi x Bc) training forward: the workspace is laid out for all Bi images, this call
// fills the rows of the range (X^T / E^T rows, f, gamma, Gram matrices, sim rows).  range_h / range_t point at the packed
// copies of the FIRST image of the range (they need not live in the arrays of the other images: a caption-sharded
// caller runs its own images from a private buffer while the gather of the others is in flight); sim is the base of
// the full [Bi, Bc] matrix.  flags: 1 = record the bench timer's start before, 2 = its stop after the launch,
// 4 = this is the first launch of a forward (resets the state's consumed flag).
extern "C" int gloria_b200_tc_local_sim_fwd_train_range(const void* range_h, const void* range_t, const void* words_h,
                                                        const float* wnorm, const int32_t* cap_lens, int Bi, int j0,
                                                        int nj, int Bc, int D, int S, int Lcap, float temp1, float temp2,
                                                        int agg, float eps, float* sim, void* workspace,
                                                        size_t workspace_bytes, int flags, float* attn_diag_raw,
                                                        int diag_lcap, void* stream) {
  GLORIA_CHECK_ARG(range_h && range_t && words_h && wnorm && cap_lens && sim && workspace, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0, "bad batch sizes %d x %d", Bi, Bc);
  GLORIA_CHECK_ARG(j0 >= 0 && nj > 0 && j0 + nj <= Bi, "bad image range [%d, %d) of %d", j0, j0 + nj, Bi);
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  if (agg == GLORIA_AGG_MAX) return fail(GLORIA_ERR_UNSUPPORTED, "agg=max has no backward: use the plain forward");
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S), lpad = gloria_b200_tc_lpad(Lcap), lp = gloria_b200_tc_lp(Lcap);
  const int sp = gloria_b200_tc_sp(S);
  const bw::TrainPlan pl = bw::train_plan(Bi, Bc, D, Spad, sp, lp);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  char* ws = (char*)workspace;
  const int R1 = Bc * lp;
  // everything below is addressed relative to image j0
  const __half* rh = (const __half*)range_h;
  const __nv_bfloat16* rt_ = (const __nv_bfloat16*)range_t;
  __nv_bfloat16* gram = (__nv_bfloat16*)(ws + pl.off_gram) + (size_t)j0 * Spad * Spad;
  __nv_bfloat16* xt = (__nv_bfloat16*)(ws + pl.off_x) + (size_t)j0 * sp * R1;
  __nv_bfloat16* et = (__nv_bfloat16*)(ws + pl.off_e) + (size_t)j0 * sp * R1;
  cublasHandle_t h = bw::cublas_handle();
  if (!h) return fail(GLORIA_ERR_DRIVER, "cublasCreate failed");
  GLORIA_CUBLAS(cublasSetStream(h, st));
  GLORIA_CUBLAS(cublasSetWorkspace(h, ws + pl.off_cublas, bw::CUBLAS_WS));
  int rc;
  if (flags & 4) GLORIA_CUDA(cudaMemsetAsync(ws + pl.off_flag, 0, 1024, st));       // fresh state: not consumed yet
  if ((rc = bw::gram_matrices(h, rt_, gram, nj, D, S, Spad, sp, st))) return rc;
  CUtensorMap rt, wt, gm, em;
  if ((rc = make_map(&rt, rh, (uint64_t)D, (uint64_t)nj * Spad, TILE))) return rc;
  if ((rc = make_map(&wt, words_h, (uint64_t)D, (uint64_t)Bc * lpad, (uint32_t)lpad))) return rc;
  if ((rc = make_map(&gm, gram, (uint64_t)Spad, (uint64_t)nj * Spad, TILE))) return rc;
  if ((rc = make_map4(&em, et, (uint64_t)lp, (uint64_t)Bc, (uint64_t)sp, (uint64_t)nj, (uint64_t)lp, (uint64_t)R1,
                      (uint64_t)sp * R1, TILE)))
    return rc;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bw::PairParams p{};
  p.wnorm = wnorm; p.cap_lens = cap_lens; p.stats = nullptr; p.dsim = nullptr;
  p.xt = xt; p.et = et;
  p.fo = (float*)(ws + pl.off_fo) + (size_t)j0 * R1; p.go = (float*)(ws + pl.off_go) + (size_t)j0 * R1;
  p.gamma = nullptr; p.sim = sim + (size_t)j0 * Bc;
  p.Bi = nj; p.Bc = Bc; p.i0 = 0; p.nc = Bc; p.D = D; p.S = S; p.NT = Spad / TILE; p.lp = lp; p.sp = sp;
  p.t1 = temp1; p.t1_log2e = temp1 * 1.4426950408889634f; p.t2 = temp2; p.eps = eps; p.agg = agg;
  p.dbg = (long long*)g_phase_clock_buffer;
  p.timer_first = (flags & 1) != 0; p.timer_last = (flags & 2) != 0;
  p.l2_hints = bw::l2_hints_mode();
  p.diag_raw = attn_diag_raw; p.diag_lcap = diag_lcap; p.diag_j0 = j0;      // pair (j0 + j, i) is diagonal when equal
  return bw::launch_pair_lpad<true>(lpad, rt, wt, gm, em, p, sms, st);
}

// Same with ctx_h / ctx_t given as the bases of the full arrays; the timer brackets the first to the last range.
extern "C" int gloria_b200_tc_local_sim_fwd_train_part(const void* ctx_h, const void* ctx_t, const void* words_h,
                                                       const float* wnorm, const int32_t* cap_lens, int Bi, int j0,
                                                       int nj, int Bc, int D, int S, int Lcap, float temp1, float temp2,
                                                       int agg, float eps, float* sim, void* workspace,
                                                       size_t workspace_bytes, void* stream) {
  GLORIA_CHECK_ARG(ctx_h && ctx_t, "null pointer");
  GLORIA_CHECK_ARG(j0 >= 0 && nj > 0 && j0 + nj <= Bi, "bad image range [%d, %d) of %d", j0, j0 + nj, Bi);
  const int Spad = gloria_b200_tc_spad(S), sp = gloria_b200_tc_sp(S);
  const int flags = (j0 == 0 ? 1 | 4 : 0) | (j0 + nj == Bi ? 2 : 0);
  return gloria_b200_tc_local_sim_fwd_train_range((const __half*)ctx_h + (size_t)j0 * Spad * D,
                                                  (const __nv_bfloat16*)ctx_t + (size_t)j0 * sp * D, words_h, wnorm, cap_lens,
                                                  Bi, j0, nj, Bc, D, S, Lcap, temp1, temp2, agg, eps, sim, workspace,
                                                  workspace_bytes, flags, nullptr, 0, stream);
}

extern "C" int gloria_b200_tc_local_sim_fwd_train(const void* ctx_h, const void* ctx_t, const void* words_h,
                                                  const float* wnorm, const int32_t* cap_lens, int Bi, int Bc, int D,
                                                  int S, int Lcap, float temp1, float temp2, int agg, float eps,
                                                  float* sim, void* workspace, size_t workspace_bytes, void* stream) {
  return gloria_b200_tc_local_sim_fwd_train_part(ctx_h, ctx_t, words_h, wnorm, cap_lens, Bi, 0, Bi, Bc, D, S, Lcap, temp1,
                                                 temp2, agg, eps, sim, workspace, workspace_bytes, stream);
}

// Training forward that also returns the attention maps of the diagonal pairs (Bi == Bc): attn_diag [Bc, diag_lcap, S],
// rows beyond each caption's length zero.  The maps come out of the fused kernel's own softmax (fp32 numerators stored by
// the diagonal pairs, one normalising launch after it) instead of a second, B-pair pass over the features.
extern "C" int gloria_b200_tc_local_sim_fwd_train_diag(const void* ctx_h, const void* ctx_t, const void* words_h,
                                                       const float* wnorm, const int32_t* cap_lens, int Bi, int Bc,
                                                       int D, int S, int Lcap, float temp1, float temp2, int agg,
                                                       float eps, float* sim, void* workspace, size_t workspace_bytes,
                                                       float* attn_diag, int diag_lcap, void* stream) {
  GLORIA_CHECK_ARG(attn_diag != nullptr && diag_lcap >= Lcap, "attn_diag buffer missing or shorter than Lcap");
  GLORIA_CHECK_ARG(Bi == Bc, "diagonal attention maps need as many images as captions, got %d x %d", Bi, Bc);
  int rc = gloria_b200_tc_local_sim_fwd_train_range(ctx_h, ctx_t, words_h, wnorm, cap_lens, Bi, 0, Bi, Bc, D, S, Lcap, temp1,
                                                    temp2, agg, eps, sim, workspace, workspace_bytes, 7, attn_diag,
                                                    diag_lcap, stream);
  if (rc) return rc;
  const int rows = Bc * diag_lcap;
  bw::normalise_diag<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(attn_diag, cap_lens, Bc, diag_lcap, S);
  GLORIA_LAUNCHED("normalise_diag");
  return GLORIA_OK;
}

// "lean" forward: sim + word-mean attention (+ the per-word statistics a later recompute backward needs)
extern "C" size_t gloria_b200_tc_mean_workspace(int Bi, int Bc, int D, int S, int Lcap) {
  if (Bi <= 0 || Bc <= 0 || gloria_b200_tc_supported(D, S, Lcap)) return 0;
  const int Spad = gloria_b200_tc_spad(S);
  return align_up((size_t)Bi * Spad * Spad * 2, 1024) + bw::CUBLAS_WS;
}

extern "C" int gloria_b200_tc_local_sim_fwd_mean(const void* ctx_h, const void* ctx_t, const void* words_h,
                                                 const float* wnorm, const int32_t* cap_lens, int Bi, int Bc, int D,
                                                 int S, int Lcap, float temp1, float temp2, int agg, float eps,
                                                 float* sim, float* attn_mean, float* stats, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  GLORIA_CHECK_ARG(ctx_h && ctx_t && words_h && wnorm && cap_lens && sim && attn_mean && workspace, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0, "bad batch sizes %d x %d", Bi, Bc);
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  if (agg == GLORIA_AGG_MAX) return fail(GLORIA_ERR_UNSUPPORTED, "agg=max: use the plain forward");
  const size_t need = gloria_b200_tc_mean_workspace(Bi, Bc, D, S, Lcap);
  if (workspace_bytes < need) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S), lpad = gloria_b200_tc_lpad(Lcap), lp = gloria_b200_tc_lp(Lcap);
  const int sp = gloria_b200_tc_sp(S);
  char* ws = (char*)workspace;
  __nv_bfloat16* gram = (__nv_bfloat16*)ws;
  cublasHandle_t h = bw::cublas_handle();
  if (!h) return fail(GLORIA_ERR_DRIVER, "cublasCreate failed");
  GLORIA_CUBLAS(cublasSetStream(h, st));
  GLORIA_CUBLAS(cublasSetWorkspace(h, ws + align_up((size_t)Bi * Spad * Spad * 2, 1024), bw::CUBLAS_WS));
  int rc;
  if ((rc = bw::gram_matrices(h, (const __nv_bfloat16*)ctx_t, gram, Bi, D, S, Spad, sp, st))) return rc;
  CUtensorMap rt, wt, gm;
  if ((rc = make_map(&rt, ctx_h, (uint64_t)D, (uint64_t)Bi * Spad, TILE))) return rc;
  if ((rc = make_map(&wt, words_h, (uint64_t)D, (uint64_t)Bc * lpad, (uint32_t)lpad))) return rc;
  if ((rc = make_map(&gm, gram, (uint64_t)Spad, (uint64_t)Bi * Spad, TILE))) return rc;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bw::PairParams p{};
  p.wnorm = wnorm; p.cap_lens = cap_lens; p.sim = sim; p.mean_out = attn_mean; p.stats_out = stats;
  p.Bi = Bi; p.Bc = Bc; p.i0 = 0; p.nc = Bc; p.D = D; p.S = S; p.NT = Spad / TILE; p.lp = lp; p.sp = sp;
  p.t1 = temp1; p.t1_log2e = temp1 * 1.4426950408889634f; p.t2 = temp2; p.eps = eps; p.agg = agg;
  p.dbg = (long long*)g_phase_clock_buffer;
  return bw::launch_pair_lpad<true>(lpad, rt, wt, gm, gm /* E^T map unused: nothing is stored */, p, sms, st);
}

// Order of the backward: everything d_ctx needs first (dR GEMM, M-term, M.R, unpack) -- image part by image part, with
// an optional caller-owned event recorded after each part -- then the caption-side gradient (dW GEMM, gamma, unpack).
// A caption-sharded caller starts the reduce_scatter of a part's d_ctx rows on that part's event, so the collectives
// overlap the next part's GEMMs and the dW GEMM (distributed.py).
extern "C" int gloria_b200_tc_local_sim_bwd_train_parts(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                                        int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                                        const float* dsim, float* d_ctx, float* d_words, void* workspace,
                                                        size_t workspace_bytes, int n_parts, void* const* part_events,
                                                        void* stream) {
  GLORIA_CHECK_ARG(n_parts >= 1 && Bi % n_parts == 0, "n_parts=%d must divide Bi=%d", n_parts, Bi);
  // d_ctx may be NULL: the image-side gradient then stays in the workspace as dRt [Bi, sp, D] fp32
  // (gloria_b200_tc_train_drt_offset) -- see gloria_b200_tc_unpack_dctx
  GLORIA_CHECK_ARG(ctx_t && words_t && cap_lens && dsim && d_words && workspace, "null pointer");
  GLORIA_CHECK_ARG(Bi > 0 && Bc > 0 && word_off >= 0 && word_off + Lcap <= Lw, "bad sizes");
  if (gloria_b200_tc_supported(D, S, Lcap)) return fail(GLORIA_ERR_UNSUPPORTED, "shape D=%d S=%d Lcap=%d", D, S, Lcap);
  cudaStream_t st = (cudaStream_t)stream;
  const int Spad = gloria_b200_tc_spad(S), lp = gloria_b200_tc_lp(Lcap), sp = gloria_b200_tc_sp(S);
  const bw::TrainPlan pl = bw::train_plan(Bi, Bc, D, Spad, sp, lp);
  if (workspace_bytes < pl.total) return fail(GLORIA_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, pl.total);
  char* ws = (char*)workspace;
  __nv_bfloat16* X = (__nv_bfloat16*)(ws + pl.off_x);
  __nv_bfloat16* E = (__nv_bfloat16*)(ws + pl.off_e);
  float* gamma = (float*)(ws + pl.off_gamma);
  cublasHandle_t h = bw::cublas_handle();
  if (!h) return fail(GLORIA_ERR_DRIVER, "cublasCreate failed");
  GLORIA_CUBLAS(cublasSetStream(h, st));
  GLORIA_CUBLAS(cublasSetWorkspace(h, ws + pl.off_cublas, bw::CUBLAS_WS));
  const int R1 = Bc * lp;
  const dim3 sgrid((unsigned)((R1 / 8 + 255) / 256), (unsigned)(sp / bw::SCALE_ROWS), (unsigned)Bi);
  const float one = 1.f, zero = 0.f;
  const int K1 = Bi * sp;
  const __nv_bfloat16* Rt = (const __nv_bfloat16*)ctx_t;
  const __nv_bfloat16* Wt = (const __nv_bfloat16*)words_t;
  float* dWt = (float*)(ws + pl.off_dwt);
  float* dRt = (float*)(ws + pl.off_drt);
  float* Mf = (float*)(ws + pl.off_m);
  __nv_bfloat16* Mb = (__nv_bfloat16*)(ws + pl.off_mb);
  timer_record(GLORIA_TIMER_TC_BWD_GEMM, 0, st);
  const int gmode = bw::bwd_gemm_mode();
  // everything the forward stored is for g = 1 and linear in g = dsim[j, i]: the own GEMMs apply it to their A operand in
  // flight (mode 0); otherwise one streaming pass scales X in place first
  const int nj = Bi / n_parts;
  bw::SideStream* side = (gmode != 0 && bw::bwd_overlap_mode() != 0) ? bw::side_stream() : nullptr;
  if (side) {
    // fork: the M-term of every part goes to the side stream first (it depends on nothing the main stream does below)
    GLORIA_CUDA(cudaEventRecord(side->fork, st));
    GLORIA_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    for (int part = 0; part < n_parts; ++part) {
      const size_t j0 = (size_t)part * nj;
      int rc;
      if ((rc = launch_mterm(E + j0 * sp * R1, (const float*)(ws + pl.off_fo) + j0 * R1, dsim + j0 * Bc, Mf + j0 * sp * sp,
                             Mb + j0 * sp * sp, nj, Bc, 0, R1, lp, sp, false, side->st)))
        return rc;
    }
    GLORIA_CUDA(cudaEventRecord(side->join, side->st));
  }
  if (gmode != 0) {
    bw::scale_x<<<sgrid, 256, 0, st>>>(X, dsim, R1, sp, Bc, 0, lp, (const int*)(ws + pl.off_flag));
    GLORIA_LAUNCHED("scale_x");
    bw::mark_consumed<<<1, 1, 0, st>>>((int*)(ws + pl.off_flag));
    GLORIA_LAUNCHED("mark_consumed");
  }
  // ---- image side, part by part
  for (int part = 0; part < n_parts; ++part) {
    const size_t j0 = (size_t)part * nj;
    const __nv_bfloat16* Xp = X + j0 * sp * R1;
    float* dRp = dRt + j0 * sp * D;
    float* Mp = Mf + j0 * sp * sp;
    __nv_bfloat16* Mbp = Mb + j0 * sp * sp;
    const __nv_bfloat16* Rp = Rt + j0 * sp * D;
    // dRt[(j,s), d] = sum_(i,l) g[j,i] X^T[(j,s),(i,l)] Wt[(i,l), d]
    int rc;
    if (gmode == 0) {
      if ((rc = acc_gemm(Xp, Wt, dRp, nj * sp, D, R1, D, true, 1, false, dsim + j0 * Bc, Bc, 1, sp, lp, false, st))) return rc;
    } else if (gmode == 1) {
      if ((rc = acc_gemm(Xp, Wt, dRp, nj * sp, D, R1, D, true, 1, false, nullptr, 0, 0, 1, 1, false, st))) return rc;
    } else {
      GLORIA_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, D, nj * sp, R1, &one, Wt, CUDA_R_16BF, D, Xp, CUDA_R_16BF, R1,
                                 &zero, dRp, CUDA_R_32F, D, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
    }
    // M_j[a, b] = sum_(i,l) E^T[(j,a),(i,l)] g[j,i] f[j,(i,l)] E^T[(j,b),(i,l)]   (own tcgen05 kernel, tc_mterm.cu)
    // (the epilogue writes the bf16 operand of the M.R GEMM directly: no fp32 M, no conversion pass)
    if (side) {
      if (part == 0) GLORIA_CUDA(cudaStreamWaitEvent(st, side->join, 0));      // join: every M_j is final
    } else if ((rc = launch_mterm(E + j0 * sp * R1, (const float*)(ws + pl.off_fo) + j0 * R1, dsim + j0 * Bc, Mp, Mbp, nj,
                                  Bc, 0, R1, lp, sp, false, st))) {
      return rc;
    }
    // dRt_j += M_j Rt_j   (M_j symmetric up to rounding; column-major: [D, sp] = Rt_j^T . M_j)
    GLORIA_CUBLAS(cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_N, D, sp, sp, &one, Rp, CUDA_R_16BF, D,
                                             (long long)sp * D, Mbp, CUDA_R_16BF, sp, (long long)sp * sp, &one, dRp,
                                             CUDA_R_32F, D, (long long)sp * D, nj, CUBLAS_COMPUTE_32F,
                                             CUBLAS_GEMM_DEFAULT));
    if (d_ctx != nullptr) {
      bw::unpack_dctx<<<dim3((S + 31) / 32, D / 32, nj), dim3(32, 8), 0, st>>>(dRp, d_ctx + j0 * D * S, D, S, sp);
      GLORIA_LAUNCHED("unpack_dctx");
    }
    if (part_events && part_events[part]) GLORIA_CUDA(cudaEventRecord((cudaEvent_t)part_events[part], st));
  }
  // ---- caption side.  dWt[(i,l), d] = sum_(j,s) g[j,i] X^T[(j,s),(i,l)] Rt[(j,s), d]   (A^T = X^T in memory)
  {
    int rc;
    if (gmode == 0) {
      if ((rc = acc_gemm(X, Rt, dWt, R1, D, K1, D, false, 1, false, dsim, 1, Bc, lp, sp, false, st))) return rc;
    } else if (gmode == 1) {
      if ((rc = acc_gemm(X, Rt, dWt, R1, D, K1, D, false, 1, false, nullptr, 0, 0, 1, 1, false, st))) return rc;
    } else {
      GLORIA_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, D, R1, K1, &one, Rt, CUDA_R_16BF, D, X, CUDA_R_16BF, R1, &zero,
                                 dWt, CUDA_R_32F, D, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
    }
  }
  timer_record(GLORIA_TIMER_TC_BWD_GEMM, 1, st);
  bw::gamma_sum<<<(R1 + 31) / 32, 256, 0, st>>>((const float*)(ws + pl.off_go), dsim, gamma, Bi, R1, Bc, 0, lp);
  GLORIA_LAUNCHED("gamma_sum");
  bw::unpack_dwords_tc<<<dim3((Lw + 31) / 32, D / 32, Bc), dim3(32, 8), 0, st>>>(dWt, gamma, Wt, cap_lens, d_words, D,
                                                                               Lw, lp, Lcap, word_off);
  GLORIA_LAUNCHED("unpack_dwords_tc");
  return GLORIA_OK;
}

// Where the packed image-side gradient dRt [Bi, sp, D] fp32 (sp = gloria_b200_tc_sp(S) rows per image, rows >= S are
// padding) sits in the training workspace; 0 = unsupported shape.
extern "C" size_t gloria_b200_tc_train_drt_offset(int Bi, int Bc, int D, int S, int Lcap) {
  if (Bi <= 0 || Bc <= 0 || gloria_b200_tc_supported(D, S, Lcap)) return 0;
  return bw::train_plan(Bi, Bc, D, gloria_b200_tc_spad(S), gloria_b200_tc_sp(S), gloria_b200_tc_lp(Lcap)).off_drt;
}

// dRt [n, sp, D] fp32 (packed, as the backward leaves it) -> d_ctx [n, D, S] (the caller's layout).  A caption-sharded
// caller reduce_scatters the PACKED rows and unpacks only the images it owns (1 / world of the transposes).
extern "C" int gloria_b200_tc_unpack_dctx(const float* drt, float* d_ctx, int n, int D, int S, void* stream) {
  GLORIA_CHECK_ARG(drt && d_ctx, "null pointer");
  GLORIA_CHECK_ARG(n > 0 && n <= 65535 && D > 0 && D % 32 == 0 && S > 0, "bad sizes n=%d D=%d S=%d", n, D, S);
  bw::unpack_dctx<<<dim3((S + 31) / 32, D / 32, n), dim3(32, 8), 0, (cudaStream_t)stream>>>(drt, d_ctx, D, S,
                                                                                           gloria_b200_tc_sp(S));
  GLORIA_LAUNCHED("unpack_dctx");
  return GLORIA_OK;
}

extern "C" int gloria_b200_tc_local_sim_bwd_train_ev(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                                     int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                                     const float* dsim, float* d_ctx, float* d_words, void* workspace,
                                                     size_t workspace_bytes, void* d_ctx_ready_event, void* stream) {
  void* ev[1] = {d_ctx_ready_event};
  return gloria_b200_tc_local_sim_bwd_train_parts(ctx_t, words_t, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, dsim, d_ctx,
                                                  d_words, workspace, workspace_bytes, 1, ev, stream);
}

extern "C" int gloria_b200_tc_local_sim_bwd_train(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                                  int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                                  const float* dsim, float* d_ctx, float* d_words, void* workspace,
                                                  size_t workspace_bytes, void* stream) {
  return gloria_b200_tc_local_sim_bwd_train_ev(ctx_t, words_t, cap_lens, Bi, Bc, D, S, Lw, Lcap, word_off, dsim, d_ctx,
                                               d_words, workspace, workspace_bytes, nullptr, stream);
}

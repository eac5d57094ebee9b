// Accumulation GEMMs of the GLoRIA backward on sm_100a: persistent CTA-PAIR kernels (thread-block clusters of two,
// tcgen05.mma.cta_group::2, 256 x 256 output tiles, TMA-fed mbarrier rings), written for the two sums the closed-form
// backward leaves after the fused training kernel (autograd of the two bmm's, gloria/loss/gloria_loss.py:40,59):
//
//     dRt[(j,s), d] = sum_(i,l)  g[j,i] X^T[(j,s),(i,l)]  Wt[(i,l), d]          (image side,   K = Bc * lp)
//     dWt[(i,l), d] = sum_(j,s)  g[j,i] X^T[(j,s),(i,l)]  Rt[(j,s), d]          (caption side, K = Bi * sp)
//
// X^T was produced for g = 1 (everything in the backward is linear in g = dsim[j,i], which exists only after the two
// cross entropies).  The scale is applied IN FLIGHT, on the A operand of both GEMMs; nothing scaled is ever written to
// HBM (the former `scale_x` streaming pass, 40 GB per step, and both cuBLAS calls are gone):
//   * image side (`TS_AK`, A = X^T rows, K-major): per 64-wide k-block the A tile lands in shared memory by TMA, eight
//     scale warps (thread = tile row) multiply each 16-byte chunk by its g[j(row), i(chunk)] in fp32, round to bf16 and
//     write the tile into TENSOR MEMORY (tcgen05.st), from where the MMA takes its A operand
//     (`tcgen05.mma ... [a_tmem], b_desc`): the tensor pipe never reads A from shared memory, so the scaling costs no
//     shared-memory bandwidth over a plain GEMM.
//   * caption side (`SC_AM`, A = the same matrix read M-major, which cannot go through TMEM: A from tensor memory
//     cannot be transposed): the scale warps (thread = one k row of one 64-column block) rewrite the tile in place in
//     shared memory, fence.proxy.async, and the MMA reads it from there.
// The A tiles have their own TMA ring and producer thread so that they run ahead of the MMAs by the depth of the scale
// stage.
//
// Work units are (n-tile, m-tile, k-split) with the n-tiles of one row block adjacent, so the clusters that share an
// A panel run in step and the panel is read from HBM once; k-splits (tiles added with red.global.add.v4.f32) are used
// only when they fill the last wave markedly better.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gloria {
namespace tc {
namespace ag {

constexpr int BM = 128;                      // rows per CTA (the pair covers 256)
constexpr int BN = 256;                      // tile columns (each CTA stages 128 of them)
constexpr int BNH = BN / 2;
constexpr int BK = KBLK;                     // 64
constexpr int A_TILE = BM * BK * 2;          // 16 KB
constexpr int B_TILE = BK * BNH * 2;         // 16 KB
constexpr int NTHREADS = 384;                // warp 0 TMA (B, and A of the plain modes), warp 1 MMA (leader CTA), warp 2 TMEM
                                             // owner, warp 3 TMA of A in the scaled modes, warps 4-11 scale / epilogue
constexpr int NTA = 8;                       // TS mode: A stages in TMEM (32 columns each)
constexpr uint32_t A_COL = 256;              // TS mode: first TMEM column of the A ring (D occupies 0..255)
constexpr int NRDY = 8;                      // AREADY / AFREE barriers (>= NTA and >= any NA)

// SS_AK / SS_AM: plain GEMM, A row-major ([M, K]) / transposed ([K, M]) through shared memory.
// TS_AK: row-major A scaled by g on its way into tensor memory.   SC_AM: transposed A scaled by g in place in shared memory.
enum Mode { SS_AK = 0, SS_AM = 1, TS_AK = 2, SC_AM = 3 };

template <int MODE> struct Cfg;
template <> struct Cfg<SS_AK> { static constexpr int NA = 6, NB = 6, ACC = 2; };
template <> struct Cfg<SS_AM> { static constexpr int NA = 6, NB = 6, ACC = 2; };
template <> struct Cfg<TS_AK> { static constexpr int NA = 6, NB = 7, ACC = 1; };
template <> struct Cfg<SC_AM> { static constexpr int NA = 7, NB = 6, ACC = 2; };

template <int MODE>
struct Smem {
  using C = Cfg<MODE>;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = C::NA * A_TILE;
  static constexpr int OFF_BAR = OFF_B + C::NB * B_TILE;
  // barrier indices
  static constexpr int AFULL = 0;                       // [NA]  scaled modes: A tile landed (local)
  static constexpr int AEMPTY = AFULL + C::NA;          // [NA]  scaled modes: the slot may be refilled (TS: scale warps have read
                                                        //       it; SC: the MMAs that read it have completed)
  static constexpr int BFULL = AEMPTY + C::NA;          // [NB]  leader: operands of the stage landed in BOTH CTAs
  static constexpr int BEMPTY = BFULL + C::NB;          // [NB]  both: MMAs that read the stage have completed
  static constexpr int AREADY = BEMPTY + C::NB;         // [NRDY] leader: scaled A is in place in both CTAs (TS: TMEM stage,
                                                        //       SC: shared-memory slot)
  static constexpr int AFREE = AREADY + NRDY;           // [NRDY] TS, both: MMAs that read the TMEM A stage have completed
  static constexpr int ACCF = AFREE + NRDY;             // [2]   both: accumulator complete
  static constexpr int ACCE = ACCF + 2;                 // [2]   leader: accumulator drained by both CTAs' epilogues
  static constexpr int NBAR = ACCE + 2;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16 + 1024;
};

struct Params {
  int M, N, K;            // C[M, N] = A[M, K] B[K, N]
  int MT, NN, S, nkb;     // row blocks of 256, column tiles of 256, k-splits, k-blocks of 64 in all
  float* C;
  int ldc;
  int epi;                // 0: store, 1: red.global.add (k-splits / accumulation into a zeroed or partial result)
  // scaled modes: weight of element (m, k) = g[(m / m_div) * g_sm + (k / k_div) * g_sk]  (null: weight 1)
  const float* g;
  int g_sm, g_sk, m_div, k_div;
  float inv_m_div, inv_k_div;
  // plain modes only -- batches and operand planes (the split-precision fp32 mode, tc_f32.cu):
  //   unit = (n tile, row block, batch b); tiles of batch b sit a_brows / b_brows tensor-map rows further down, its C
  //   c_bstride elements further on.  The k-blocks are nterms runs of nkb_t: run t reads plane pa[t] of A and plane
  //   pb[t] of B (a_prows / b_prows tensor-map rows per plane), all into the same accumulator.
  int nb, a_brows, b_brows;
  long long c_bstride;
  int nterms, nkb_t, a_prows, b_prows;
  unsigned char pa[8], pb[8];
  int pf;                 // L2 prefetch distance in k-blocks (0 = off)
  int exp;                // development experiments (GLORIA_B200_GEMM_EXP): 1 = TS without the scale pipeline (static A)
  long long* dbg;         // phase clocks (only read when built with -DGLORIA_PHASE_CLOCKS)
};

// floor(x / d) for 0 <= x < 2^24 from the reciprocal (one fix-up step each way)
__device__ __forceinline__ int fdiv(int x, int d, float inv) {
  int q = __float2int_rd((float)x * inv);
  const int r = x - q * d;
  if (r >= d) ++q;
  if (r < 0) --q;
  return q;
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tm), "r"(x), "r"(y) : "memory");
}

// ---------------------------------------------------------------- cluster / 2-CTA PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose mbarrier may live in the peer CTA of the pair (complete_tx lands on `mbar_cluster`)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t mbar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(mbar_cluster)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem desc, 128 rows per CTA] * B[smem desc, N/2 columns per CTA]
__device__ __forceinline__ void umma_pair_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// ... with A taken from tensor memory (lanes = rows, 8 columns of packed bf16 pairs per K = 16)
__device__ __forceinline__ void umma_pair_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once every tcgen05 op issued so far has completed) on the barrier at this offset in the CTAs of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns, thread i of the warp writes TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

#ifdef GLORIA_PHASE_CLOCKS
#define GTIMED(acc, ...) do { long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; } while (0)
#else
#define GTIMED(acc, ...) do { __VA_ARGS__; } while (0)
#endif

// unit u -> (n tile, row block, k-split); the n tiles of a row block are adjacent (they share the A panel)
struct Unit {
  int n, m, b, kb0, kb1;
  __device__ Unit(int u, const Params& p) {
    n = u % p.NN;
    const int r = u / p.NN;
    m = r % p.MT;
    const int s = r / p.MT;
    if (p.nb > 1) {                      // batched: no k-splits
      b = s; kb0 = 0; kb1 = p.nkb;
    } else {
      b = 0;
      kb0 = (int)((long long)p.nkb * s / p.S);
      kb1 = (int)((long long)p.nkb * (s + 1) / p.S);
    }
  }
};

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
acc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  using L = Smem<MODE>;
  using C = Cfg<MODE>;
  constexpr bool SCALED = MODE == TS_AK || MODE == SC_AM;
  constexpr int NAd = C::NA > 0 ? C::NA : 1;          // (TS has no A slots in shared memory)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + L::OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::OFF_BAR + L::NBAR * 8);
  auto bar = [&](int idx) { return bars + 8u * (uint32_t)idx; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  const int nunits = p.MT * p.NN * p.S * p.nb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(bar(L::AFULL + s), 1); mbar_init(bar(L::AEMPTY + s), MODE == TS_AK ? 4 : 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(bar(L::BFULL + s), 1); mbar_init(bar(L::BEMPTY + s), 1); }
    for (int s = 0; s < NRDY; ++s) { mbar_init(bar(L::AREADY + s), 8); mbar_init(bar(L::AFREE + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar(L::ACCF + s), 1); mbar_init(bar(L::ACCE + s), 16); }
    fence_barrier_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 2) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: this CTA's half of B (+ its rows of A
    // in the plain modes); completion is counted on the LEADER's barrier
    if (lane == 0) {
      uint32_t nb = 0;
#ifdef GLORIA_PHASE_CLOCKS
      long long pt_all = clock64(), pw_b = 0;
#endif
      for (int u = cl; u < nunits; u += ncl) {
        const Unit un(u, p);
        const int m0 = un.m * 2 * BM + (int)rank * BM;
        const int n0 = un.n * BN + (int)rank * BNH;
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          const int sb = (int)(nb % C::NB);
          const uint32_t phb = (nb / C::NB) & 1u;
          ++nb;
          const uint32_t full_leader = mapa(bar(L::BFULL + sb), 0);
          if (p.pf > 0 && p.nterms <= 1 && p.nb <= 1 && kb + p.pf < un.kb1) {                 // pull the tiles of k-block kb + pf into L2 now
            const int kp = (kb + p.pf) * BK;
            tma_prefetch_2d(&tm_b, n0, kp);
            tma_prefetch_2d(&tm_b, n0 + 64, kp);
            if (MODE == SS_AK) tma_prefetch_2d(&tm_a, kp, m0);
            if (MODE == SS_AM) { tma_prefetch_2d(&tm_a, m0, kp); tma_prefetch_2d(&tm_a, m0 + 64, kp); }
          }
          GTIMED(pw_b, mbar_wait(bar(L::BEMPTY + sb), phb ^ 1));
          if (leader) mbar_expect_tx(bar(L::BFULL + sb), SCALED ? 2 * B_TILE : 2 * (A_TILE + B_TILE));
          // k-block -> (term, k-block inside the term) and the tensor-map rows of this batch's planes
          int kk = kb, aoff = 0, boff = 0;
          if (!SCALED) {
            if (p.nterms > 1) {
              const int t = kb / p.nkb_t;
              kk = kb - t * p.nkb_t;
              aoff = (int)p.pa[t] * p.a_prows;
              boff = (int)p.pb[t] * p.b_prows;
            }
            aoff += un.b * p.a_brows;
            boff += un.b * p.b_brows;
            const uint32_t adst = base + L::OFF_A + sb * A_TILE;
            if (MODE == SS_AK) {
              tma_load_2d_pair(adst, &tm_a, kk * BK, m0 + aoff, full_leader);
            } else {                         // A^T in memory: two [64 k x 64 m] boxes -> M-major tile
              tma_load_2d_pair(adst, &tm_a, m0, kk * BK + aoff, full_leader);
              tma_load_2d_pair(adst + A_TILE / 2, &tm_a, m0 + 64, kk * BK + aoff, full_leader);
            }
          }
          const uint32_t bdst = base + L::OFF_B + sb * B_TILE;
          tma_load_2d_pair(bdst, &tm_b, n0, kk * BK + boff, full_leader);
          tma_load_2d_pair(bdst + B_TILE / 2, &tm_b, n0 + 64, kk * BK + boff, full_leader);
        }
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (p.dbg) { long long* d = p.dbg + (size_t)blockIdx.x * 40 + 8; d[0] = clock64() - pt_all; d[2] = pw_b; }
#endif
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ scaled modes: TMA producer of A (own ring, local
    // barriers: the tile is consumed by this CTA's scale warps), decoupled from B so that it runs ahead of the MMAs
    if (SCALED && lane == 0 && p.exp != 1) {
      uint32_t na = 0;
#ifdef GLORIA_PHASE_CLOCKS
      long long pw_a = 0;
#endif
      for (int u = cl; u < nunits; u += ncl) {
        const Unit un(u, p);
        const int m0 = un.m * 2 * BM + (int)rank * BM;
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          const int sa = (int)(na % C::NA);
          const uint32_t pha = (na / C::NA) & 1u;
          ++na;
          GTIMED(pw_a, mbar_wait(bar(L::AEMPTY + sa), pha ^ 1));
          mbar_expect_tx(bar(L::AFULL + sa), A_TILE);
          const uint32_t adst = base + L::OFF_A + sa * A_TILE;
          if (MODE == TS_AK) {
            tma_load_2d(adst, &tm_a, kb * BK, m0, bar(L::AFULL + sa));
          } else {
            tma_load_2d(adst, &tm_a, m0, kb * BK, bar(L::AFULL + sa));
            tma_load_2d(adst + A_TILE / 2, &tm_a, m0 + 64, kb * BK, bar(L::AFULL + sa));
          }
        }
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (p.dbg) p.dbg[(size_t)blockIdx.x * 40 + 9] = pw_a;
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, (MODE == SS_AM || MODE == SC_AM) ? 1 : 0, 1);
      const uint64_t a_k0 = make_smem_desc(base + L::OFF_A, 16, 1024);          // K-major A tile
      const uint64_t a_m0 = make_smem_desc(base + L::OFF_A, A_TILE / 2, 1024);  // M-major A tile (two 64-row blocks)
      const uint64_t b_n0 = make_smem_desc(base + L::OFF_B, B_TILE / 2, 1024);  // N-major B tile (two 64-column blocks)
      uint32_t nb = 0, nt = 0;
#ifdef GLORIA_PHASE_CLOCKS
      long long mt_all = clock64(), mw_b = 0, mw_a = 0, mw_e = 0;
#endif
      for (int u = cl; u < nunits; u += ncl) {
        const Unit un(u, p);
        const int as = (int)(nt % C::ACC);
        const uint32_t pha = (nt / C::ACC) & 1u;
        ++nt;
        GTIMED(mw_e, mbar_wait(bar(L::ACCE + as), pha ^ 1));  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(as * BN);
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          const int sb = (int)(nb % C::NB);
          const uint32_t phb = (nb / C::NB) & 1u;
          // the A-side stage of this k-block: TMEM stage (TS), shared-memory slot (SC), the B stage itself (plain)
          const int sa = MODE == TS_AK ? (int)(nb % NTA) : MODE == SC_AM ? (int)(nb % NAd) : sb;
          const uint32_t pha2 = MODE == TS_AK ? (nb / NTA) & 1u : (nb / NAd) & 1u;
          ++nb;
          GTIMED(mw_b, mbar_wait(bar(L::BFULL + sb), phb));
          if (SCALED && p.exp != 1) GTIMED(mw_a, mbar_wait(bar(L::AREADY + sa), pha2));
          tc_fence_after();
          const uint64_t bd = b_n0 + (uint64_t)((sb * B_TILE) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t acc = (uint32_t)((kb != un.kb0) | (k != 0));
            if (MODE == TS_AK) {
              umma_pair_ts(d_tmem, tmem + A_COL + (uint32_t)(sa * 32 + k * 8), bd + 128 * k, idesc, acc);
            } else if (MODE == SS_AK) {
              umma_pair_ss(d_tmem, a_k0 + (uint64_t)((sa * A_TILE) >> 4) + 2 * k, bd + 128 * k, idesc, acc);
            } else {
              umma_pair_ss(d_tmem, a_m0 + (uint64_t)((sa * A_TILE) >> 4) + 128 * k, bd + 128 * k, idesc, acc);
            }
          }
          umma_commit_pair(bar(L::BEMPTY + sb), 3);
          if (MODE == TS_AK) umma_commit_pair(bar(L::AFREE + sa), 3);
          if (MODE == SC_AM) umma_commit_pair(bar(L::AEMPTY + sa), 3);
        }
        umma_commit_pair(bar(L::ACCF + as), 3);
      }
#ifdef GLORIA_PHASE_CLOCKS
      if (p.dbg) { long long* d = p.dbg + (size_t)blockIdx.x * 40; d[0] = clock64() - mt_all; d[1] = mw_b; d[2] = mw_a; d[3] = mw_e; d[4] = nb; }
#endif
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ scale warps (scaled modes) + epilogue
    const int grp = (warp - 4) >> 2;                  // two groups of four warps
    const int q = warp & 3;                           // lane quarter of tensor memory this warp may touch
    const int row = q * 32 + lane;                    // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t nk = 0, nt = 0;
#ifdef GLORIA_PHASE_CLOCKS
    long long st_all = clock64(), sw_full = 0, sw_free = 0, s_math = 0, s_st = 0, s_epi = 0, s_arr = 0;
#endif
    for (int u = cl; u < nunits; u += ncl) {
      const Unit un(u, p);
      const int m_glob = un.m * 2 * BM + (int)rank * BM + row;
      if (SCALED && p.exp != 1) {
        // k-blocks alternate between the two groups of four warps.
        // TS: thread = tile row (m fixed), the 8 chunks of its 128-byte row run along k.
        // SC: thread = (64-column block, k row): k fixed, the 8 chunks run along m.
        const int blk = row >> 6, krow = row & 63;
        const int m_base = un.m * 2 * BM + (int)rank * BM;
        // weight of each 8-wide chunk of k-block kb (chunks never straddle m_div / k_div along the contiguous axis).
        // The index along this thread's fixed axis is hoisted out of the k-block loop where it can be (TS: the row's m).
        const float* g_row = nullptr;                         // TS: g + (m / m_div) * g_sm, null when the row is padding
        int mi[8];                                            // SC: (m_c / m_div) * g_sm of the 8 chunks, -1 beyond M
        if (p.g != nullptr) {
          if (MODE == TS_AK) {
            if (m_glob < p.M) g_row = p.g + (size_t)fdiv(m_glob, p.m_div, p.inv_m_div) * p.g_sm;
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int m = m_base + blk * 64 + c * 8;
              mi[c] = m < p.M ? fdiv(m, p.m_div, p.inv_m_div) * p.g_sm : -1;
            }
          }
        }
        auto weights = [&](int kb, float* w) {
          if (p.g == nullptr) {
#pragma unroll
            for (int c = 0; c < 8; ++c) w[c] = 1.f;
          } else if (MODE == TS_AK) {
            const int k0 = kb * BK;
            int q0 = fdiv(k0, p.k_div, p.inv_k_div);
            int rem = k0 - q0 * p.k_div;                       // position of chunk 0 inside its k_div block
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              w[c] = (g_row != nullptr && k0 + c * 8 < p.K) ? __ldg(g_row + (size_t)q0 * p.g_sk) : 0.f;
              rem += 8;
              if (rem >= p.k_div) { rem -= p.k_div; ++q0; }
            }
          } else {
            const int k = kb * BK + krow;
            const float* gk = k < p.K ? p.g + (size_t)fdiv(k, p.k_div, p.inv_k_div) * p.g_sk : nullptr;
#pragma unroll
            for (int c = 0; c < 8; ++c) w[c] = (gk != nullptr && mi[c] >= 0) ? __ldg(gk + mi[c]) : 0.f;
          }
        };
        int kb = un.kb0 + (int)((grp + 2u - (nk & 1u)) & 1u);      // first k-block of this unit owned by this group
        float wn[8];
        if (kb < un.kb1) weights(kb, wn);
        for (; kb < un.kb1; kb += 2) {
          const uint32_t n = nk + (uint32_t)(kb - un.kb0);
          const int sa = (int)(n % C::NA);
          const uint32_t pha = (n / C::NA) & 1u;
          const int ta = (int)(n % NTA);
          const uint32_t pht = (n / NTA) & 1u;
          float w[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) w[c] = wn[c];
          if (kb + 2 < un.kb1) weights(kb + 2, wn);            // next k-block's weights: their latency is off the chain
          GTIMED(sw_full, mbar_wait(bar(L::AFULL + sa), pha));
#ifdef GLORIA_PHASE_CLOCKS
          long long _tm = clock64();
#endif
          uint8_t* src = smem + L::OFF_A + sa * A_TILE + (MODE == TS_AK ? (size_t)row * 128 : (size_t)blk * (A_TILE / 2) + (size_t)krow * 128);
          const int sw = MODE == TS_AK ? (row & 7) : (krow & 7);   // 128-byte swizzle: logical chunk c sits at c ^ (row & 7)
          // bf16x2 multiplies (one instruction per two elements; the weight is rounded to bf16 first, as the M-term
          // kernel does: a 2^-9 relative error on g[j,i], random over the 512 pairs a gradient element sums)
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));
            const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
            const __nv_bfloat162 w2 = __float2bfloat162_rn(w[c]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162 t = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&vi[k]), w2);
              r[c * 4 + k] = *reinterpret_cast<const uint32_t*>(&t);
            }
          }
          if (MODE == TS_AK) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(L::AEMPTY + sa));       // the slot may be refilled
#ifdef GLORIA_PHASE_CLOCKS
            s_math += clock64() - _tm;
#endif
            GTIMED(sw_free, mbar_wait(bar(L::AFREE + ta), pht ^ 1));   // MMAs of k-block n - NTA have read this TMEM stage
            tc_fence_after();
#ifdef GLORIA_PHASE_CLOCKS
            _tm = clock64();
#endif
            tmem_st32(tmem + lane_addr + A_COL + (uint32_t)(ta * 32), r);
            tmem_st_wait();
            tc_fence_before();
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(src + ((c ^ sw) << 4)) = make_uint4(r[c * 4], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]);
            fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor pipe's reads
#ifdef GLORIA_PHASE_CLOCKS
            s_math += clock64() - _tm;
            _tm = clock64();
#endif
          }
#ifdef GLORIA_PHASE_CLOCKS
          s_st += clock64() - _tm;
          _tm = clock64();
#endif
          __syncwarp();
          if (lane == 0) {
            const int rb = L::AREADY + (MODE == TS_AK ? ta : sa);
            if (leader) mbar_arrive(bar(rb));
            else mbar_arrive_cluster(mapa(bar(rb), 0));
          }
#ifdef GLORIA_PHASE_CLOCKS
          s_arr += clock64() - _tm;
#endif
        }
        nk += (uint32_t)(un.kb1 - un.kb0);
      }
      // ---- epilogue: this warp's 32 rows x 128 columns (group = column half) of the accumulator
      const int as = (int)(nt % C::ACC);
      const uint32_t pha = (nt / C::ACC) & 1u;
      ++nt;
#ifdef GLORIA_PHASE_CLOCKS
      long long _te = clock64();
#endif
      if (lane == 0) {                                          // one polling lane per warp, backing off: in the plain modes
        uint32_t spins = 0;                                     // this wait spans the whole main loop of a tile
        while (!mbar_try_wait(bar(L::ACCF + as), pha)) {
          __nanosleep(SCALED ? 64 : 512);
          if (++spins > (1u << 24)) __trap();
        }
      }
      __syncwarp();
      tc_fence_after();
      const int ncol0 = un.n * BN + grp * BNH;
      float* crow = p.C + (size_t)un.b * p.c_bstride + (size_t)m_glob * p.ldc + ncol0;
#pragma unroll 1
      for (int c = 0; c < BNH / 16; ++c) {
        float v[16];
        tmem_ld16(tmem + lane_addr + (uint32_t)(as * BN + grp * BNH + c * 16), v);
        tmem_ld_wait();
        if (m_glob < p.M && ncol0 + c * 16 < p.N) {           // N is a multiple of 16
          if (p.epi == 0) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              *reinterpret_cast<float4*>(crow + c * 16 + k4 * 4) = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
          } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              red_add_v4(crow + c * 16 + k4 * 4, v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar(L::ACCE + as));
        else mbar_arrive_cluster(mapa(bar(L::ACCE + as), 0));
      }
#ifdef GLORIA_PHASE_CLOCKS
      s_epi += clock64() - _te;
#endif
    }
#ifdef GLORIA_PHASE_CLOCKS
    if (p.dbg && q == 0 && lane == 0) {
      long long* d = p.dbg + (size_t)blockIdx.x * 40 + 16 + 8 * grp;
      d[0] = clock64() - st_all; d[1] = sw_full; d[2] = sw_free; d[3] = s_math; d[4] = s_st; d[5] = s_epi; d[6] = s_arr;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA frees tensor memory (or exits) while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

template <int MODE>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t st) {
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nunits = p.MT * p.NN * p.S * p.nb;
  int ncl = sms / 2;
  if (ncl > nunits) ncl = nunits;
  GLORIA_CUDA(cudaFuncSetAttribute(acc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<MODE>::BYTES));
  acc_gemm_kernel<MODE><<<2 * ncl, NTHREADS, Smem<MODE>::BYTES, st>>>(ma, mb, p);
  GLORIA_LAUNCHED("acc_gemm_kernel");
  return GLORIA_OK;
}

// number of k-splits that brings the unit count closest (from below) to a whole number of waves, at most `smax`
int pick_splits(int tiles, int ncl, int nkb, int smax) {
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= smax; ++s) {
    if (s > 1 && nkb / s < 64) break;              // keep units long against their epilogue
    const long long units = (long long)tiles * s;
    const long long waves = (units + ncl - 1) / ncl;
    const double eff = (double)units / (double)(waves * ncl);
    if (eff > best_eff + 0.08) { best_eff = eff; best = s; }   // a split costs a memset, atomics and L2 locality: only for a clear gain
  }
  return best;
}

}  // namespace ag

// C[b] [M, N] (fp32, row pitch ldc) = or += sum_t A[b, pa[t]] B[b, pb[t]] with bf16 operands in HBM.
//   a_kmajor: A planes are [M, K] (K contiguous); otherwise A^T is given, [K, M] (M contiguous).   B planes are [K, N].
//   g != null: A[m, k] is multiplied by g[(m / m_div) * g_sm + (k / k_div) * g_sk] in fp32 and rounded to bf16 on the way to
//   the tensor cores.  The divisor along the CONTIGUOUS axis of A must be a multiple of 8 (16-byte chunks carry one weight).
//   ksplit: 0 = choose; > 1 or accumulate: tiles are added with red.global.add (C must hold the value to add to).
//   Batches / planes (plain modes): see GemmEx in tc_common.cuh.  Every element inside the tensor-map extents must be
//   finite (rows read past a batch's K are multiplied by zero-filled columns); for transposed A with batches or planes
//   K must be a multiple of 64.
int acc_gemm_ex(const GemmEx& e, cudaStream_t st) {
  const int M = e.M, N = e.N, K = e.K;
  if (M <= 0 || N <= 0 || K <= 0 || (N & 15) || (K & 7) || (!e.a_kmajor && (M & 7)))
    return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: bad shape M=%d N=%d K=%d", M, N, K);
  if (e.g && (e.m_div <= 0 || e.k_div <= 0 || ((e.a_kmajor ? e.k_div : e.m_div) & 7)))
    return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: bad weight blocks m_div=%d k_div=%d", e.m_div, e.k_div);
  const int nb = e.nb > 0 ? e.nb : 1, nterms = e.nterms > 0 ? e.nterms : 1;
  const bool scaled = e.g != nullptr || e.force_scaled_path;
  const bool ext = nb > 1 || nterms > 1;
  if (ext && scaled) return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: batches / planes only in the plain modes");
  if (nterms != 1 && nterms != 3 && nterms != 6) return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: nterms %d", nterms);
  if (ext && !e.a_kmajor && (K & 63)) return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: transposed A with batches / planes needs K %% 64 == 0");
  ag::Params p{};
  p.M = M; p.N = N; p.K = K;
  p.MT = (M + 2 * ag::BM - 1) / (2 * ag::BM);
  p.NN = (N + ag::BN - 1) / ag::BN;
  p.nkb_t = (K + ag::BK - 1) / ag::BK;
  p.nkb = p.nkb_t * nterms;
  p.nb = nb; p.nterms = nterms;
  p.a_brows = (int)e.a_brows; p.b_brows = (int)e.b_brows; p.c_bstride = e.c_bstride;
  p.a_prows = (int)e.a_prows; p.b_prows = (int)e.b_prows;
  // split-precision terms, smallest first (tensor memory accumulates round-toward-zero: the large term goes last).
  // planes: 0 = leading bf16 piece, 1 = first residual, 2 = second residual
  static const unsigned char T6A[6] = {2, 1, 0, 1, 0, 0}, T6B[6] = {0, 1, 2, 0, 1, 0};
  static const unsigned char T3A[3] = {1, 0, 0}, T3B[3] = {0, 1, 0};
  for (int t = 0; t < nterms; ++t) {
    p.pa[t] = nterms == 6 ? T6A[t] : nterms == 3 ? T3A[t] : 0;
    p.pb[t] = nterms == 6 ? T6B[t] : nterms == 3 ? T3B[t] : 0;
  }
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  p.S = nb > 1 ? 1 : e.ksplit > 0 ? e.ksplit : ag::pick_splits(p.MT * p.NN, sms / 2, p.nkb, 8);
  if (p.S > p.nkb) p.S = p.nkb;
  p.C = e.C; p.ldc = e.ldc;
  p.epi = (p.S > 1 || e.accumulate) ? 1 : 0;
  p.g = e.g; p.g_sm = e.g_sm; p.g_sk = e.g_sk; p.m_div = e.m_div > 0 ? e.m_div : 1; p.k_div = e.k_div > 0 ? e.k_div : 1;
  p.inv_m_div = 1.0f / (float)p.m_div; p.inv_k_div = 1.0f / (float)p.k_div;
  static const int exp_mode = [] { const char* x = getenv("GLORIA_B200_GEMM_EXP"); return x ? atoi(x) : 0; }();
  static const int pf_dist = [] { const char* x = getenv("GLORIA_B200_GEMM_PREFETCH"); return x ? atoi(x) : 0; }();
  p.exp = exp_mode; p.pf = pf_dist;
  if (M >= (1 << 24) || K >= (1 << 24)) return fail(GLORIA_ERR_BAD_ARG, "acc_gemm: M, K must be below 2^24");
  p.dbg = (long long*)g_phase_clock_buffer;
  if (p.S > 1 && !e.accumulate) {
    if (e.ldc == N) GLORIA_CUDA(cudaMemsetAsync(e.C, 0, (size_t)M * e.ldc * sizeof(float), st));
    else GLORIA_CUDA(cudaMemset2DAsync(e.C, (size_t)e.ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  }
  // tensor-map extents in rows: all planes and batches of the operand
  const uint64_t a_unit = e.a_kmajor ? (uint64_t)M : (uint64_t)K, b_unit = (uint64_t)K;
  const uint64_t a_rows = e.a_rows > 0 ? (uint64_t)e.a_rows : a_unit, b_rows = e.b_rows > 0 ? (uint64_t)e.b_rows : b_unit;
  CUtensorMap ma, mb;
  int rc;
  if (e.a_kmajor) {
    if ((rc = make_map(&ma, e.A, (uint64_t)K, a_rows, ag::BM))) return rc;
  } else {
    if ((rc = make_map(&ma, e.A, (uint64_t)M, a_rows, ag::BK))) return rc;
  }
  if ((rc = make_map(&mb, e.B, (uint64_t)N, b_rows, ag::BK))) return rc;
  if (e.a_kmajor) return scaled ? ag::launch<ag::TS_AK>(ma, mb, p, st) : ag::launch<ag::SS_AK>(ma, mb, p, st);
  return scaled ? ag::launch<ag::SC_AM>(ma, mb, p, st) : ag::launch<ag::SS_AM>(ma, mb, p, st);
}

int acc_gemm(const void* A, const void* B, float* C, int M, int N, int K, int ldc, bool a_kmajor, int ksplit,
             bool accumulate, const float* g, int g_sm, int g_sk, int m_div, int k_div, bool force_scaled_path,
             cudaStream_t st) {
  GemmEx e{};
  e.A = A; e.B = B; e.C = C; e.M = M; e.N = N; e.K = K; e.ldc = ldc; e.a_kmajor = a_kmajor; e.ksplit = ksplit;
  e.accumulate = accumulate; e.g = g; e.g_sm = g_sm; e.g_sk = g_sk; e.m_div = m_div; e.k_div = k_div;
  e.force_scaled_path = force_scaled_path;
  return acc_gemm_ex(e, st);
}

}  // namespace tc
}  // namespace gloria

using namespace gloria;

// Exported for the parity tests (tests/test_gpu_acc_gemm.py) and for callers that want the primitive alone.
extern "C" int gloria_b200_acc_gemm(const void* A, const void* B, float* C, int M, int N, int K, int a_kmajor, int ksplit,
                                    int accumulate, const float* g, int g_sm, int g_sk, int m_div, int k_div,
                                    int force_scaled_path, void* stream) {
  GLORIA_CHECK_ARG(A && B && C, "null pointer");
  return tc::acc_gemm(A, B, C, M, N, K, N, a_kmajor != 0, ksplit, accumulate != 0, g, g_sm, g_sk, m_div, k_div,
                      force_scaled_path != 0, (cudaStream_t)stream);
}

// Split-precision form for the parity tests of the fp32 tensor-core mode (tc_f32.cu): dense operand planes
//   A [3][nb * (a_kmajor ? M : K), a_kmajor ? K : M],  B [3][nb * K, N]  (bf16 pieces of fp32 matrices),  C [nb][M, N];
//   nterms = 3 or 6 piece products, smallest first.
extern "C" int gloria_b200_acc_gemm_planes(const void* A, const void* B, float* C, int M, int N, int K, int a_kmajor, int nterms,
                                           int nb, int ksplit, int accumulate, void* stream) {
  GLORIA_CHECK_ARG(A && B && C && nb > 0, "null pointer / bad batch count");
  tc::GemmEx e{};
  const long long arows = a_kmajor ? M : K;
  e.A = A; e.B = B; e.C = C; e.M = M; e.N = N; e.K = K; e.ldc = N; e.a_kmajor = a_kmajor != 0; e.ksplit = ksplit;
  e.accumulate = accumulate != 0;
  e.nb = nb; e.a_brows = arows; e.b_brows = K; e.c_bstride = (long long)M * N;
  e.nterms = nterms; e.a_prows = arows * nb; e.b_prows = (long long)K * nb;
  e.a_rows = 3 * arows * nb; e.b_rows = 3LL * K * nb;
  return tc::acc_gemm_ex(e, (cudaStream_t)stream);
}

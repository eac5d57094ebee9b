// Gradient through |C_l| ("M term") of the GLoRIA backward on sm_100a tensor cores:
//     M_j[a, b] = sum_k  Eo^T[(j,a), k] * w[j, k] * Eo^T[(j,b), k],      k = (caption i, word l),  w = dsim[j,i] * f[j,k]
// for every image j (a, b = region rows), i.e. a batched  E diag(w) E^T  with K = nc*lp (53,248 at B = 512).
// A plain GEMM needs the scaled copy  Bo = diag(w) Eo  in HBM (another 21 GB written and read); here the scaling
// happens on the way: per 64-wide k-block the CTA loads the raw region tiles of image j by TMA, four warps scale ONE
// of them (the A operand) by w in shared memory (bf16x2 multiplies, swizzle-preserving copy), and tcgen05 multiplies
// it against the raw tiles.  M_j is symmetric: unit (j, a) computes only the blocks (a, b >= a) and mirrors them.
#include "common.cuh"
#include "tc_common.cuh"

namespace gloria {
namespace tc {
namespace mt {

constexpr int TILE_BYTES = TILE * 128;                           // 128 rows x 64 bf16
constexpr int NSTAGE = 3;
constexpr int OFF_AS = MAX_NT * TILE_BYTES;                      // scaled A tile inside a stage
constexpr int OFF_W = OFF_AS + TILE_BYTES;                       // 64 bf16 weights of the k-block
constexpr int STAGE = OFF_W + 1024;                              // keeps every stage 1024-byte aligned (SWIZZLE_128B)
constexpr int OFF_BAR = NSTAGE * STAGE;
enum { B_FULL = 0, B_SCALED = NSTAGE, B_EMPTY = 2 * NSTAGE, B_ACCF = 3 * NSTAGE, B_ACCE, B_COUNT };
constexpr int SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16 + 1024;
constexpr int NTHREADS = 384;                                    // warp 0 TMA, warp 1 MMA, warps 4-7 / 8-11 scale (k-blocks
                                                                 // alternate between the two groups), warps 4-7 epilogue

struct Params {
  const float* f;        // [Bi, R1]
  const float* g;        // [Bi, Bc] or nullptr (f already carries dsim)
  float* M;              // [Bi, sp, sp]
  __nv_bfloat16* Mb;     // [Bi, sp, sp] bf16 output instead of M (final, non-accumulating launch), or null
  int Bi, Bc, i0, R1, lp, sp, NT, accumulate;
};

// Per image the symmetric M_j needs the blocks (a, b >= a) of its NT x NT tile grid.  A CLUSTER OF TWO CTAs takes one image
// at a time and walks its K range once, in step: rank 0 owns row tile 0 (blocks (0, b)), rank 1 the row tiles 1..NT-1
// (blocks (1,1), (1,2), (2,2) at NT = 3) -- three 128 x 128 accumulators each at NT = 3, so the pair stays balanced.
// The tiles both CTAs need (1 .. NT-1) are loaded ONCE by rank 0's producer and MULTICAST into both CTAs' stages, so the
// E^T rows of an image cross L2 -> SM once per CTA but HBM -> L2 exactly once and the pair cannot drift apart: rank 0
// refills a stage only when BOTH CTAs' MMAs have released it (its `empty` barrier counts rank 1's commit too).  Before, the
// row tiles of an image were separate units (38.7 GB of DRAM reads per step for 20 GB of operand), then two unsynchronised
// CTAs (33.7 GB).
struct Role {
  int t0, nl, ns;            // first tile held, tiles held (tile t lives in slot t in BOTH CTAs), tiles scaled
  int ssrc[2], sdst[2];      // scale pass s: slot ssrc[s] -> slot sdst[s]
  int nm;                    // MMAs per k-step: (a slot, b slot) -> accumulator m; block (ra[m], rb[m]) of the tile grid
  int sa[3], sb[3], ra[3], rb[3];
};
__device__ __forceinline__ Role make_role(int rank, int NT) {
  Role r{};
  if (rank == 0) {           // row tile 0: A' = tile 0 scaled -> slot 3
    r.t0 = 0; r.nl = NT; r.ns = 1; r.nm = NT;
    r.ssrc[0] = 0; r.sdst[0] = 3;
    for (int b = 0; b < NT; ++b) { r.sa[b] = 3; r.sb[b] = b; r.ra[b] = 0; r.rb[b] = b; }
  } else {                   // row tiles 1..NT-1: tile 1 scaled -> slot 3, tile 2 scaled -> slot 0 (tile 0 is not held here)
    r.t0 = 1; r.nl = NT - 1; r.ns = NT - 1; r.nm = 0;
    for (int a = 1; a < NT; ++a) {
      r.ssrc[a - 1] = a; r.sdst[a - 1] = a == 1 ? 3 : 0;
      for (int b = a; b < NT; ++b) {
        r.sa[r.nm] = r.sdst[a - 1]; r.sb[r.nm] = b; r.ra[r.nm] = a; r.rb[r.nm] = b;
        ++r.nm;
      }
    }
  }
  return r;
}

// One TMA load delivered to the same shared-memory offset (and signalled on the mbarrier at the same offset) of every CTA
// in `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], "
      "[%4], %5;"
      ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
mterm_kernel(const __grid_constant__ CUtensorMap tm_e, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  auto bar = [&](int idx) { return bars + 8u * (uint32_t)idx; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT;
  const int nkb = (p.R1 + KBLK - 1) / KBLK;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const Role role = make_role((int)rank, NT);
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  const bool idle = role.nm == 0;                                  // NT = 1: the second CTA of the pair has no block

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar(B_FULL + s), 1);
      mbar_init(bar(B_SCALED + s), 256);       // owner group: A' written; other group: phase of `full` observed
      mbar_init(bar(B_EMPTY + s), (rank == 0 && NT > 1) ? 2 : 1);   // rank 0: its own MMAs' commit and its peer's
    }
    mbar_init(bar(B_ACCF), 1);
    mbar_init(bar(B_ACCE), 128);
    fence_barrier_init();
    tma_prefetch_desc(&tm_e);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (idle) {
    // nothing to do
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int j = cl; j < p.Bi; j += ncl) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(B_EMPTY + st), ph ^ 1);           // rank 0: released by both CTAs; rank 1: by its own MMAs
          mbar_expect_tx(bar(B_FULL + st), (uint32_t)role.nl * TILE_BYTES);
          if (rank == 0) {
            // tile 0 is rank 0's alone; tiles 1 .. NT-1 go to both CTAs with one load each
            tma_load_2d(base + st * STAGE, &tm_e, kb * KBLK, j * p.sp, bar(B_FULL + st));
            for (int t = 1; t < NT; ++t)
              tma_load_2d_mc(base + st * STAGE + t * TILE_BYTES, &tm_e, kb * KBLK, j * p.sp + t * TILE, bar(B_FULL + st), 3);
          }
          if (++st == NSTAGE) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE, TILE, 0, 0);   // A' (K-major) x raw tile (K-major)
      const uint64_t d0 = make_smem_desc(base, 16, 1024);
      int st = 0; uint32_t ph = 0, nu = 0;
      for (int j = cl; j < p.Bi; j += ncl) {
        mbar_wait(bar(B_ACCE), (nu & 1) ^ 1);                    // previous image's accumulators have been read
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(B_SCALED + st), ph);                     // tiles landed and A' written
          tc_fence_after();
          const uint64_t ds = d0 + (uint64_t)((st * STAGE) >> 4);
          for (int m = 0; m < role.nm; ++m) {
            const uint64_t da = ds + (uint64_t)((role.sa[m] * TILE_BYTES) >> 4);
            const uint64_t db = ds + (uint64_t)((role.sb[m] * TILE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + (uint32_t)(m * TILE), da + 2 * k, db + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          }
          if (rank == 0) umma_commit(bar(B_EMPTY + st));
          else umma_commit_mc(bar(B_EMPTY + st), 3);           // frees the stage here AND tells rank 0's producer
          if (++st == NSTAGE) { st = 0; ph ^= 1; }
        }
        umma_commit(bar(B_ACCF));
        ++nu;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ scale warps (thread = tile row) + epilogue
    // two groups of four warps take alternate k-blocks, so one group's wait -> scale -> publish chain overlaps the other's
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t64 = (threadIdx.x - 128) & 127;                   // 0..127 inside the group
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t nk = 0, nu = 0;                                     // running k-block / image counters of this CTA
    for (int j = cl; j < p.Bi; j += ncl) {
      const float* fr = p.f + (size_t)j * p.R1;
      const float* gr = p.g ? p.g + (size_t)j * p.Bc + p.i0 : nullptr;
      // weight of column t64 of a k-block; loaded one k-block (of this group) ahead: its L2 latency is off the chain
      auto load_w = [&](int kb) {
        float w = 0.f;
        const int k = kb * KBLK + t64;
        if (t64 < KBLK && kb < nkb && k < p.R1) w = __ldg(fr + k) * (gr ? __ldg(gr + k / p.lp) : 1.f);
        return w;
      };
      // Both groups observe EVERY k-block's `full` barrier: the owner (k-block parity == group) scales it, the other
      // group only needs to have seen the phase (its next wait on this stage is for the following phase, and a parity
      // wait cannot tell phase k from phase k+2).  Both arrive on `scaled`, so the MMA of k-block n -- and with it the
      // refill of its stage -- cannot run before BOTH groups have seen `full`(n): without that, a group coming out of
      // the epilogue could find its stage already two phases on and wait forever (seen as a launch failure with few
      // k-blocks per unit).
      int kb_mine = (int)((grp + 2u - (nk & 1u)) & 1u);          // first k-block of this image owned by this group
      float w_next = load_w(kb_mine);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t n = nk + (uint32_t)kb;
        const int st = (int)(n % NSTAGE);
        const uint32_t ph = (n / NSTAGE) & 1u;
        const bool mine = (n & 1u) == (uint32_t)grp;
        float w_cur = 0.f;
        if (mine) {
          w_cur = w_next;
          w_next = load_w(kb + 2);
        }
        // the stage's previous use has been released (the producer waited for it) once its tiles have landed
        mbar_wait(bar(B_FULL + st), ph);
        if (!mine) {
          mbar_arrive(bar(B_SCALED + st));
          continue;
        }
        uint8_t* sbase = smem + (size_t)st * STAGE;
        __nv_bfloat16* wsm = reinterpret_cast<__nv_bfloat16*>(sbase + OFF_W);
        if (t64 < KBLK) wsm[t64] = __float2bfloat16_rn(w_cur);
        if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
        // A' = (held tile) * w per column; 128-byte swizzle: chunk c of row r sits at c ^ (r & 7)
        for (int s = 0; s < role.ns; ++s) {
          const uint8_t* src = sbase + (size_t)role.ssrc[s] * TILE_BYTES + (size_t)row * 128;
          uint8_t* dst = sbase + (size_t)role.sdst[s] * TILE_BYTES + (size_t)row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int pc = (c ^ (row & 7)) << 4;
            const uint4 v = *reinterpret_cast<const uint4*>(src + pc);
            const uint4 wv = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(wsm) + c * 16);
            const uint32_t vi[4] = {v.x, v.y, v.z, v.w}, wi[4] = {wv.x, wv.y, wv.z, wv.w};
            uint32_t oi[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162 r2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&vi[k]),
                                                *reinterpret_cast<const __nv_bfloat162*>(&wi[k]));
              oi[k] = *reinterpret_cast<const uint32_t*>(&r2);
            }
            *reinterpret_cast<uint4*>(dst + pc) = make_uint4(oi[0], oi[1], oi[2], oi[3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar(B_SCALED + st));
      }
      nk += (uint32_t)nkb;
      if (grp == 0) {
        // ---- epilogue: this CTA's blocks from TMEM, each mirrored across the diagonal
        mbar_wait(bar(B_ACCF), nu & 1);
        tc_fence_after();
        float* Mj = p.M + (size_t)j * p.sp * p.sp;
        __nv_bfloat16* Mbj = p.Mb ? p.Mb + (size_t)j * p.sp * p.sp : nullptr;
        for (int m = 0; m < role.nm; ++m) {
          const int a = role.ra[m], b = role.rb[m];
          const int r_glob = a * TILE + row;
#pragma unroll 1
          for (int c = 0; c < TILE / 16; ++c) {
            float v[16];
            tmem_ld16(tmem + lane_addr + (uint32_t)(m * TILE + c * 16), v);
            tmem_ld_wait();
            const int col0 = b * TILE + c * 16;
            if (r_glob < p.sp && col0 + 16 <= p.sp) {            // sp is a multiple of 16: whole group in or out
              float4* d1 = reinterpret_cast<float4*>(Mj + (size_t)r_glob * p.sp + col0);
              if (p.accumulate) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const float4 t = d1[k4];
                  v[4 * k4] += t.x; v[4 * k4 + 1] += t.y; v[4 * k4 + 2] += t.z; v[4 * k4 + 3] += t.w;
                }
              }
              if (Mbj != nullptr) {                              // bf16 copy for the M.R GEMM (final values only)
                uint4* o = reinterpret_cast<uint4*>(Mbj + (size_t)r_glob * p.sp + col0);
                o[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                o[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
              } else {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) d1[k4] = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
              }
              if (b > a) {                                       // mirror (lanes write consecutive elements: coalesced)
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const size_t o2 = (size_t)(col0 + k) * p.sp + r_glob;
                  if (Mbj != nullptr) {
                    Mbj[o2] = __float2bfloat16_rn(v[k]);
                  } else {
                    // (with accumulate the mirror holds the same running value as the block itself: write, do not re-add)
                    Mj[o2] = v[k];
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(B_ACCE));
      }
      ++nu;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // no CTA leaves while its peer may still multicast to it or signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace mt

// M[j] (+)= Eo_j^T diag(w_j) Eo_j  for all images; Et is the [Bi*sp, R1] bf16 operand matrix (row-major)
int launch_mterm(const void* Et, const float* f, const float* g, float* M, void* Mb, int Bi, int Bc, int i0, int R1, int lp,
                 int sp, bool accumulate, cudaStream_t st) {
  CUtensorMap em;
  int rc;
  if ((rc = make_map(&em, Et, (uint64_t)R1, (uint64_t)Bi * sp, TILE))) return rc;
  mt::Params p;
  p.f = f; p.g = g; p.M = M; p.Mb = (__nv_bfloat16*)Mb; p.Bi = Bi; p.Bc = Bc; p.i0 = i0; p.R1 = R1; p.lp = lp; p.sp = sp;
  p.NT = (sp + TILE - 1) / TILE; p.accumulate = accumulate ? 1 : 0;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = 2 * min(sms / 2, Bi);                     // clusters of two: one image per cluster at a time
  GLORIA_CUDA(cudaFuncSetAttribute(mt::mterm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mt::SMEM_BYTES));
  mt::mterm_kernel<<<grid, mt::NTHREADS, mt::SMEM_BYTES, st>>>(em, p);
  GLORIA_LAUNCHED("mterm_kernel");
  return GLORIA_OK;
}

}  // namespace tc
}  // namespace gloria

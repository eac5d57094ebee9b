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
  int Bi, Bc, i0, R1, lp, sp, NT, accumulate;
};

__global__ void __launch_bounds__(NTHREADS, 1) mterm_kernel(const __grid_constant__ CUtensorMap tm_e, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  auto bar = [&](int idx) { return bars + 8u * (uint32_t)idx; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT;
  const int nkb = (p.R1 + KBLK - 1) / KBLK;
  const int nunits = p.Bi * NT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar(B_FULL + s), 1);
      mbar_init(bar(B_SCALED + s), 256);       // owner group: A' written; other group: phase of `full` observed
      mbar_init(bar(B_EMPTY + s), 1);
    }
    mbar_init(bar(B_ACCF), 1);
    mbar_init(bar(B_ACCE), 128);
    fence_barrier_init();
    tma_prefetch_desc(&tm_e);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int j = u / NT, a = u % NT;
        const int nb = NT - a;                                   // region tiles a .. NT-1
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(B_EMPTY + st), ph ^ 1);
          mbar_expect_tx(bar(B_FULL + st), (uint32_t)nb * TILE_BYTES);
          for (int b = 0; b < nb; ++b)
            tma_load_2d(base + st * STAGE + b * TILE_BYTES, &tm_e, kb * KBLK, j * p.sp + (a + b) * TILE, bar(B_FULL + st));
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
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int nb = NT - u % NT;
        mbar_wait(bar(B_ACCE), (nu & 1) ^ 1);                    // previous unit's accumulators have been read
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(B_SCALED + st), ph);                     // tiles landed and A' written
          tc_fence_after();
          const uint64_t ds = d0 + (uint64_t)((st * STAGE) >> 4);
          const uint64_t da = ds + (uint64_t)(OFF_AS >> 4);
          for (int b = 0; b < nb; ++b) {
            const uint64_t db = ds + (uint64_t)((b * TILE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + (uint32_t)(b * TILE), da + 2 * k, db + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(bar(B_EMPTY + st));
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
    uint32_t nk = 0, nu = 0;                                     // running k-block / unit counters of this CTA
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int j = u / NT, a = u % NT;
      const int nb = NT - a;
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
      int kb_mine = (int)((grp + 2u - (nk & 1u)) & 1u);          // first k-block of this unit owned by this group
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
        // A' = (tile a) * w per column; tile a is the stage's first tile; 128-byte swizzle: chunk c of row r sits at c ^ (r & 7)
        const uint8_t* src = sbase + (size_t)row * 128;
        uint8_t* dst = sbase + OFF_AS + (size_t)row * 128;
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
        fence_proxy_async_smem();
        mbar_arrive(bar(B_SCALED + st));
      }
      nk += (uint32_t)nkb;
      if (grp == 0) {
        // ---- epilogue: blocks (a, a+b) from TMEM, mirrored to (a+b, a)
        mbar_wait(bar(B_ACCF), nu & 1);
        tc_fence_after();
        const int r_glob = a * TILE + row;
        float* Mj = p.M + (size_t)j * p.sp * p.sp;
        for (int b = 0; b < nb; ++b) {
#pragma unroll 1
          for (int c = 0; c < TILE / 16; ++c) {
            float v[16];
            tmem_ld16(tmem + lane_addr + (uint32_t)(b * TILE + c * 16), v);
            tmem_ld_wait();
            const int col0 = (a + b) * TILE + c * 16;
            if (r_glob < p.sp) {
              if (col0 + 16 <= p.sp) {                           // sp is a multiple of 16: whole group in or out
                float4* d1 = reinterpret_cast<float4*>(Mj + (size_t)r_glob * p.sp + col0);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  float4 o = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
                  if (p.accumulate) { const float4 t = d1[k4]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                  d1[k4] = o;
                }
                if (b > 0) {                                     // mirror (lanes write consecutive floats: coalesced)
#pragma unroll
                  for (int k = 0; k < 16; ++k) {
                    float* d2 = Mj + (size_t)(col0 + k) * p.sp + r_glob;
                    *d2 = p.accumulate ? *d2 + v[k] : v[k];
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace mt

// M[j] (+)= Eo_j^T diag(w_j) Eo_j  for all images; Et is the [Bi*sp, R1] bf16 operand matrix (row-major)
int launch_mterm(const void* Et, const float* f, const float* g, float* M, int Bi, int Bc, int i0, int R1, int lp, int sp,
                 bool accumulate, cudaStream_t st) {
  CUtensorMap em;
  int rc;
  if ((rc = make_map(&em, Et, (uint64_t)R1, (uint64_t)Bi * sp, TILE))) return rc;
  mt::Params p;
  p.f = f; p.g = g; p.M = M; p.Bi = Bi; p.Bc = Bc; p.i0 = i0; p.R1 = R1; p.lp = lp; p.sp = sp;
  p.NT = (sp + TILE - 1) / TILE; p.accumulate = accumulate ? 1 : 0;
  int dev = 0, sms = 0;
  GLORIA_CUDA(cudaGetDevice(&dev));
  GLORIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = min(sms, Bi * p.NT);
  GLORIA_CUDA(cudaFuncSetAttribute(mt::mterm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mt::SMEM_BYTES));
  mt::mterm_kernel<<<grid, mt::NTHREADS, mt::SMEM_BYTES, st>>>(em, p);
  GLORIA_LAUNCHED("mterm_kernel");
  return GLORIA_OK;
}

}  // namespace tc
}  // namespace gloria

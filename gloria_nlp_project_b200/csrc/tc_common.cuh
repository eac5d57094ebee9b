// sm_100a building blocks written as inline PTX: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences) and the UMMA shared-memory + instruction descriptors.
// Bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) and the PTX ISA.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace gloria {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so a waiting
// warp does not compete for issue slots with the working warps of its scheduler.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU (each iteration may sleep up to
// the hint above, so the bound is tens of seconds).  No printf here (its call frame costs registers in every role).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar)
      : "memory");
}

// L2 eviction-priority policies for the bulk copies of the fused training kernel: its 45 GB write stream (X^T, E^T rows)
// would otherwise push the re-used region / Gram tiles out of L2 (22 GB of DRAM re-reads against 0.5 GB of inputs)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;"
      ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3,
                                                  uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
               : "memory");
}

// shared memory (SWIZZLE_128B box) -> global, 3-D tensor map; elements outside the tensor's extents are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on an mbarrier once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (Blackwell).
//   K-major  operand: rows of 64 bf16 (128 B); SBO = 1024 B between 8-row groups; LBO unused (1).
//   MN-major operand: 64 MN-elements (128 B) contiguous per K row, 8 K rows per 1024-B atom;
//                     SBO = stride between 8-K groups, LBO = stride between 64-wide MN blocks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: (BF16 x BF16 | F16 x F16) -> F32.  a/b format field: 0 = F16, 1 = BF16.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int ab_bf16 = 1) {
  return (1u << 4) | ((uint32_t)ab_bf16 << 7) | ((uint32_t)ab_bf16 << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------- shared by the forward and backward kernels
constexpr int KBLK = 64;                 // bf16 per 128-byte swizzle row
constexpr int TILE = 128;                // regions per score tile == words per context tile == channels per chunk
constexpr int MAX_NT = 3;                // Spad <= 384

// Static unit schedule shared by all warp roles.  A unit is (caption i, chunk of `c` consecutive images); units are
// numbered chunk-major (u = chunk * Bc + caption) and dealt round-robin to the CTAs, so in every round all CTAs sweep
// the same two or three image chunks in step and an image's region tiles are L2 hits for every SM but the first.
// The chunk length minimises rounds * (c + 1/4): the quarter pair is the cost of a unit switch (the operand pipeline
// runs across units; only the SIMT warps re-read the caption's masks).  At Bc = 64 captions x 512 images per rank
// (8-GPU shard of B = 512) this is 7 rounds of 32 images = 224 pair slots against 221.4 ideal (one caption per CTA
// with the images split in two would be 256).
struct Units {
  int Bi, Bc, ncta, c, nunits, u;
  int i, j, j_end;
  __device__ Units(int Bi_, int Bc_) : Bi(Bi_), Bc(Bc_), ncta(gridDim.x), i(0), j(0), j_end(0) {
    int best_c = Bi > 0 ? Bi : 1, best_n = 1;
    long long best = -1;
    const int kmax = Bi < 128 ? Bi : 128;
    for (int k = 1; k <= kmax; ++k) {
      const int cc = (Bi + k - 1) / k;
      const int nch = (Bi + cc - 1) / cc;
      const long long rounds = ((long long)Bc * nch + ncta - 1) / ncta;
      const long long cost = rounds * (4 * cc + 1);
      if (best < 0 || cost < best) { best = cost; best_c = cc; best_n = nch; }
    }
    c = best_c;
    nunits = Bc * best_n;
    u = (int)blockIdx.x - ncta;
  }
  __device__ bool next_caption() {
    u += ncta;
    if (u >= nunits) return false;
    const int ch = u / Bc;
    i = u - ch * Bc;
    j = ch * c;
    j_end = min(Bi, j + c);
    return true;
  }
};

extern void* g_phase_clock_buffer;
// Flat (caption, image) iterator over the same schedule, for roles that look one unit ahead.
struct UnitIter {
  Units u;
  int j;
  bool started;
  __device__ UnitIter(int Bi, int Bc) : u(Bi, Bc), j(0), started(false) {}
  __device__ bool next() {
    if (started && ++j < u.j_end) return true;
    started = true;
    if (!u.next_caption()) return false;
    j = u.j;
    return true;
  }
  __device__ int cap() const { return u.i; }
};

// host side (tc_local.cu): 2-D bf16 tensor map [rows, inner] (inner contiguous), box [box_rows, 64], SWIZZLE_128B
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint32_t box_rows);
// tc_mterm.cu: M[j] (+)= Eo_j^T diag(g[j,i] f[j,k]) Eo_j for all images (g may be null); Mb != null (and !accumulate):
// the result is written as bf16 to Mb instead of fp32 to M
int launch_mterm(const void* Et, const float* f, const float* g, float* M, void* Mb, int Bi, int Bc, int i0, int R1, int lp,
                 int sp, bool accumulate, cudaStream_t st);
// tc_gemm.cu, general form.  Plain modes take batches and operand planes: the operands are dense 2-D bf16 arrays of
// a_rows / b_rows rows in all ([rows, K] for row-major A, [rows, M] for transposed A, [rows, N] for B); batch b starts
// a_brows / b_brows rows further down (C: c_bstride elements further on), plane q a_prows / b_prows rows further down.
// nterms = 3 / 6: operands are given as bf16 pieces of fp32 values (plane 0 = leading piece, 1 / 2 = residuals) and the
// products hi.lo + lo.hi + hi.hi (3) or all products down to 2^-24 (6) are accumulated into one fp32 result.
struct GemmEx {
  const void* A; const void* B; float* C;
  int M, N, K, ldc;
  bool a_kmajor; int ksplit; bool accumulate;
  const float* g; int g_sm, g_sk, m_div, k_div; bool force_scaled_path;
  int nb; long long a_brows, b_brows, c_bstride;
  int nterms; long long a_prows, b_prows;
  long long a_rows, b_rows;          // tensor-map extents (0: one plane, one batch)
};
int acc_gemm_ex(const GemmEx& e, cudaStream_t st);
// tc_gemm.cu: C[M, N] (fp32, row pitch ldc) = or += A B on the CTA-pair tcgen05 GEMM; see the definition for the arguments
int acc_gemm(const void* A, const void* B, float* C, int M, int N, int K, int ldc, bool a_kmajor, int ksplit,
             bool accumulate, const float* g, int g_sm, int g_sk, int m_div, int k_div, bool force_scaled_path,
             cudaStream_t st);
// 4-D bf16 tensor [imgs, rows, mid, inner] (pitches in elements), box [1, box_rows, 1, 64], SWIZZLE_128B (TMA stores
// of E^T rows: clipped at `inner` columns per caption and at `rows` region rows per image)
int make_map4(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t mid, uint64_t rows, uint64_t imgs,
              uint64_t mid_pitch, uint64_t row_pitch, uint64_t img_pitch, uint32_t box_rows);
// 3-D bf16 tensor [rows, mid, inner] with pitches in elements, box [box_rows, 1, 64], SWIZZLE_128B (TMA stores)
int make_map3(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t mid, uint64_t rows, uint64_t mid_pitch,
              uint64_t row_pitch, uint32_t box_rows);

}  // namespace tc
}  // namespace gloria

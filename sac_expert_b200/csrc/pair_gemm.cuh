// CTA-pair (cta_group::2) tensor-core GEMM building block - NOT on the product path yet.
// One cluster of two CTAs computes C[256, 256] = A[256, K] . B[256, K]^T for one batch entry:
//   * CTA r stages rows [128 r, 128 r + 128) of A and rows [128 r, 128 r + 128) of B (its HALF of the N = 256 operand)
//     as bf16 hi / lo planes in its own shared memory (K-major SWIZZLE_128B, same layout as the single-CTA kernels);
//   * the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256: the pair's tensor cores read both
//     halves of B, every CTA accumulates its 128 rows x 256 columns in its own TMEM;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster signals the same mbarrier in both CTAs.
// So each operand byte is converted once per PAIR and each SM holds half of B - the mechanism the weight-stationary
// design needs.  Exercised by saceo_test_pair_gemm / tests/test_gpu_gather_gemm.py::test_pair_gemm.
#pragma once
#include "tc_gemm.cuh"

namespace saceo {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// arrives (once the pair's earlier MMAs have retired) on the mbarrier at the same shared offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

constexpr int PG_NT = 512, PG_N = 256;
constexpr int PG_PLANE = TC_BM * 128;                 // one 128-row x 64-k bf16 plane
constexpr int PG_MAIN = 4 * PG_PLANE > (PG_NT / 32) * 32 * (PG_N / 4 + 4) * 4 ? 4 * PG_PLANE : (PG_NT / 32) * 32 * (PG_N / 4 + 4) * 4;
constexpr int PG_BYTES = PG_MAIN + 1024 + 64;

// grid.x = 2 * batch, cluster (2,1,1), 512 threads.  A: [batch][256][K] row-major, B: [batch][256][K] row-major (op(B)^T),
// C: [batch][256][256].  K is a multiple of 64.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PG_NT, 1) k_pair_gemm(TcP q) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_hi = sb, a_lo = sb + PG_PLANE, b_hi = sb + 2 * PG_PLANE, b_lo = sb + 3 * PG_PLANE;
  const uint32_t bars = sb + PG_MAIN, tmem_slot = bars + 16;
  const GemmP& p = q.g;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* __restrict__ A = p.A + (long long)pair * p.sAa;
  const float* __restrict__ B = p.B + (long long)pair * p.sBa;
  const int m0 = (int)rank * TC_BM;                   // rows of A / C held by this CTA; also its half of B's rows

  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(PG_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  Slab<TC_BM, PG_NT> sa, sw;
  sa.init(A, p.lda, 1, m0, 2 * TC_BM);
  sw.init(B, p.ldb, 1, m0, 2 * TC_BM);
  const int nk = p.K / TC_BK;
  sa.ld(0, p.K); sw.ld(0, p.K);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                  // barriers initialised and TMEM allocated in both CTAs
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = lds_u32(tmem_slot);
  constexpr uint32_t IDESC = umma_idesc(2 * TC_BM, PG_N);       // M = 256 across the pair
  for (int kc = 0; kc < nk; ++kc) {
    if (kc > 0) mbar_wait(bars, (uint32_t)((kc - 1) & 1));      // the pair's MMAs of the previous slab released the planes
    sa.st(a_hi, a_lo); sw.st(b_hi, b_lo);
    if (kc + 1 < nk) { sa.ld((kc + 1) * TC_BK, p.K); sw.ld((kc + 1) * TC_BK, p.K); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                // both CTAs' planes are in place
    if (rank == 0 && threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t ko = kk * 32;
        umma_f16_pair(tmem, umma_desc(a_hi + ko), umma_desc(b_hi + ko), IDESC, (kc | kk) ? 1u : 0u);
        umma_f16_pair(tmem, umma_desc(a_hi + ko), umma_desc(b_lo + ko), IDESC, 1u);
        umma_f16_pair(tmem, umma_desc(a_lo + ko), umma_desc(b_hi + ko), IDESC, 1u);
      }
      umma_commit_pair(bars);
    }
  }
  mbar_wait(bars, (uint32_t)((nk - 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  tc_epilogue<PG_N, PG_NT>(q, sb, tmem, pair, 0, (long long)pair * p.sCa, m0, 0, warp, lane);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                  // nobody frees TMEM while the peer may still be read by the pair
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(PG_N) : "memory");
  }
}

}  // namespace saceo

// Hardware probe for the building blocks of the warp-specialised fused-MLP kernels (csrc/mlp_ws.cuh), run once on a
// B200 before the kernels were written (results in profiles/r2_umma_probe.log):
//   test 1: tcgen05.mma kind::f16 with an MN-MAJOR B operand (SWIZZLE_128B, LBO = stride between 64-element N atoms,
//           SBO = stride between 8-row K groups) read from the weight-plane image layout [s = n>>6][k row][64 n];
//   test 2: mixed operand formats in one MMA: A = bf16 planes (gradients), B = fp16 planes (weights), both K-major;
//   test 3: the B stage of test 1 delivered by cp.async.bulk (TMA engine) + mbarrier expect_tx instead of ld/st.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../sac_expert_b200/csrc/tc_gemm.cuh"

using namespace saceo;

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// a_fmt / b_fmt: 0 = f16, 1 = bf16; b_mn: B operand MN-major
__host__ __device__ constexpr uint32_t idesc2(int M, int N, uint32_t a_fmt, uint32_t b_fmt, uint32_t b_mn) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 1: A fp16 K-major [128][64] (k < K used), B = image stage [4][32][64] fp16 (MN-major), K = 32, N = 256
// mode 2: A bf16 K-major [128][64], B fp16 K-major [128 rows n][64 k], K = 64, N = 128
// mode 3: as 1 but the B stage arrives by cp.async.bulk
__global__ void __launch_bounds__(128, 1) k_probe(int mode, const uint16_t* __restrict__ Aimg, const uint16_t* __restrict__ Bimg,
                                                  float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sb, sB = sb + 16384, bars = sb + 16384 + 32768, slot = bars + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1); mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A image: 16 KB, copied verbatim (already swizzled on the host)
  for (int i = threadIdx.x; i < 16384 / 16; i += 128)
    sts128(sA + i * 16, reinterpret_cast<const uint4*>(Aimg)[i]);
  const int bbytes = mode == 2 ? 16384 : 16384;   // mode 1/3: [4][32][128 B] = 16 KB; mode 2: [128][128 B] = 16 KB
  if (mode != 3) {
    for (int i = threadIdx.x; i < bbytes / 16; i += 128)
      sts128(sB + i * 16, reinterpret_cast<const uint4*>(Bimg)[i]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = lds_u32(slot);
  if (threadIdx.x == 0) {
    if (mode == 3) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8), "r"(16384u) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sB), "l"(Bimg), "r"(16384u), "r"(bars + 8) : "memory");
      mbar_wait(bars + 8, 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 2) {
      const uint32_t id = idesc2(128, 128, 1u, 0u, 0u);
      for (int kk = 0; kk < 4; ++kk)
        umma_f16(tmem, umma_desc(sA + kk * 32), umma_desc(sB + kk * 32), id, kk ? 1u : 0u);
    } else {
      const uint32_t id = idesc2(128, 256, 0u, 0u, 1u);
      for (int kk = 0; kk < 2; ++kk)     // K = 16 per MMA = two 8-row groups = 2048 B of the stage
        umma_f16(tmem, umma_desc(sA + kk * 32), desc_mn(sB + kk * 2048, 4096, 1024), id, kk ? 1u : 0u);
    }
    umma_commit(bars);
  }
  mbar_wait(bars, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int N = mode == 2 ? 128 : 256;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

static uint16_t f2h(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static uint16_t f2b(float x) { __nv_bfloat16 h = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float b2f(uint16_t u) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }

int main() {
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  int fails = 0;
  for (int mode : {1, 3, 2}) {     // the mixed-format MMA last: it traps (illegal instruction) on sm_100a
    const int N = mode == 2 ? 128 : 256, K = mode == 2 ? 64 : 32;
    std::vector<float> A(128 * 64, 0.f), B((size_t)K * N);
    std::vector<uint16_t> Aimg(128 * 64, 0), Bimg(8192, 0);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) {
      const float x = rnd();
      const uint16_t u = mode == 2 ? f2b(x) : f2h(x);
      A[m * 64 + k] = mode == 2 ? b2f(u) : h2f(u);
      // K-major SW128: row m: 128 B, chunk (k>>3) ^ (m&7)
      Aimg[(m >> 3) * 512 + (m & 7) * 64 + (((k >> 3) ^ (m & 7)) << 3) + (k & 7)] = u;
    }
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) {
      const uint16_t u = f2h(rnd());
      B[(size_t)k * N + n] = h2f(u);
      if (mode == 2) {   // K-major: row n, 64 k
        Bimg[(n >> 3) * 512 + (n & 7) * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = u;
      } else {           // plane image stage [s = n>>6][r = k (32)][64 n], 16-byte chunk index XOR (r & 7)
        const int s = n >> 6, c = (n & 63) >> 3;
        Bimg[s * 2048 + k * 64 + ((c ^ (k & 7)) << 3) + (n & 7)] = u;
      }
    }
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 256 * 4);
    cudaMemcpy(dA, Aimg.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 256 * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_probe<<<1, 128, 65536>>>(mode, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost);
    double worst = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[m * 64 + k] * B[(size_t)k * N + n];
      worst = fmax(worst, fabs(acc - D[m * N + n])); ref_max = fmax(ref_max, fabs(acc));
    }
    const bool ok = worst < 1e-4 * ref_max + 1e-5;
    printf("mode %d (%s): max abs err %.3e (ref max %.3f) %s\n", mode,
           mode == 1 ? "MN-major B, ld/st staged" : mode == 2 ? "A bf16 x B fp16" : "MN-major B via cp.async.bulk",
           worst, ref_max, ok ? "OK" : "MISMATCH");
    fails += !ok;
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf(fails ? "PROBE_FAIL\n" : "PROBE_OK\n");
  return fails ? 1 : 0;
}

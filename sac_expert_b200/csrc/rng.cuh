// Philox4x32-10 counter RNG for the in-kernel draws of "perf mode":
// minibatch indices (np.random.randint, sac_eo/common/buffers.py:135), Gaussian action noise
// (np.random.normal, sac_eo/actors/continuous_actors.py:297,350) and the expert-row shuffle
// (self.rng.shuffle, sac_eo/algs/SAC_expert.py:301-303).  The reference draws from NumPy MT19937 /
// PCG64 streams that cannot be reproduced on device, so parity runs inject the draws
// (saceo_set_draws) and this generator is only tested statistically.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace saceo {

struct Philox {
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint64_t p0 = (uint64_t)M0 * c[0];
    uint64_t p1 = (uint64_t)M1 * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  // counter = (i, agent, step, stream), key = 64-bit seed
  __host__ __device__ static inline void gen(uint64_t seed, uint32_t i, uint32_t agent, uint32_t step,
                                             uint32_t stream, uint32_t (&out)[4]) {
    uint32_t c[4] = {i, agent, step, stream};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) round(c, k);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// (0,1] uniform from 32 bits
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// four standard normals from one Philox block (Box-Muller)
__device__ __forceinline__ void normal4(const uint32_t (&r)[4], float (&z)[4]) {
  float u0 = u01(r[0]), u1 = u01(r[1]), u2 = u01(r[2]), u3 = u01(r[3]);
  float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u1, &s0, &c0);
  sincosf(6.283185307179586f * u3, &s1, &c1);
  z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
}

// uniform integer in [0, n) by multiply-high (bias <= n / 2^32, documented in DESIGN.md)
__device__ __forceinline__ uint32_t below(uint32_t x, uint32_t n) {
  return (uint32_t)(((uint64_t)x * (uint64_t)n) >> 32);
}

}  // namespace saceo

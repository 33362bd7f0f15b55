// Weight-plane image layout shared by the optimiser kernels (elem.cuh::k_adam) and the warp-specialised fused MLP
// kernels (mlp_ws.cuh).  Image of a [K x 256] matrix W[i][j] (fp16 hi plane + fp16 lo plane, lo = rn16(w - hi)):
//     [t = i>>5][plane hi|lo][s = j>>6][r = i&31][64 j]      16-byte chunks XOR (r & 7)   (SWIZZLE_128B atoms)
// One net = the W0 stages (32 input rows each, zero-padded to a multiple of 32) followed by the 8 W1 stages.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace saceo {

constexpr int WS_STAGE = 32768;

// bytes of the plane image of one net with K0 inputs: W0 stages (32 input rows each, zero-padded) then the 8 W1 stages
__host__ __device__ inline long long ws_image_bytes(int K0) { return (long long)((K0 + 31) / 32 + 8) * WS_STAGE; }
// byte offset (hi plane; lo plane = +16384) of the 16-byte chunk holding W[row][8*c8 .. 8*c8+7]; row counts from the
// first W0 row, the W1 rows follow at row index 32 * n0
__host__ __device__ inline long long ws_image_off(int row, int c8) {
  const int t = row >> 5, r = row & 31, s = c8 >> 3, c = c8 & 7;
  return (long long)t * WS_STAGE + s * 4096 + r * 128 + ((c ^ (r & 7)) << 4);
}

// used by k_adam: the 4 consecutive parameters at flat index e (multiple of 4) of a net with K0 inputs -> image bytes
__device__ __forceinline__ void ws_planes_store4(uint8_t* __restrict__ img, int K0, long long e, const float (&x)[4]) {
  const long long oW1 = (long long)K0 * 256 + 256;
  int row, col;
  if (e < (long long)K0 * 256) { row = (int)(e >> 8); col = (int)(e & 255); }
  else if (e >= oW1 && e < oW1 + 65536) { const int e2 = (int)(e - oW1); row = ((K0 + 31) / 32) * 32 + (e2 >> 8); col = e2 & 255; }
  else return;
  uint32_t h0, h1, l0, l1;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(x[1]), "f"(x[0]));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(x[3]), "f"(x[2]));
  const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l0) : "f"(x[1] - f0.y), "f"(x[0] - f0.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l1) : "f"(x[3] - f1.y), "f"(x[2] - f1.x));
  uint8_t* dst = img + ws_image_off(row, col >> 3) + (col & 7) * 2;
  *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(dst + 16384) = make_uint2(l0, l1);
}


}  // namespace saceo

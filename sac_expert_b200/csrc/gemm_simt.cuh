// Generic batched fp32 GEMM on CUDA cores with a fused epilogue.
//
// One launch covers every (agent, net) pair of the population: blockIdx.z = agent * nnet + net.
// It is the exact-fp32 engine (gemm_mode 0) and, in every mode, the engine for the awkward
// shapes that do not belong on tensor cores: first layers with K = S or S+A in
// {11,14,17,23,27,35,...}, heads with N in {1, A, 2A, S+1}, and the E/2-row expert GEMMs.
//
//   C[M,N] = epi( op(A)[M,K] . op(B)[K,N] )       row-major storage, leading dims lda/ldb/ldc
//   ONES:  A is transposed storage X[K, M-1] and row M-1 of op(A) is all ones, so that
//          [dW ; db] = [X, 1]^T . dY lands directly in the flat Keras layout [W | b]
//          (sac_eo/common/nn_utils.py:162-182).
//   epilogue:  v = acc (+ bias[n]) (+ addend[m,n]);  epi 1: v = act(v);  epi 2: v *= act'(aux[m,n])
//          where act' is expressed from the POST-activation value (relu: h>0, tanh: 1-h^2,
//          elu: h>0 ? 1 : h+1).  addend and aux share C's shape and strides.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace saceo {

enum { EPI_NONE = 0, EPI_ACT = 1, EPI_MUL_DACT = 2 };
enum { ACT_RELU = 0, ACT_TANH = 1, ACT_ELU = 2, ACT_LINEAR = 3 };

struct GemmP {
  const float* A; const float* B; float* C;
  const float* bias; const float* addend; const float* aux;
  int M, N, K;
  int lda, ldb, ldc;
  long long sAa, sAn;   // agent / net strides (floats)
  long long sBa, sBn;
  long long sCa, sCn;
  long long sba, sbn;   // bias
  int nnet;
  int epi, act;
  int m_off;            // first output row handled by this launch (tail launches after a tensor-core main part)
  int f16;              // tensor-core engine only: split operands into fp16 planes (forward passes) instead of bf16
};

__device__ __forceinline__ float apply_act(int act, float v) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_TANH: return tanhf(v);
    case ACT_ELU:  return v > 0.f ? v : expm1f(v);
    default:       return v;
  }
}
__device__ __forceinline__ float dact_from_out(int act, float h) {
  switch (act) {
    case ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case ACT_TANH: return 1.f - h * h;
    case ACT_ELU:  return h > 0.f ? 1.f : h + 1.f;
    default:       return 1.f;
  }
}

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <bool TA, bool TB, bool ONES>
__global__ void __launch_bounds__(SG_THREADS) k_gemm_simt(GemmP p) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
  const int z = blockIdx.z;
  const int agent = z / p.nnet, net = z - agent * p.nnet;
  const float* __restrict__ A = p.A + agent * p.sAa + net * p.sAn;
  const float* __restrict__ B = p.B + agent * p.sBa + net * p.sBn;
  const long long offC = agent * p.sCa + net * p.sCn;
  float* __restrict__ C = p.C + offC;
  const int m0 = p.m_off + blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
    // ---- stage A tile: As[k][m] = op(A)(m0+m, k0+k)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (TA) { m = t & 63; k = (t >> 6) + 4 * i; }   // storage contiguous in m
      else    { k = t & 15; m = (t >> 4) + 16 * i; }  // storage contiguous in k
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < p.M && gk < p.K) {
        if (ONES && gm == p.M - 1) v = 1.f;
        else v = TA ? A[(long long)gk * p.lda + gm] : A[(long long)gm * p.lda + gk];
      }
      As[k][m] = v;
    }
    // ---- stage B tile: Bs[k][n] = op(B)(k0+k, n0+n)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (TB) { k = t & 15; n = (t >> 4) + 16 * i; }  // storage contiguous in k
      else    { n = t & 63; k = (t >> 6) + 4 * i; }   // storage contiguous in n
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < p.N && gk < p.K) v = TB ? B[(long long)gn * p.ldb + gk] : B[(long long)gk * p.ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = p.bias ? p.bias + agent * p.sba + net * p.sbn : nullptr;
  const float* addend = p.addend ? p.addend + offC : nullptr;
  const float* aux = p.aux ? p.aux + offC : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      const long long o = (long long)gm * p.ldc + gn;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (addend) v += addend[o];
      if (p.epi == EPI_ACT) v = apply_act(p.act, v);
      else if (p.epi == EPI_MUL_DACT) v *= dact_from_out(p.act, aux[o]);
      C[o] = v;
    }
  }
}

}  // namespace saceo

// Skinny batched GEMM: one of the two output dimensions is <= 32.
//
// Covers every contraction of the update that is bandwidth- or latency-bound rather than
// tensor-core material: the E (<= 32) expert rows through the actor and the 2x512 dynamics models
// (weights streamed once, frozen), the ones-row of the [dW; db] trick (bias gradients = column sums),
// the row tails left over by the 128-row tensor-core tiles, and - with the roles of the two output
// dimensions swapped - the thin heads N in {1, A, 2A} and the backward-to-action slices (N = A).
//
//   C(i, j) = epi( sum_k A(i,k) B(k,j) ),  i < Ms <= 32 ("short" dim),  j < Nl ("long" dim)
//   generic element strides for all three operands, so transposes and role swaps are free.
// One CTA = 128 threads = 128 consecutive j; every thread keeps the Ms accumulators of its column in
// registers, the short operand's k-panel is broadcast from shared memory.
#pragma once
#include "gemm_simt.cuh"

namespace saceo {

struct SkinnyP {
  const float* A; const float* B; float* C;
  const float* bias; const float* addend; const float* aux;
  int Ms, Nl, K;
  long long a_si, a_sk, b_sk, b_sj, c_si, c_sj;
  long long sAa, sAn, sBa, sBn, sCa, sCn, sba, sbn;
  int nnet, epi, act;
  int bias_on_i;     // bias indexed by the short index i (role-swapped heads) or by j
  int ones_row;      // short row index whose A(i,.) is all ones (-1: none)
  int i_off;         // first short row handled (row tails)
  int ones_col;      // long index j whose B(.,j) is all ones (-1: none) - role-swapped [dW; db] GEMMs
};

constexpr int SK_THREADS = 128, SK_KT = 32;

// BT: the long operand is contiguous along k (b_sk == 1): its 128 x 32 panel is read with lanes along k
// (coalesced 128-byte rows) and transposed through shared memory; otherwise lanes already walk j.
template <int MS, bool BT>   // MS: compile-time bound on the short dim: 1, 4, 8, 16, 32
__global__ void __launch_bounds__(SK_THREADS) k_gemm_skinny(SkinnyP p) {
  __shared__ __align__(16) float As[SK_KT][MS];
  __shared__ float Bs[BT ? SK_KT : 1][BT ? SK_THREADS + 1 : 1];
  const int z = blockIdx.z;
  const int agent = z / p.nnet, net = z - agent * p.nnet;
  const float* __restrict__ A = p.A + agent * p.sAa + net * p.sAn;
  const float* __restrict__ B = p.B + agent * p.sBa + net * p.sBn;
  const long long offC = agent * p.sCa + net * p.sCn;
  const int j = blockIdx.x * SK_THREADS + threadIdx.x;
  const bool jin = j < p.Nl;
  const int ms = p.Ms - p.i_off < MS ? p.Ms - p.i_off : MS;
  float acc[MS];
#pragma unroll
  for (int i = 0; i < MS; ++i) acc[i] = 0.f;
  const float* bcol = B + (long long)j * p.b_sj;
  for (int k0 = 0; k0 < p.K; k0 += SK_KT) {
    // stage the short operand's k-panel (MS x 32)
    for (int e = threadIdx.x; e < SK_KT * MS; e += SK_THREADS) {
      int i, k;
      if (p.a_sk == 1) { k = e % SK_KT; i = e / SK_KT; } else { i = e % MS; k = e / MS; }
      const int gi = p.i_off + i, gk = k0 + k;
      float v = 0.f;
      if (i < ms && gk < p.K) v = (gi == p.ones_row) ? 1.f : __ldg(A + (long long)gi * p.a_si + (long long)gk * p.a_sk);
      As[k][i] = v;
    }
    const bool kfull = (k0 + SK_KT <= p.K);
    if (BT) {
      const int j0 = blockIdx.x * SK_THREADS;
      const int kk = threadIdx.x & (SK_KT - 1), jr = threadIdx.x / SK_KT;      // 4 rows of 32 k per pass
      float t[SK_KT];
#pragma unroll
      for (int i = 0; i < SK_KT; ++i) {            // 32 independent loads in flight per thread
        const int gj = j0 + jr + 4 * i;
        const bool ok = (gj < p.Nl) && (kfull || k0 + kk < p.K) && (gj != p.ones_col);
        const float* q = B + (long long)(ok ? gj : 0) * p.b_sj + (ok ? k0 + kk : 0);
        const float v = __ldg(q);
        t[i] = ok ? v : ((gj == p.ones_col && gj < p.Nl && k0 + kk < p.K) ? 1.f : 0.f);
      }
#pragma unroll
      for (int i = 0; i < SK_KT; ++i) Bs[kk][jr + 4 * i] = t[i];
    }
    __syncthreads();
    float b[SK_KT];
    if (BT) {
#pragma unroll
      for (int k = 0; k < SK_KT; ++k) b[k] = Bs[k][threadIdx.x];
    } else if (jin && kfull && j != p.ones_col) {
      const float* q = bcol + (long long)k0 * p.b_sk;
#pragma unroll
      for (int k = 0; k < SK_KT; ++k) b[k] = __ldg(q + (long long)k * p.b_sk);   // unconditional: all in flight
    } else {
#pragma unroll
      for (int k = 0; k < SK_KT; ++k)
        b[k] = (jin && k0 + k < p.K) ? (j == p.ones_col ? 1.f : __ldg(bcol + (long long)(k0 + k) * p.b_sk)) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < SK_KT; ++k) {
#pragma unroll
      for (int i = 0; i < MS; ++i) acc[i] = fmaf(As[k][i], b[k], acc[i]);
    }
    __syncthreads();
  }
  if (!jin) return;
  const float* bias = p.bias ? p.bias + agent * p.sba + net * p.sbn : nullptr;
  const float* addend = p.addend ? p.addend + offC : nullptr;
  const float* aux = p.aux ? p.aux + offC : nullptr;
  float* __restrict__ C = p.C + offC;
#pragma unroll
  for (int i = 0; i < MS; ++i) {
    if (i >= ms) break;
    const int gi = p.i_off + i;
    const long long o = (long long)gi * p.c_si + (long long)j * p.c_sj;
    float v = acc[i];
    if (bias) v += __ldg(bias + (p.bias_on_i ? gi : j));
    if (addend) v += addend[o];
    if (p.epi == EPI_ACT) v = apply_act(p.act, v);
    else if (p.epi == EPI_MUL_DACT) v *= dact_from_out(p.act, aux[o]);
    C[o] = v;
  }
}

// Builds the skinny problem for rows [i_off, M) of a GemmP whose M side is short ...
static inline SkinnyP skinny_rows(bool TA, bool TB, bool ONES, const GemmP& g, int i_off) {
  SkinnyP s{};
  s.A = g.A; s.B = g.B; s.C = g.C; s.bias = g.bias; s.addend = g.addend; s.aux = g.aux;
  s.Ms = g.M; s.Nl = g.N; s.K = g.K; s.i_off = i_off;
  s.a_si = TA ? 1 : g.lda; s.a_sk = TA ? g.lda : 1;
  s.b_sk = TB ? 1 : g.ldb; s.b_sj = TB ? g.ldb : 1;
  s.c_si = g.ldc; s.c_sj = 1;
  s.sAa = g.sAa; s.sAn = g.sAn; s.sBa = g.sBa; s.sBn = g.sBn; s.sCa = g.sCa; s.sCn = g.sCn; s.sba = g.sba; s.sbn = g.sbn;
  s.nnet = g.nnet; s.epi = g.epi; s.act = g.act; s.bias_on_i = 0;
  s.ones_row = ONES ? g.M - 1 : -1;
  s.ones_col = -1;
  return s;
}
// ... and the role-swapped problem when the N side is short: C^T(n, m) = sum_k B^T(n,k) A^T(k,m)
static inline SkinnyP skinny_cols(bool TA, bool TB, bool ONES, const GemmP& g) {
  SkinnyP s{};
  s.A = g.B; s.B = g.A; s.C = g.C; s.bias = g.bias; s.addend = g.addend; s.aux = g.aux;
  s.Ms = g.N; s.Nl = g.M; s.K = g.K; s.i_off = 0;
  s.a_si = TB ? g.ldb : 1; s.a_sk = TB ? 1 : g.ldb;          // short operand = op(B)^T : (n,k)
  s.b_sk = TA ? g.lda : 1; s.b_sj = TA ? 1 : g.lda;          // long operand  = op(A)^T : (k,m)
  s.c_si = 1; s.c_sj = g.ldc;
  s.sAa = g.sBa; s.sAn = g.sBn; s.sBa = g.sAa; s.sBn = g.sAn; s.sCa = g.sCa; s.sCn = g.sCn; s.sba = g.sba; s.sbn = g.sbn;
  s.nnet = g.nnet; s.epi = g.epi; s.act = g.act; s.bias_on_i = 1;
  s.ones_row = -1;
  s.ones_col = ONES ? g.M - 1 : -1;
  return s;
}
static inline void skinny_launch(const SkinnyP& s, int nagents, cudaStream_t st) {
  const int ms = s.Ms - s.i_off;
  dim3 grid((s.Nl + SK_THREADS - 1) / SK_THREADS, 1, nagents * s.nnet), block(SK_THREADS);
  const bool bt = (s.b_sk == 1 && s.b_sj != 1);
#define SK_GO(MSV) do { if (bt) k_gemm_skinny<MSV, true><<<grid, block, 0, st>>>(s); \
                        else k_gemm_skinny<MSV, false><<<grid, block, 0, st>>>(s); } while (0)
  if (ms <= 1) SK_GO(1);
  else if (ms <= 4) SK_GO(4);
  else if (ms <= 8) SK_GO(8);
  else if (ms <= 16) SK_GO(16);
  else SK_GO(32);
#undef SK_GO
}

}  // namespace saceo

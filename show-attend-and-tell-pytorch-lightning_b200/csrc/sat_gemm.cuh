// GEMM dispatch: tcgen05 tensor-core core (bf16 operands, shapes that meet the TMA/UMMA tile
// constraints) or the SIMT FFMA core (fp32 parity mode and odd shapes).  Both run the same
// epilogue functors, so every fused stage exists in both precisions.
#pragma once
#include <type_traits>

#include "sat_gemm_simt.cuh"
#include "sat_gemm_tc.cuh"

constexpr int SAT_MAX_SPLITK = 16;

template <typename TA, typename TW>
static bool gemm_tn_uses_tc(bool use_tc, const GemmOperandA& A, const TW* W, int64_t ldw, int M, int N) {
  if constexpr (std::is_same<TA, bf16>::value && std::is_same<TW, bf16>::value)
    return use_tc && M > 0 && N > 0 && N % 4 == 0 && tc::operands_ok(A, W, ldw);
  return false;
}

// splitk > 1 (tensor-core core only): CTA z of the grid accumulates a slice of K and the epilogue (an EpiStore with
// zstride set) writes partial z; the consumer kernel adds the partials in fixed order.
template <typename TA, typename TW, typename Epi>
static int gemm_tn(bool use_tc, const GemmOperandA& A, const TW* W, int64_t ldw, int M, int N, const Epi& epi,
                   cudaStream_t stream, int splitk = 1) {
  if constexpr (std::is_same<TA, bf16>::value && std::is_same<TW, bf16>::value) {
    if (gemm_tn_uses_tc<TA, TW>(use_tc, A, W, ldw, M, N)) return tc::launch<Epi>(A, W, ldw, M, N, epi, stream, splitk);
  }
  SAT_REQUIRE(splitk == 1, "split-K is only available on the tensor-core GEMM core");
  return launch_gemm_tn_simt<TA, TW, Epi>(A, W, ldw, M, N, epi, stream);
}

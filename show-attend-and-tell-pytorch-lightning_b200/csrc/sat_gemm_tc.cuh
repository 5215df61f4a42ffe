// tcgen05 (5th-gen tensor core) GEMM core for sm_100a:  C[M,N] = epi( sum_seg A_seg[M,K_seg] * W[N,K]^T ),
// bf16 operands (both K-major), fp32 accumulation in tensor memory.
//
//   warp 0    : TMA producer  (cp.async.bulk.tensor.2d, 128B-swizzled tiles, 4-stage mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16)
//   warps 2-9 : epilogue: tcgen05.ld (32 lanes x 16 columns per instruction) -> shared-memory transpose -> fused
//               epilogue functor with coalesced global accesses (two warps per TMEM lane quarter, each half the columns)
//
// The epilogue functors are the same ones the SIMT core runs (LSTM cell, tanh+add, CE, plain store ...), so
// every fused stage of the decoder exists on both cores.  Tiles: BM = 128, BN in {64,128}, BK = 64.
// M / N / K tails are handled by TMA out-of-bounds zero fill plus predicated epilogue stores.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <type_traits>

#include "sat_gemm_simt.cuh"

namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quarter: the fused epilogues are issue/latency-bound
constexpr int THREADS = 64 + EPI_WARPS * 32;
// operand ring depth: 4 x 24 KB at BN = 64 and 3 x 32 KB at BN = 128 -> 96 KB per CTA, two CTAs per SM, so one CTA's
// epilogue overlaps the other's TMA/MMA main loop (the kernels are not persistent)
template <int BN> struct Stages { static constexpr int value = BN >= 128 ? 3 : 4; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 format: version 1 at bit 46,
// stride-byte-offset = 8 rows * 128 B, layout type 2 = SWIZZLE_128B at bits 61..63).
__device__ __forceinline__ uint64_t make_smem_desc(const void* p) {
  const uint32_t a = smem_u32(p);
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
template <int BN> __device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Maps {
  CUtensorMap a[2];   // A segments: dims {K_s, M}, box {64, 128}
  CUtensorMap w;      // W: dims {Ktot, N}, box {64, BN}
};

// Epilogue functors come in two kinds.  The default kind is called per (row, 4 columns) after the accumulator tile has
// been transposed through shared memory (coalesced global accesses).  A functor that declares
//   static constexpr bool kRowOwner = true;
// takes over the whole epilogue of a tile instead (run<BN>()): each thread keeps the accumulator ROW that tcgen05.ld
// hands it, which makes per-row reductions over the tile's columns (soft-max statistics, arg-max) thread-local.
template <typename T, typename = void> struct is_row_owner : std::false_type {};
template <typename T> struct is_row_owner<T, std::void_t<decltype(T::kRowOwner)>> : std::bool_constant<T::kRowOwner> {};
constexpr int AUX_FLOATS = 160;      // small per-CTA scratch that does not alias the operand ring (bias tile of a row-owner epilogue)

template <int BN>
struct Smem {
  static constexpr int STAGES = Stages<BN>::value;
  alignas(1024) bf16 a[STAGES][BM * BK];
  alignas(1024) bf16 w[STAGES][BN * BK];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
  alignas(16) float aux[AUX_FLOATS];
};

// ===== standard epilogue: TMEM -> registers -> shared (transpose) -> coalesced fused epilogue =====
// A thread owns one accumulator ROW after tcgen05.ld; global accesses of the epilogue functors want a warp on
// consecutive COLUMNS of one row.  The operand ring is idle once tmem_full fires, so it is reused as a [128][BN+4] fp32
// staging tile (conflict-free: row stride = 4 mod 32 banks).  Called by the 8 epilogue warps (warp = 2..9).
template <int BN, typename Epi>
__device__ __forceinline__ void epilogue_tile(float* tile, uint32_t tmem, bool has_acc, int warp, int lane, int m0, int n0, int M, int N,
                                              const Epi& epi) {
  const int q = warp & 3;                    // TMEM lane quarter this warp may access
  constexpr int LDT = BN + 4;
  const int row = q * 32 + lane;
  const int half = (warp - 2) >> 2;          // which half of the accumulator columns this warp moves
  constexpr int CH = BN / 2;
  if (has_acc) {
#pragma unroll 1
    for (int c = half * CH; c < (half + 1) * CH; c += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(&tile[row * LDT + c + g * 4]) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
  } else {
    for (int c = half * CH; c < (half + 1) * CH; c += 4) *reinterpret_cast<float4*>(&tile[row * LDT + c]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
  constexpr int LPR = BN / 4;                // lanes per row
  constexpr int RPW = 32 / LPR;              // rows per warp pass
  const int ew = warp - 2;                   // 0..EPI_WARPS-1
  const int lr = lane / LPR, lc = (lane % LPR) * 4;
  const int n = n0 + lc;
  constexpr int NIT = BM / (EPI_WARPS * RPW);   // rows per thread
  constexpr int UN = 4;                      // rows whose epilogue operands are loaded before any dependent math
  static_assert(NIT % UN == 0, "row loop must divide");
#pragma unroll 1
  for (int i0 = 0; i0 < NIT; i0 += UN) {
    typename Epi::Ctx ctx[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int m = m0 + ew * RPW + lr + (i0 + u) * EPI_WARPS * RPW;
      ok[u] = m < M && n < N;
      if (ok[u]) ctx[u] = epi.load(m, n);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int r = ew * RPW + lr + (i0 + u) * EPI_WARPS * RPW;
      if (ok[u]) {
        const float4 a = *reinterpret_cast<const float4*>(&tile[r * LDT + lc]);
        const float a4[4] = {a.x, a.y, a.z, a.w};
        epi.apply(m0 + r, n, a4, ctx[u]);
      }
    }
  }
}

template <int BN, typename Epi>
__global__ void __launch_bounds__(THREADS)
gemm_tn_tc_kernel(const __grid_constant__ Maps maps, int M, int N, int k0, int k1, Epi epi) {
  // gridDim.z = split-K factor: CTA z accumulates k blocks [z*nkb/S, (z+1)*nkb/S) (partials are summed by the consumer)
  extern __shared__ uint8_t smem_raw[];
  constexpr int STAGES = Stages<BN>::value;
  Smem<BN>& s = *reinterpret_cast<Smem<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb0 = (k0 + BK - 1) / BK, nkb1 = (k1 + BK - 1) / BK;
  const int nkb_all = nkb0 + nkb1;
  const int kb_begin = (int)(((long)nkb_all * blockIdx.z) / gridDim.z);
  const int kb_end = (int)(((long)nkb_all * (blockIdx.z + 1)) / gridDim.z);
  const int nkb = kb_end - kb_begin;
  constexpr uint32_t STAGE_BYTES = (BM * BK + BN * BK) * sizeof(bf16);

  // Programmatic dependent launch: let the next kernel of the stream start its own prologue now; this kernel's prologue
  // (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of its predecessor and only then waits for it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // Inside the decoder drivers W is always a packed weight matrix, never written within the launch chain: thread 0 (which
  // just initialised the barriers) starts the weight tiles of the first ring pass before griddepcontrol.wait; the
  // activation tiles follow it.  (sat_linear, with a caller-supplied W, launches without the PDL attribute.)
  const int npre = nkb < STAGES ? nkb : STAGES;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
    for (int i = 0; i < npre; ++i) {
      const int kb = kb_begin + i;
      const int kw = kb >= nkb0 ? k0 + (kb - nkb0) * BK : kb * BK;
      mbar_expect_tx(&s.full[i], STAGE_BYTES);
      tma_load_2d(&maps.w, &s.full[i], s.w[i], kw, n0);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // predecessor grid complete and its writes visible

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb_begin + i;
        const int st = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const bool seg1 = kb >= nkb0;
        const int ka = (seg1 ? kb - nkb0 : kb) * BK;            // k offset inside the A segment
        const int kw = seg1 ? k0 + ka : ka;                      // k offset inside W
        if (i >= npre) {                                         // first pass: tx count armed and W already in flight
          mbar_wait(&s.empty[st], ph ^ 1);
          mbar_expect_tx(&s.full[st], STAGE_BYTES);
          tma_load_2d(&maps.w, &s.full[st], s.w[st], kw, n0);
        }
        tma_load_2d(seg1 ? &maps.a[1] : &maps.a[0], &s.full[st], s.a[st], ka, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      for (int kb = 0; kb < nkb; ++kb) {      // local k-block index (split-K: this CTA's share)
        const int st = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&s.full[st], ph);
        tcgen05_fence_after();
        const uint64_t ad = make_smem_desc(s.a[st]), bd = make_smem_desc(s.w[st]);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) umma(tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
        umma_commit(&s.empty[st]);            // frees the smem stage when these MMAs retire
      }
      umma_commit(&s.tmem_full);              // accumulator complete
    }
  } else {
    if constexpr (is_row_owner<Epi>::value) {
      // row-owner epilogue (soft-max statistics / arg-max of the vocabulary projection): the functor stages what it needs
      // in s.aux before the accumulator is complete, then reads TMEM itself
      epi.prologue(s.aux, n0, N, warp, lane);
      mbar_wait(&s.tmem_full, 0);
      tcgen05_fence_after();
      static_assert((size_t)BM * (BN / 2 + 8) * sizeof(float) <= sizeof(s.a) + sizeof(s.w), "row-owner scratch must fit the ring");
      epi.template run<BN, EPI_WARPS>(reinterpret_cast<uint8_t*>(&s.a[0][0]), s.aux, tmem, nkb > 0, warp, lane, m0, n0, M, N, (int)blockIdx.x);
    } else {
      mbar_wait(&s.tmem_full, 0);
      tcgen05_fence_after();
      static_assert((size_t)BM * (BN + 4) * sizeof(float) <= sizeof(s.a) + sizeof(s.w), "staging tile must fit the ring");
      epilogue_tile<BN, Epi>(reinterpret_cast<float*>(&s.a[0][0]), tmem, nkb > 0, warp, lane, m0, n0, M, N, epi);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
  }
}

// =============================================================================================
// Persistent variant for row-owner epilogues (the vocabulary projection: M = T*B rows, N = V, only K = E <= 256 deep).  One
// CTA per SM owns a CONTIGUOUS run of output tiles in m-major order (n fastest), so that
//   * the A row block of the run (128 rows x K, <= 64 KB) is loaded ONCE and stays resident in shared memory while the W
//     tiles stream through a 6-stage ring: half the L2 -> SM traffic of reloading both operands per tile (this GEMM is
//     bound by that traffic, not by the tensor pipe: 2000 tiles x 128 KB per pass at BASELINE configs[1]);
//   * TWO accumulator buffers in tensor memory let the MMA warp fill one while the 16 epilogue warps drain the other;
//   * launch, barrier-init and TMEM-allocation latency is paid once per CTA instead of once per 128x128 tile.
// =============================================================================================
constexpr int PERS_WSTAGES = 6;
constexpr int PERS_AKB = 4;                   // resident A: up to 4 k-blocks of 64 (K <= 256)
constexpr int PERS_EPI_WARPS = 16;            // four per scheduler: the row-owner epilogues are chains of dependent ALU / MUFU ops
constexpr int PERS_THREADS = 64 + PERS_EPI_WARPS * 32;
struct SmemPers {
  alignas(1024) bf16 a[PERS_AKB][BM * BK];
  alignas(1024) bf16 w[PERS_WSTAGES][128 * BK];
  alignas(1024) uint8_t scratch[BM * (128 / 2 + 8) * sizeof(float)];   // row-owner epilogue scratch (not aliased with the ring here): 36 KB
  alignas(8) uint64_t full[PERS_WSTAGES];
  uint64_t empty[PERS_WSTAGES];
  uint64_t a_full, a_empty;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  alignas(16) float aux[AUX_FLOATS];
};

template <typename Epi>
__global__ void __launch_bounds__(PERS_THREADS, 1)
gemm_tn_tc_persistent_kernel(const __grid_constant__ Maps maps, int M, int N, int K, Epi epi) {
  static_assert(is_row_owner<Epi>::value, "the persistent kernel runs row-owner epilogues");
  constexpr int BN = 128;
  extern __shared__ uint8_t smem_raw[];
  SmemPers& s = *reinterpret_cast<SmemPers*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BK - 1) / BK;                                  // <= PERS_AKB (checked by the launcher)
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + BM - 1) / BM;
  const int ntiles = tiles_n * tiles_m;
  const int t0 = (int)(((long)ntiles * blockIdx.x) / gridDim.x), t1 = (int)(((long)ntiles * (blockIdx.x + 1)) / gridDim.x);
  constexpr uint32_t W_BYTES = BN * BK * sizeof(bf16), A_KB_BYTES = BM * BK * sizeof(bf16);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < PERS_WSTAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.a_full, 1);
    mbar_init(&s.a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], PERS_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
  }
  if (warp == 1) {   // two accumulator buffers of BN fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, aload = 0;                        // W k-block counter across tiles, number of A loads so far
      int cur_mb = -1;
      for (int tile = t0; tile < t1; ++tile) {
        const int mb = tile / tiles_n, n0 = (tile - mb * tiles_n) * BN;
        if (mb != cur_mb) {                               // new row block: (re)load the resident A operand
          mbar_wait(&s.a_empty, (aload & 1) ^ 1);         // every MMA that read the previous block has retired
          mbar_expect_tx(&s.a_full, A_KB_BYTES * (uint32_t)nkb);
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&maps.a[0], &s.a_full, s.a[kb], kb * BK, mb * BM);
          ++aload;
          cur_mb = mb;
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % PERS_WSTAGES;
          const uint32_t ph = (it / PERS_WSTAGES) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          mbar_expect_tx(&s.full[st], W_BYTES);
          tma_load_2d(&maps.w, &s.full[st], s.w[st], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      uint32_t it = 0, lt = 0, aload = 0;
      int cur_mb = -1;
      for (int tile = t0; tile < t1; ++tile, ++lt) {
        const int mb = tile / tiles_n;
        if (mb != cur_mb) {
          mbar_wait(&s.a_full, aload & 1);
          ++aload;
          cur_mb = mb;
        }
        const uint32_t buf = lt & 1;
        mbar_wait(&s.tmem_empty[buf], ((lt >> 1) & 1) ^ 1);       // the epilogue has drained this buffer (free at first use)
        tcgen05_fence_after();
        const uint32_t acc = tmem + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % PERS_WSTAGES;
          const uint32_t ph = (it / PERS_WSTAGES) & 1;
          mbar_wait(&s.full[st], ph);
          tcgen05_fence_after();
          const uint64_t ad = make_smem_desc(s.a[kb]), bd = make_smem_desc(s.w[st]);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) umma(acc, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
          umma_commit(&s.empty[st]);
        }
        umma_commit(&s.tmem_full[buf]);
        if (tile + 1 >= t1 || (tile + 1) / tiles_n != mb) umma_commit(&s.a_empty);   // last tile of this row block
      }
    }
  } else {
    uint32_t lt = 0;
    for (int tile = t0; tile < t1; ++tile, ++lt) {
      const int mb = tile / tiles_n, nt = tile - mb * tiles_n;
      const int m0 = mb * BM, n0 = nt * BN;
      const uint32_t buf = lt & 1;
      epi.prologue(s.aux, n0, N, warp, lane);
      mbar_wait(&s.tmem_full[buf], (lt >> 1) & 1);
      tcgen05_fence_after();
      epi.template run<BN, PERS_EPI_WARPS>(s.scratch, s.aux, tmem + buf * BN, nkb > 0, warp, lane, m0, n0, M, N, nt);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&s.tmem_empty[buf]);          // this warp no longer reads the buffer
      asm volatile("bar.sync 1, %0;" ::"n"(PERS_EPI_WARPS * 32) : "memory");   // scratch / aux are reused by the next tile
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN));
  }
}

// =============================================================================================
// "NT" core: C[N1,N2] = sum_k A[k,n1] * B[k,n2]  -- the weight-gradient GEMMs dW = dY^T X, whose contraction runs over the
// T*B rows of two row-major activation buffers.  Both operands are therefore MN-major for the tensor core: a TMA box of
// {64 columns, BK rows} lands as BK 128-byte rows (128B swizzle), which is exactly the canonical MN-major SWIZZLE_128B
// layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with SBO = 1024 B (eight k rows) and LBO = BK*128 B (the next
// 64-column box); one tcgen05.mma consumes 16 k rows = two 8-row groups, so the descriptor advances by 2048 B per MMA.
// gridDim.z splits the k range; partials are written per z (EpiStore::zstride) and summed in fixed order by the
// consumer (param_grads_finalize_kernel), which keeps the gradients bit-reproducible.
// =============================================================================================
struct MapsNT {
  CUtensorMap a;      // A: dims {N1, Krows}, box {64, BK}
  CUtensorMap b;      // B: dims {N2, Krows}, box {64, BK}
};

template <int BN>
struct SmemNT {
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  alignas(1024) bf16 a[STAGES][2 * BK * 64];              // two boxes: columns n1 0..63 | 64..127, each [BK rows][64]
  alignas(1024) bf16 b[STAGES][(BN / 64) * BK * 64];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

// MN-major, 128B-swizzled shared-memory matrix descriptor (see above)
__device__ __forceinline__ uint64_t make_smem_desc_mn(const void* p, uint32_t lbo_bytes) {
  const uint32_t a = smem_u32(p);
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int BN, typename Epi>
__global__ void __launch_bounds__(THREADS)
gemm_nt_tc_kernel(const __grid_constant__ MapsNT maps, int N1, int N2, int Krows, Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  static_assert(BN == 64 || BN == 128, "MN-major B tiles are made of 64-column boxes");
  constexpr int STAGES = SmemNT<BN>::STAGES;
  SmemNT<BN>& s = *reinterpret_cast<SmemNT<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_all = (Krows + BK - 1) / BK;
  const int kb_begin = (int)(((long)nkb_all * blockIdx.z) / gridDim.z);
  const int kb_end = (int)(((long)nkb_all * (blockIdx.z + 1)) / gridDim.z);
  const int nkb = kb_end - kb_begin;
  constexpr uint32_t BOX_BYTES = BK * 64 * sizeof(bf16);
  constexpr uint32_t STAGE_BYTES = (2 + BN / 64) * BOX_BYTES;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // both operands are activations written by predecessors

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int st = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const int krow = (kb_begin + i) * BK;
        mbar_wait(&s.empty[st], ph ^ 1);
        mbar_expect_tx(&s.full[st], STAGE_BYTES);
        tma_load_2d(&maps.a, &s.full[st], &s.a[st][0], m0, krow);
        tma_load_2d(&maps.a, &s.full[st], &s.a[st][BK * 64], m0 + 64, krow);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(&maps.b, &s.full[st], &s.b[st][j * BK * 64], n0 + 64 * j, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>() | (1u << 15) | (1u << 16);      // A and B are MN-major
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&s.full[st], ph);
        tcgen05_fence_after();
        const uint64_t ad = make_smem_desc_mn(&s.a[st][0], BOX_BYTES), bd = make_smem_desc_mn(&s.b[st][0], BOX_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk)      // 16 k rows = 2048 B = 128 descriptor units per MMA
          umma(tmem, ad + 128 * kk, bd + 128 * kk, idesc, (kb | kk) != 0);
        umma_commit(&s.empty[st]);
      }
      umma_commit(&s.tmem_full);
    }
  } else {
    mbar_wait(&s.tmem_full, 0);
    tcgen05_fence_after();
    static_assert((size_t)BM * (BN + 4) * sizeof(float) <= sizeof(s.a) + sizeof(s.b), "staging tile must fit the ring");
    epilogue_tile<BN, Epi>(reinterpret_cast<float*>(&s.a[0][0]), tmem, nkb > 0, warp, lane, m0, n0, N1, N2, epi);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
  }
}

// ---- host side ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// 2D bf16 tensor map: inner dim = K (contiguous), outer = rows; box {64, box_rows}; 128B swizzle; OOB -> 0
static int make_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
  auto enc = get_encode();
  SAT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld k=%lld ld=%lld", (int)r, (long long)rows,
              (long long)k, (long long)ld);
  return 0;
}

// generic 2D bf16 tensor map of a row-major [rows, cols] buffer: box {box_cols, box_rows}, no swizzle, OOB -> 0
static int make_map_plain(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  auto enc = get_encode();
  SAT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (plain) failed (%d) rows=%lld cols=%lld box=%dx%d", (int)r, (long long)rows,
              (long long)cols, box_cols, box_rows);
  return 0;
}

static inline bool operands_ok(const GemmOperandA& A, const void* W, int64_t ldw) {
  if (A.nseg < 1 || A.nseg > 2) return false;
  for (int i = 0; i < A.nseg; ++i) {
    if (A.k[i] % 8 != 0 || A.ld[i] % 8 != 0 || (reinterpret_cast<uintptr_t>(A.p[i]) & 15)) return false;
  }
  // every segment but the last must fill whole 64-wide k blocks (W is addressed with a running k offset)
  if (A.nseg == 2 && A.k[0] % BK != 0) return false;
  return ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
}

template <int BN, typename Epi>
static int launch_bn(const GemmOperandA& A, const bf16* W, int64_t ldw, int M, int N, const Epi& epi, cudaStream_t stream,
                     int splitk = 1) {
  Maps maps;
  int ktot = 0;
  for (int i = 0; i < A.nseg; ++i) {
    SAT_TRY(make_map(&maps.a[i], A.p[i], M, A.k[i], A.ld[i], BM));
    ktot += A.k[i];
  }
  if (A.nseg == 1) maps.a[1] = maps.a[0];
  SAT_TRY(make_map(&maps.w, W, N, ktot, ldw, BN));
  auto kern = gemm_tn_tc_kernel<BN, Epi>;
  constexpr int smem = (int)sizeof(Smem<BN>) + 1024;
  static bool attr_set[64] = {false};         // the attribute is per device
  int dev_now = 0;
  SAT_CUDA(cudaGetDevice(&dev_now));
  if (!attr_set[dev_now & 63]) {
    SAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev_now & 63] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((N + BN - 1) / BN, (M + BM - 1) / BM, splitk);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = sat_pdl_allowed() ? 1 : 0;
  SAT_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, M, N, A.k[0], A.nseg == 2 ? A.k[1] : 0, epi));
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

template <typename Epi>
static int launch(const GemmOperandA& A, const bf16* W, int64_t ldw, int M, int N, const Epi& epi, cudaStream_t stream,
                  int splitk = 1) {
  const long mt = (M + BM - 1) / BM;
  const long tiles128 = (long)((N + 127) / 128) * mt;
  if (tiles128 >= 148 && N >= 128 && splitk == 1) return launch_bn<128, Epi>(A, W, ldw, M, N, epi, stream);
  // skinny GEMMs (the per-step M = B projections): narrower N tiles put more CTAs to work and halve the serial
  // epilogue work per thread (the fused LSTM / store epilogues are latency-bound, not throughput-bound)
  const long tiles64 = (long)((N + 63) / 64) * mt * splitk;
  if (tiles64 < 120 && N >= 64) return launch_bn<32, Epi>(A, W, ldw, M, N, epi, stream, splitk);
  return launch_bn<64, Epi>(A, W, ldw, M, N, epi, stream, splitk);
}

// split-K factor for skinny GEMMs (few output tiles, long K): enough CTAs to cover the SMs, >= 4 k blocks each
static inline int pick_splitk(int M, int N, int ktot) {
  const long tiles = (long)((N + 63) / 64) * ((M + BM - 1) / BM);
  const int nkb = (ktot + BK - 1) / BK;
  int s = 1;
  while (s < 16 && tiles * (s * 2) <= 160 && nkb / (s * 2) >= 4) s *= 2;
  return s;
}


// persistent launcher (row-owner epilogues, single A segment of K <= 256, BN = 128)
static inline bool persistent_ok(const GemmOperandA& A) { return A.nseg == 1 && A.k[0] <= PERS_AKB * BK; }

template <typename Epi>
static int launch_persistent(const GemmOperandA& A, const bf16* W, int64_t ldw, int M, int N, const Epi& epi, cudaStream_t stream) {
  Maps maps;
  SAT_TRY(make_map(&maps.a[0], A.p[0], M, A.k[0], A.ld[0], BM));
  maps.a[1] = maps.a[0];
  SAT_TRY(make_map(&maps.w, W, N, A.k[0], ldw, 128));
  auto kern = gemm_tn_tc_persistent_kernel<Epi>;
  constexpr int smem = (int)sizeof(SmemPers) + 1024;
  static bool attr_set[64] = {false};
  static int n_sm[64] = {0};
  int dev_now = 0;
  SAT_CUDA(cudaGetDevice(&dev_now));
  if (!attr_set[dev_now & 63]) {
    SAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SAT_CUDA(cudaDeviceGetAttribute(&n_sm[dev_now & 63], cudaDevAttrMultiProcessorCount, dev_now));
    attr_set[dev_now & 63] = true;
  }
  const int ntiles = ((N + 127) / 128) * ((M + BM - 1) / BM);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ntiles < n_sm[dev_now & 63] ? ntiles : n_sm[dev_now & 63]);
  cfg.blockDim = dim3(PERS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = sat_pdl_allowed() ? 1 : 0;
  SAT_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, M, N, A.k[0], epi));
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

// ---- NT launchers ------------------------------------------------------------------------------------
// 2D bf16 tensor map of a row-major [rows, cols] buffer read as {64 columns, BK rows} boxes (MN-major operand tiles)
static int make_map_mn(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  auto enc = get_encode();
  SAT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {64u, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (MN-major) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
  return 0;
}

static inline bool nt_operands_ok(const void* A, int64_t lda, const void* B, int64_t ldb) {
  return lda % 8 == 0 && ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
}

template <int BN, typename Epi>
static int launch_nt_bn(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, int Krows, int N1, int N2, const Epi& epi,
                        cudaStream_t stream, int splitk) {
  MapsNT maps;
  SAT_TRY(make_map_mn(&maps.a, A, Krows, N1, lda));
  SAT_TRY(make_map_mn(&maps.b, B, Krows, N2, ldb));
  auto kern = gemm_nt_tc_kernel<BN, Epi>;
  constexpr int smem = (int)sizeof(SmemNT<BN>) + 1024;
  static bool attr_set[64] = {false};
  int dev_now = 0;
  SAT_CUDA(cudaGetDevice(&dev_now));
  if (!attr_set[dev_now & 63]) {
    SAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev_now & 63] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((N2 + BN - 1) / BN, (N1 + BM - 1) / BM, splitk);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = sat_pdl_allowed() ? 1 : 0;
  SAT_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, N1, N2, Krows, epi));
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

template <typename Epi>
static int launch_nt(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, int Krows, int N1, int N2, const Epi& epi,
                     cudaStream_t stream, int splitk) {
  if (N2 > 64) return launch_nt_bn<128, Epi>(A, lda, B, ldb, Krows, N1, N2, epi, stream, splitk);
  return launch_nt_bn<64, Epi>(A, lda, B, ldb, Krows, N1, N2, epi, stream, splitk);
}

// split-K factor of a weight-gradient GEMM: enough CTAs for ~2 per SM, at least 4 k blocks each
static inline int pick_splitk_nt(int N1, int N2, int Krows) {
  const long tiles = (long)((N1 + BM - 1) / BM) * ((N2 + (N2 > 64 ? 127 : 63)) / (N2 > 64 ? 128 : 64));
  const int nkb = (Krows + BK - 1) / BK;
  int s = 1;
  while (s < 32 && tiles * (s * 2) <= 320 && nkb / (s * 2) >= 4) s *= 2;
  return s;
}

}  // namespace tc

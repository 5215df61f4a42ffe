// K1 v3: fused attention step (forward) as a TMA-fed shared-memory pipeline, one CTA per caption row.
//
//   warp 8 (one elected lane) is the producer: it streams the caption's P tile (L x A) and then its
//   annotation tile (L x D) through a ring of NST 16-KB stages with 1-D TMA bulk copies
//   (cp.async.bulk + mbarrier complete_tx), running up to NST stages (96 KB) ahead of the consumers, so
//   the annotation rows are already in flight while the scores and the softmax are being computed.
//   warps 0-7 are consumers: scores from the P stages, softmax over L, alpha-weighted context from the
//   annotation stages (16-byte shared-memory loads), beta gate, stores.
//   Every byte of P / annotations is read from global memory exactly once per step; 2 CTAs per SM keep
//   ~190 KB of loads in flight per SM, which is what an HBM-bound kernel needs on B200.
#pragma once
#include <stdlib.h>

#include "sat_kernels.cuh"

__device__ __forceinline__ uint32_t sat_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sat_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sat_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sat_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sat_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sat_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sat_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sat_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = sat_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"   // suspend-time hint: sleep, don't spin
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D TMA bulk copy global -> shared (size a multiple of 16 B, both addresses 16 B aligned)
__device__ __forceinline__ void sat_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sat_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(sat_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sat_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int ATTP_CWARPS = 8;                       // consumer warps
constexpr int ATTP_CONSUMERS = ATTP_CWARPS * 32;
constexpr int ATTP_THREADS = ATTP_CONSUMERS + 32;    // + producer warp
constexpr int ATTP_NST = 6;
constexpr int ATTP_STAGE_BYTES = 16384;
constexpr int ATTP_KA = 2;                           // attention_dim <= 256 on the pipelined kernel

struct AttPipeSmem {
  uint64_t full[ATTP_NST];
  uint64_t empty[ATTP_NST];
  float red_a[ATTP_CWARPS];
  float red_b[ATTP_CWARPS];
};

template <typename T, bool kExact>
__global__ void __launch_bounds__(ATTP_THREADS)
attention_step_fwd_pipe_kernel(const T* __restrict__ ann, const T* __restrict__ P, const float* __restrict__ wf,
                               const float* __restrict__ hp, int64_t ldhp, const int32_t* __restrict__ lens, int t, int ncap,
                               int L, int D, int A, float scale, float* __restrict__ alpha, int64_t ld_alpha,
                               float* __restrict__ qsave, T* __restrict__ z, T* __restrict__ gz, T* __restrict__ beta,
                               int64_t ld_z) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int VN = Vec16<T>::N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  AttPipeSmem* hd = reinterpret_cast<AttPipeSmem*>(smem_raw);
  float* e = reinterpret_cast<float*>(smem_raw + sizeof(AttPipeSmem));   // [L4]
  const int L4 = (L + 3) & ~3;
  float* qs = e + L4;                   // [A]
  float* ws = qs + A;                   // [A]
  const int NV = D / VN;
  const int RG = NV >= ATTP_CONSUMERS ? 1 : ATTP_CONSUMERS / NV;
  float* red = ws + A;                  // [RG * D]
  // stage ring: offset computed with integer arithmetic from the (128-byte aligned) dynamic smem base so that the
  // compiler keeps the shared address space (LDS instead of generic LD)
  const uint32_t stage_off = (uint32_t)((sizeof(AttPipeSmem) + sizeof(float) * (size_t)(L4 + 2 * A + RG * D) + 127) & ~(size_t)127);
  uint8_t* stages = smem_raw + stage_off;

  const bool active = lens == nullptr || t < lens[b];
  float* alpha_b = alpha + (int64_t)b * ld_alpha;
  if (!active) {
    for (int l = tid; l < L; l += ATTP_THREADS) alpha_b[l] = 0.0f;
    for (int d = tid; d < D; d += ATTP_THREADS) {
      z[(int64_t)b * ld_z + d] = from_f<T>(0.f);
      gz[(int64_t)b * ld_z + d] = from_f<T>(0.f);
      if (beta) beta[(int64_t)b * ld_z + d] = from_f<T>(0.f);
    }
    if (qsave) for (int a = tid; a < A; a += ATTP_THREADS) qsave[(int64_t)b * A + a] = 0.0f;
    return;
  }
  const int img = b / ncap;
  const int RCP = ATTP_STAGE_BYTES / (A * (int)sizeof(T));      // P rows per stage
  const int RCA = ATTP_STAGE_BYTES / (D * (int)sizeof(T));      // annotation rows per stage
  const int nP = (L + RCP - 1) / RCP, nA = (L + RCA - 1) / RCA;

  if (tid == 0) {
    for (int i = 0; i < ATTP_NST; ++i) {
      sat_mbar_init(&hd->full[i], 1);
      sat_mbar_init(&hd->empty[i], ATTP_CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const float* hp_b = hp + (int64_t)b * ldhp;
  for (int a = tid; a < A; a += ATTP_THREADS) {
    const float q = hp_b[a];
    qs[a] = q;
    ws[a] = wf[a];
    if (qsave) qsave[(int64_t)b * A + a] = q;
  }
  __syncthreads();

  if (warp == ATTP_CWARPS) {
    // ===== producer =====
    if (lane == 0) {
      const T* Pb = P + (int64_t)img * L * A;
      const T* ab = ann + (int64_t)img * L * D;
      for (int i = 0; i < nP + nA; ++i) {
        const int st = i % ATTP_NST;
        const uint32_t ph = (uint32_t)(i / ATTP_NST) & 1u;
        sat_mbar_wait(&hd->empty[st], ph ^ 1u);
        const void* src;
        uint32_t bytes;
        if (i < nP) {
          const int r0 = i * RCP, rows = min(RCP, L - r0);
          src = Pb + (int64_t)r0 * A;
          bytes = (uint32_t)(rows * A * (int)sizeof(T));
        } else {
          const int r0 = (i - nP) * RCA, rows = min(RCA, L - r0);
          src = ab + (int64_t)r0 * D;
          bytes = (uint32_t)(rows * D * (int)sizeof(T));
        }
        sat_mbar_expect_tx(&hd->full[st], bytes);
        sat_bulk_g2s(stages + (size_t)st * ATTP_STAGE_BYTES, src, bytes, &hd->full[st]);
      }
    }
    return;
  }

  // ===== consumers (256 threads, named barrier 1) =====
  // lane-resident slices of q and w_f: lane owns attention columns lane*4 + 128*k .. +3  (A <= 128 * ATTP_KA)
  float qreg[ATTP_KA][4], wreg[ATTP_KA][4];
#pragma unroll
  for (int k = 0; k < ATTP_KA; ++k) {
    const int a = lane * 4 + 128 * k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qreg[k][i] = a < A ? qs[a + i] : 0.0f;
      wreg[k][i] = a < A ? ws[a + i] : 0.0f;
    }
  }
  int it = 0;   // running chunk index (same sequence as the producer)
  for (int i = 0; i < nP; ++i, ++it) {
    const int st = it % ATTP_NST;
    sat_mbar_wait(&hd->full[st], (uint32_t)(it / ATTP_NST) & 1u);
    const T* Ps = reinterpret_cast<const T*>(stages + (size_t)st * ATTP_STAGE_BYTES);
    const int r0 = i * RCP, rows = min(RCP, L - r0);
    for (int l = warp; l < rows; l += ATTP_CWARPS) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < ATTP_KA; ++k) {
        const int a = lane * 4 + 128 * k;
        if (a < A) {
          const float4 p = ld4(Ps + (size_t)l * A + a);
          s = fmaf(wreg[k][0], sat_tanh<kExact>(p.x + qreg[k][0]), s);
          s = fmaf(wreg[k][1], sat_tanh<kExact>(p.y + qreg[k][1]), s);
          s = fmaf(wreg[k][2], sat_tanh<kExact>(p.z + qreg[k][2]), s);
          s = fmaf(wreg[k][3], sat_tanh<kExact>(p.w + qreg[k][3]), s);
        }
      }
      s = warp_sum(s);
      if (lane == 0) e[r0 + l] = s * scale;
    }
    __syncwarp();
    if (lane == 0) sat_mbar_arrive(&hd->empty[st]);
  }
  sat_named_bar(1, ATTP_CONSUMERS);
  // softmax over L (consumer-only reductions)
  float mx = -INFINITY;
  for (int l = tid; l < L; l += ATTP_CONSUMERS) mx = fmaxf(mx, e[l]);
  mx = warp_max(mx);
  if (lane == 0) hd->red_a[warp] = mx;
  sat_named_bar(1, ATTP_CONSUMERS);
  mx = hd->red_a[0];
#pragma unroll
  for (int w2 = 1; w2 < ATTP_CWARPS; ++w2) mx = fmaxf(mx, hd->red_a[w2]);
  float sum = 0.0f;
  for (int l = tid; l < L; l += ATTP_CONSUMERS) {
    const float p = sat_exp<kExact>(e[l] - mx);
    e[l] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) hd->red_b[warp] = sum;
  sat_named_bar(1, ATTP_CONSUMERS);
  sum = 0.0f;
#pragma unroll
  for (int w2 = 0; w2 < ATTP_CWARPS; ++w2) sum += hd->red_b[w2];
  for (int l = tid; l < L; l += ATTP_CONSUMERS) {
    const float al = e[l] / sum;
    e[l] = al;
    alpha_b[l] = al;
  }
  sat_named_bar(1, ATTP_CONSUMERS);

  // context from the annotation stages: thread (rg, cv0) owns column vector cv0 and the rows
  // [rg*RPT, rg*RPT+RPT) of every stage (consecutive rows -> constant-stride, unrolled LDS.128)
  float acc[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) acc[i] = 0.0f;
  const int rg = RG == 1 ? 0 : tid / NV;
  const int cv0 = RG == 1 ? tid : tid - rg * NV;
  const int RPT = (RCA + RG - 1) / RG;
  const bool worker = rg < RG && cv0 < NV;
  for (int j = 0; j < nA; ++j, ++it) {
    const int st = it % ATTP_NST;
    sat_mbar_wait(&hd->full[st], (uint32_t)(it / ATTP_NST) & 1u);
    const T* As = reinterpret_cast<const T*>(stages + (size_t)st * ATTP_STAGE_BYTES);
    const int r0 = j * RCA, rows = min(RCA, L - r0);
    if (worker) {
      const int lb = rg * RPT, le = min(rows, lb + RPT);
      const T* ap = As + (size_t)lb * D + cv0 * VN;
      const float* ep = e + r0 + lb;
      int l = lb;
      for (; l + 4 <= le; l += 4, ap += 4 * (size_t)D, ep += 4) {
        float v[4][VN];
#pragma unroll
        for (int u = 0; u < 4; ++u) Vec16<T>::load_shared(ap + (size_t)u * D, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float al = ep[u];
#pragma unroll
          for (int i2 = 0; i2 < VN; ++i2) acc[i2] = fmaf(al, v[u][i2], acc[i2]);
        }
      }
      for (; l < le; ++l, ap += D, ++ep) {
        float v[VN];
        Vec16<T>::load_shared(ap, v);
        const float al = ep[0];
#pragma unroll
        for (int i2 = 0; i2 < VN; ++i2) acc[i2] = fmaf(al, v[i2], acc[i2]);
      }
    }
    __syncwarp();
    if (lane == 0) sat_mbar_arrive(&hd->empty[st]);
  }
  if (rg < RG && cv0 < NV) {
#pragma unroll
    for (int i = 0; i < VN; ++i) red[rg * D + cv0 * VN + i] = acc[i];
  }
  sat_named_bar(1, ATTP_CONSUMERS);
  for (int d = tid; d < D; d += ATTP_CONSUMERS) {
    float zs = 0.0f;
    for (int r = 0; r < RG; ++r) zs += red[r * D + d];
    const float bt = sat_sigmoid<kExact>(hp_b[A + d]);
    z[(int64_t)b * ld_z + d] = from_f<T>(zs);
    gz[(int64_t)b * ld_z + d] = from_f<T>(bt * zs);
    if (beta) beta[(int64_t)b * ld_z + d] = from_f<T>(bt);
  }
}

static inline size_t attention_fwd_pipe_smem(int L, int D, int A, int vn) {
  const int NV = D / vn;
  const int RG = NV >= ATTP_CONSUMERS ? 1 : ATTP_CONSUMERS / NV;
  return sizeof(AttPipeSmem) + sizeof(float) * (size_t)(((L + 3) & ~3) + 2 * A + (size_t)RG * D) + 128 +
         (size_t)ATTP_NST * ATTP_STAGE_BYTES;
}

// The pipelined kernel needs: one row of P and of the annotations to fit a stage, and at most 256 16-byte
// column vectors per annotation row (D <= 2048 bf16 / 1024 fp32); other shapes use the plain kernel.
template <typename T>
static inline bool attention_pipe_ok(int L, int D, int A) {
  static int mode = -2;                       // SAT_ATT_MODE=0 forces the plain single-pass kernel (A/B timing)
  if (mode == -2) {
    const char* e = getenv("SAT_ATT_MODE");
    mode = e ? atoi(e) : -1;
  }
  if (mode == 0) return false;
  (void)L;
  return D * (int)sizeof(T) <= ATTP_STAGE_BYTES && A <= 128 * ATTP_KA && D / Vec16<T>::N <= ATTP_CONSUMERS;
}

// launch helper shared by the training and decode drivers
template <typename T, bool kExact>
static int launch_attention_fwd(const T* ann, const T* P, const float* wf, const float* hp, int64_t ldhp, const int32_t* lens,
                                int t, int rows, int ncap, int L, int D, int A, float scale, float* alpha, int64_t ld_alpha,
                                float* qsave, T* z, T* gz, T* beta, int64_t ld_z, cudaStream_t st) {
  if (!attention_pipe_ok<T>(L, D, A)) {
    const size_t sm1 = attention_fwd_smem(L, D, A, Vec16<T>::N);
    auto k1 = attention_step_fwd_kernel<T, kExact>;
    if (sm1 > 48 * 1024) SAT_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    k1<<<rows, ATT_THREADS, sm1, st>>>(ann, P, wf, hp, ldhp, lens, t, ncap, L, D, A, scale, alpha, ld_alpha, qsave, z, gz, beta, ld_z);
    SAT_COUNT_LAUNCH();
    SAT_LAUNCH_OK();
    return 0;
  }
  const size_t smem = attention_fwd_pipe_smem(L, D, A, Vec16<T>::N);
  auto kern = attention_step_fwd_pipe_kernel<T, kExact>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    SAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  kern<<<rows, ATTP_THREADS, smem, st>>>(ann, P, wf, hp, ldhp, lens, t, ncap, L, D, A, scale, alpha, ld_alpha, qsave, z, gz, beta,
                                         ld_z);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

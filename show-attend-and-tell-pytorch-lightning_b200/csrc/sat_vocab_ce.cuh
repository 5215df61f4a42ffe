// Vocabulary projection fused with what consumes it (DeepOutput's last Linear, model.py:130, + LabelSmoothing, util.py:105-112,
// + accuracy, model.py:596-597; or + log-softmax / masks / arg-max of greedy decoding, model.py:330-343): row-owner epilogues
// of the tcgen05 GEMM core (sat_gemm_tc.cuh).  The [T*B, V] logits never reach HBM:
//   pass 1 (VOCAB_STATS)   per (row, 128-column tile): max, sum exp(x - max), sum x, arg-max, and the target logit
//   ce_finalize_kernel     per row: log-sum-exp, token loss, arg-max
//   pass 2 (VOCAB_DLOGITS) recomputes the tile and writes dlogits = (softmax - target distribution) / N_tok in the operand
//                          dtype -- the only [T*B, V] array of the training step (it feeds dpre = dlogits * Wo and dWo)
//   VOCAB_GREEDY           per (row, tile): soft-max statistics of x / temperature and the best non-masked word;
//                          greedy_finalize_kernel turns them into the row's candidate (score, word)
// After tcgen05.ld a thread holds one accumulator row, so all of these reductions are thread-local; the two column halves
// of a tile (two warps per TMEM lane quarter) are combined through the drained operand ring.
#pragma once
#include "sat_gemm_tc.cuh"

namespace tc {

constexpr float LOG2E_F = 1.4426950408889634f;
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

enum { VOCAB_STATS = 0, VOCAB_DLOGITS = 1, VOCAB_GREEDY = 2 };

struct VocabArgs {
  const float* bias;        // [N] over the storage vocabulary, or nullptr
  int V0;                   // true vocabulary size: columns >= V0 are padding and never count
  int NT;                   // 128-column tiles per row (row pitch of `stats`)
  float4* stats;            // [M][NT]
  // cross entropy (STATS / DLOGITS): row m = t*B + b, target y = caps[b*caplen + t + 1], active iff t < lens[b]
  const int32_t* caps;
  const int32_t* lens;
  int B, caplen;
  float* row_xt;            // [M] logit of the target word (written by the one thread that owns its column)
  const float* row_lse;     // [M]                                     (DLOGITS)
  const float* inv_ntok_p;  //                                         (DLOGITS)
  float smoothing;
  bf16* dlogits;            // [M, ldd]                                (DLOGITS)
  int64_t ldd;
  // greedy decode (GREEDY): row r is live iff alive[r] != 0
  const int32_t* alive;
  float inv_temp;
  int tokPAD, tokSTART, tokEND, tokUNK, step0;
};

template <int MODE>
struct EpiVocab {
  static constexpr bool kRowOwner = true;
  VocabArgs a;

  // bias tile of this CTA -> aux[0..128); -inf beyond the true vocabulary (padding and the N tail), so that padded
  // columns drop out of every maximum / sum without further tests
  __device__ __forceinline__ void prologue(float* aux, int n0, int N, int warp, int lane) const {
    const int i = (warp - 2) * 32 + lane;
    if (i < 128) {
      const int n = n0 + i;
      aux[i] = n < a.V0 ? (a.bias ? a.bias[n] : 0.0f) : -INFINITY;
    }
    (void)N;
  }

  template <int BN>
  __device__ __forceinline__ void run(uint8_t* scratch, const float* aux, uint32_t tmem, bool has_acc, int warp, int lane, int m0,
                                      int n0, int M, int N) const {
    static_assert(BN == 128, "the vocabulary epilogues work on 128-column tiles");
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane, m = m0 + r;
    const int cbase = half * 64;                                   // first tile column of this thread
    asm volatile("bar.sync 1, 256;" ::: "memory");                 // bias tile staged by prologue()
    bool active = m < M;
    int y = -1;
    if (MODE == VOCAB_GREEDY) {
      active = active && a.alive[m] != 0;
    } else {
      const int t = m / a.B, b = m - t * a.B;
      active = active && t < a.lens[b];
      if (active) y = a.caps[(int64_t)b * a.caplen + t + 1];
    }
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase;

    if (MODE == VOCAB_DLOGITS) {
      // ---- pass 2: dlogits tile, staged as bf16 [128][128 + 8] and written out with whole rows per half warp ----
      constexpr int PITCH = (BN + 8) * 2;                          // bytes; 272: quarter-warp 16-byte stores hit distinct banks
      const float lse = active ? a.row_lse[m] : 0.0f;
      const float inv_ntok = *a.inv_ntok_p;
      const float sv = a.smoothing / (float)a.V0, conf = 1.0f - a.smoothing;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        float v[16];
        if (has_acc) tmem_ld16(taddr + ch * 16, v);
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0f;
        }
        const int c0 = cbase + ch * 16;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float x = v[j + u] + aux[c0 + j + u];
            float p = ex2_fast((x - lse) * LOG2E_F) - sv;
            if (n0 + c0 + j + u == y) p -= conf;
            p2[u] = (active && x > -INFINITY) ? p * inv_ntok : 0.0f;
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(p2[0], p2[1]);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(scratch + (size_t)r * PITCH + (size_t)c0 * 2);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int ew = warp - 2;
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 16 + ew * 2 + (lane >> 4);
        const int cc = (lane & 15) * 8;
        const int mm = m0 + rr, n = n0 + cc;
        if (mm < M && n < N)
          *reinterpret_cast<uint4*>(a.dlogits + (int64_t)mm * a.ldd + n) = *reinterpret_cast<const uint4*>(scratch + (size_t)rr * PITCH + (size_t)cc * 2);
      }
      return;
    }

    // ---- STATS / GREEDY: running soft-max statistics of this thread's 64 columns ----
    float mx = -INFINITY, se = 0.0f, sx = 0.0f, xt = 0.0f, bestv = -INFINITY;
    int arg = 0x7fffffff, has_xt = 0;
    const float xs = MODE == VOCAB_GREEDY ? a.inv_temp : 1.0f;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      float v[16];
      if (has_acc) tmem_ld16(taddr + ch * 16, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
      const int c0 = cbase + ch * 16, col0 = n0 + c0;
      const float old = mx;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = (v[j] + aux[c0 + j]) * xs;
        v[j] = x;
        mx = fmaxf(mx, x);
      }
      if (MODE == VOCAB_GREEDY) {
        // candidate set: <START>, <PAD> never; <END>, <UNK> not at step 0 (model.py:333,340); first index wins ties
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = col0 + j;
          const bool masked = col == a.tokSTART || col == a.tokPAD || (a.step0 && (col == a.tokEND || col == a.tokUNK));
          if (!masked && v[j] > bestv) { bestv = v[j]; arg = col; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (v[j] > bestv) { bestv = v[j]; arg = col0 + j; }      // arg-max over every word (model.py:596)
        if ((unsigned)(y - col0) < 16u) {
          const int jy = y - col0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j == jy) xt = v[j];
          has_xt = 1;
        }
      }
      if (mx > -INFINITY) {
        if (mx > old) se *= ex2_fast((old - mx) * LOG2E_F);        // old = -inf: se is still 0
        float s0 = 0.0f, s1 = 0.0f, t0 = 0.0f, t1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          s0 += ex2_fast((v[j] - mx) * LOG2E_F);
          s1 += ex2_fast((v[j + 1] - mx) * LOG2E_F);
          t0 += v[j] > -INFINITY ? v[j] : 0.0f;
          t1 += v[j + 1] > -INFINITY ? v[j + 1] : 0.0f;
        }
        se += s0 + s1;
        sx += t0 + t1;
      }
    }
    // combine the two column halves of the tile (same row, warps w and w + 4) through the drained ring
    float* xch = reinterpret_cast<float*>(scratch);
    if (half == 1) {
      float* o = xch + r * 8;
      o[0] = mx; o[1] = se; o[2] = sx; o[3] = bestv; o[4] = __int_as_float(arg); o[5] = xt; o[6] = __int_as_float(has_xt);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0 && active) {
      const float* o = xch + r * 8;
      const float mx1 = o[0], se1 = o[1], sx1 = o[2], bv1 = o[3];
      const int arg1 = __float_as_int(o[4]);
      const float mm = fmaxf(mx, mx1);
      float s = 0.0f;
      if (mx > -INFINITY) s += se * ex2_fast((mx - mm) * LOG2E_F);
      if (mx1 > -INFINITY) s += se1 * ex2_fast((mx1 - mm) * LOG2E_F);
      if (bv1 > bestv) { bestv = bv1; arg = arg1; }                // ties keep the lower columns (half 0)
      if (MODE == VOCAB_GREEDY) {
        a.stats[(int64_t)m * a.NT + blockIdx.x] = make_float4(mm, s, bestv, __int_as_float(arg));
      } else {
        a.stats[(int64_t)m * a.NT + blockIdx.x] = make_float4(mm, s, sx + sx1, __int_as_float(arg));
        if (has_xt) a.row_xt[m] = xt;
        else if (__float_as_int(o[6]) != 0) a.row_xt[m] = o[5];
      }
    }
    (void)N;
  }
};

// Row-wise finish of the fused cross entropy: log-sum-exp over the tile statistics, token loss (util.py:105-112) and arg-max
// (model.py:596).  One warp per row, fixed combination order.
static __global__ void __launch_bounds__(256)
ce_finalize_kernel(const float4* __restrict__ stats, int NT, const float* __restrict__ row_xt, const int32_t* __restrict__ lens, int B,
                   int M, int V0, float smoothing, float* __restrict__ row_lse, float* __restrict__ row_loss,
                   int32_t* __restrict__ row_argmax) {
  SAT_PDL_TRIGGER();
  SAT_PDL_WAIT();
  const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int t = m / B, b = m - t * B;
  if (t >= lens[b]) {
    if (lane == 0) { row_loss[m] = 0.0f; row_argmax[m] = -1; row_lse[m] = 0.0f; }
    return;
  }
  const float4* st = stats + (int64_t)m * NT;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int i = lane; i < NT; i += 32) {
    const float4 s4 = st[i];
    if (s4.x > mx) { mx = s4.x; arg = __float_as_int(s4.w); }      // a tile's max is the value of its arg-max; tiles ascend
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.0f, sx = 0.0f;
  for (int i = lane; i < NT; i += 32) {
    const float4 s4 = st[i];
    se += s4.y * __expf(s4.x - mx);
    sx += s4.z;
  }
  se = warp_sum(se);
  sx = warp_sum(sx);
  if (lane == 0) {
    const float lse = mx + __logf(se);
    const float nll = lse - row_xt[m];
    const float smooth = lse - sx / (float)V0;
    row_lse[m] = lse;
    row_loss[m] = (1.0f - smoothing) * nll + smoothing * smooth;
    row_argmax[m] = arg;
  }
}

// Greedy decode: candidate (score, word) of every live row from the tile statistics -- the k = 1 output of
// row_topk_kernel (log-softmax(x / T), masks, + parent score; model.py:330-351) without materialising the logits.
static __global__ void __launch_bounds__(256)
greedy_finalize_kernel(const float4* __restrict__ stats, int NT, const int32_t* __restrict__ alive, const float* __restrict__ top_scores,
                       int R, int step0, float* __restrict__ cand_val, int32_t* __restrict__ cand_idx) {
  SAT_PDL_TRIGGER();
  SAT_PDL_WAIT();
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R || alive[r] == 0) return;
  const float4* st = stats + (int64_t)r * NT;
  float mx = -INFINITY, bv = -INFINITY;
  int arg = 0x7fffffff;
  for (int i = lane; i < NT; i += 32) {
    const float4 s4 = st[i];
    mx = fmaxf(mx, s4.x);
    if (s4.z > bv) { bv = s4.z; arg = __float_as_int(s4.w); }
  }
  mx = warp_max(mx);
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > bv || (ov == bv && oa < arg)) { bv = ov; arg = oa; }
  }
  float se = 0.0f;
  for (int i = lane; i < NT; i += 32) {
    const float4 s4 = st[i];
    se += s4.y * __expf(s4.x - mx);
  }
  se = warp_sum(se);
  if (lane == 0) {
    const float lp = (bv - mx) - __logf(se);
    cand_val[r] = step0 ? lp : lp + top_scores[r];
    cand_idx[r] = arg;
  }
}

template <int MODE>
static int launch_vocab(const GemmOperandA& A, const bf16* W, int64_t ldw, int M, int N, const VocabArgs& va, cudaStream_t stream) {
  EpiVocab<MODE> epi{va};
  return launch_bn<128, EpiVocab<MODE>>(A, W, ldw, M, N, epi, stream);
}

}  // namespace tc

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
#include <stdlib.h>

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

  // NW epilogue warps (8: one-tile-per-CTA kernel, 16: persistent kernel): warp w reads TMEM lane quarter w & 3 (rows) and the
  // column part (w - 2) >> 2 of the tile, 128 / (NW / 4) columns.
  template <int BN, int NW>
  __device__ __forceinline__ void run(uint8_t* scratch, const float* aux, uint32_t tmem, bool has_acc, int warp, int lane, int m0,
                                      int n0, int M, int N, int tile_n) const {
    static_assert(BN == 128, "the vocabulary epilogues work on 128-column tiles");
    static_assert(NW == 8 || NW == 16, "two or four column parts");
    constexpr int PARTS = NW / 4, PCOLS = BN / PARTS, NCH = PCOLS / 16;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int r = q * 32 + lane, m = m0 + r;
    const int cbase = part * PCOLS;                                // first tile column of this thread
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");     // bias tile staged by prologue()
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
    const bool tile_full = n0 + BN <= a.V0;                        // no padded / out-of-range column in this tile (CTA-uniform)
    if (MODE == VOCAB_DLOGITS) {
      // ---- pass 2: dlogits tile, staged as bf16 [128][128 + 8] and written out with whole rows per half warp ----
      constexpr int PITCH = (BN + 8) * 2;                          // bytes; 272: quarter-warp 16-byte stores hit distinct banks
      const float nlse = active ? -a.row_lse[m] * LOG2E_F : -INFINITY;     // finished rows: exp2(-inf) = 0, no overflow
      const float inv_ntok = active ? *a.inv_ntok_p : 0.0f;        // finished rows: all-zero gradient
      const float sv = a.smoothing / (float)a.V0, conf = 1.0f - a.smoothing;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        float v[16];
        if (has_acc) tmem_ld16(taddr + ch * 16, v);
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0f;
        }
        const int c0 = cbase + ch * 16;
        const int jy = y - (n0 + c0);                              // target column inside this chunk, or out of [0,16)
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float x = v[j + u] + aux[c0 + j + u];
            float p = ex2_fast(fmaf(x, LOG2E_F, nlse)) - sv;       // softmax - s/V   (x = -inf: -s/V, zeroed below)
            if (j + u == jy) p -= conf;
            p2[u] = (tile_full || x > -INFINITY) ? p * inv_ntok : 0.0f;
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(p2[0], p2[1]);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(scratch + (size_t)r * PITCH + (size_t)c0 * 2);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
      const int ew = warp - 2;
#pragma unroll 1
      for (int it = 0; it < BM / (NW * 2); ++it) {
        const int rr = it * (NW * 2) + ew * 2 + (lane >> 4);
        const int cc = (lane & 15) * 8;
        const int mm = m0 + rr, n = n0 + cc;
        if (mm < M && n < N)
          *reinterpret_cast<uint4*>(a.dlogits + (int64_t)mm * a.ldd + n) = *reinterpret_cast<const uint4*>(scratch + (size_t)rr * PITCH + (size_t)cc * 2);
      }
      return;
    }

    // ---- STATS / GREEDY: running soft-max statistics of this thread's columns (short dependency chains: pairwise trees) ----
    float mx = -INFINITY, se = 0.0f, sx = 0.0f, xt = 0.0f, bestv = -INFINITY;
    int arg = 0x7fffffff, has_xt = 0;
    const float xs = MODE == VOCAB_GREEDY ? a.inv_temp : 1.0f;
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
      float v[16];
      if (has_acc) tmem_ld16(taddr + ch * 16, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
      const int c0 = cbase + ch * 16, col0 = n0 + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (v[j] + aux[c0 + j]) * xs;
      // chunk arg-max as a tree of (value, index) pairs; a tie keeps the lower index (torch.argmax / topk order)
      float cv[16];
      int ci[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        cv[j] = v[j];
        ci[j] = j;
        if (MODE == VOCAB_GREEDY) {
          // candidate set: <START>, <PAD> never; <END>, <UNK> not at step 0 (model.py:333,340)
          const int col = col0 + j;
          if (col == a.tokSTART || col == a.tokPAD || (a.step0 && (col == a.tokEND || col == a.tokUNK))) cv[j] = -INFINITY;
        }
      }
#pragma unroll
      for (int w = 1; w < 16; w <<= 1) {
#pragma unroll
        for (int j = 0; j + w < 16; j += 2 * w) {
          const bool keep = cv[j] >= cv[j + w];
          cv[j] = keep ? cv[j] : cv[j + w];
          ci[j] = keep ? ci[j] : ci[j + w];
        }
      }
      if (cv[0] > bestv) { bestv = cv[0]; arg = col0 + ci[0]; }    // strictly greater: earlier chunks win ties
      float cm;                                                    // chunk max over EVERY column (soft-max statistics ignore the masks)
      if (MODE == VOCAB_GREEDY) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = fmaxf(v[2 * j], v[2 * j + 1]);
        cm = fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7])));
      } else {
        cm = cv[0];
        if ((unsigned)(y - col0) < 16u) {
          const int jy = y - col0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j == jy) xt = v[j];
          has_xt = 1;
        }
      }
      if (cm > -INFINITY) {
        if (cm > mx) {
          se *= ex2_fast((mx - cm) * LOG2E_F);                     // mx = -inf: se is still 0
          mx = cm;
        }
        const float nm = -mx * LOG2E_F;
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          s4[j & 3] += ex2_fast(fmaf(v[j], LOG2E_F, nm));
          if (MODE != VOCAB_GREEDY) t4[j & 3] += (tile_full || v[j] > -INFINITY) ? v[j] : 0.0f;
        }
        se += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        sx += (t4[0] + t4[1]) + (t4[2] + t4[3]);
      }
    }
    // combine the column parts of the tile (same row, warps w, w + 4, ..) through shared memory, lower columns first
    float* xch = reinterpret_cast<float*>(scratch);
    if (part > 0) {
      float* o = xch + ((part - 1) * BM + r) * 8;
      o[0] = mx; o[1] = se; o[2] = sx; o[3] = bestv; o[4] = __int_as_float(arg); o[5] = xt; o[6] = __int_as_float(has_xt);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    if (part == 0 && active) {
#pragma unroll
      for (int p2 = 1; p2 < PARTS; ++p2) {
        const float* o = xch + ((p2 - 1) * BM + r) * 8;
        const float mx1 = o[0], se1 = o[1], bv1 = o[3];
        const float mm = fmaxf(mx, mx1);
        float s = 0.0f;
        if (mx > -INFINITY) s += se * ex2_fast((mx - mm) * LOG2E_F);
        if (mx1 > -INFINITY) s += se1 * ex2_fast((mx1 - mm) * LOG2E_F);
        mx = mm;
        se = s;
        sx += o[2];
        if (bv1 > bestv) { bestv = bv1; arg = __float_as_int(o[4]); }   // ties keep the lower columns
        if (__float_as_int(o[6]) != 0) { xt = o[5]; has_xt = 1; }
      }
      if (MODE == VOCAB_GREEDY) {
        a.stats[(int64_t)m * a.NT + tile_n] = make_float4(mx, se, bestv, __int_as_float(arg));
      } else {
        a.stats[(int64_t)m * a.NT + tile_n] = make_float4(mx, se, sx, __int_as_float(arg));
        if (has_xt) a.row_xt[m] = xt;
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
  // persistent kernel (one CTA per SM, double-buffered tensor memory) at every size, so that a row's statistics do not depend
  // on the batch it is computed in (the two kernels split a tile's columns differently); SAT_VOCAB_PERSIST=0 forces the
  // one-tile-per-CTA kernel (A/B timing)
  static const int pers_mode = getenv("SAT_VOCAB_PERSIST") ? atoi(getenv("SAT_VOCAB_PERSIST")) : 1;
  if (pers_mode != 0 && persistent_ok(A)) return launch_persistent<EpiVocab<MODE>>(A, W, ldw, M, N, epi, stream);
  return launch_bn<128, EpiVocab<MODE>>(A, W, ldw, M, N, epi, stream);
}

}  // namespace tc

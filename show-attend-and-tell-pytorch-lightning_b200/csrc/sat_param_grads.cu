// sat_train_param_grads: the parameter gradients of the teacher-forced step (what autograd accumulates into the 18 decoder
// parameters of model.py:158-199 when PL calls loss.backward()), computed inside the library from the buffers that
// sat_train_forward / sat_train_backward left behind:
//   * weight gradients  dW = dY^T X  are reductions over the T*B time-major rows: "NT" GEMMs on the tcgen05 core (both
//     operands MN-major straight out of the activation buffers, sat_gemm_tc.cuh) or the SIMT core (fp32 parity mode),
//     split over k with per-split partials;
//   * bias / f_att gradients are column sums (colsum_partial_kernel);
//   * the embedding gradient is a segment sum over the rows that fed each word (embed_grad_kernel: one warp per word walks
//     the token list in row order -- no atomics, no sort);
//   * ONE finalize kernel adds the partials in fixed order, scales by the upstream gradient where the operand does not
//     carry it yet, undoes the gate interleave / operand padding and writes the gradients with the reference's shapes.
// Everything is deterministic: the reference trains with deterministic=True (train.py:271).
#include "sat_gemm.cuh"
#include "sat_kernels.cuh"

namespace {

// ---- column sums: part[chunk][c] = sum over the chunk's rows of X[r][c] ---------------------------------------------
constexpr int CS_ROWS = 64;
template <typename T>
__global__ void __launch_bounds__(128) colsum_partial_kernel(const T* __restrict__ X, int64_t ld, int R, int C, float* __restrict__ part) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= C) return;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld4(X + (int64_t)(r + u) * ld + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; r < r1; ++r) {
    const float4 v = ld4(X + (int64_t)r * ld + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * C + c) = acc;
}

template <typename T>
static int launch_colsum(const T* X, int64_t ld, int R, int C, float* part, cudaStream_t st) {
  dim3 grid((C / 4 + 127) / 128, (R + CS_ROWS - 1) / CS_ROWS);
  colsum_partial_kernel<T><<<grid, 128, 0, st>>>(X, ld, R, C, part);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

// ---- out[i, :] = sum over the ncap caption rows of image i (fixed order) ---------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) ncap_sum_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t per_img, int ncap,
                                                       int64_t n_out) {
  const int64_t i4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i4 >= n_out) return;
  const int64_t img = i4 / per_img, off = i4 - img * per_img;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < ncap; ++c) {
    const float4 v = ld4(in + (img * ncap + c) * per_img + off);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  st4(out + i4, acc);
}

// ---- embedding gradient: dEmb[v, :] = sum over the rows m with tok[m] == v of dXe[m, :] ------------------------------
// A CTA owns EG_TPC consecutive vocabulary entries.  The token list is staged in shared memory chunk by chunk; for every
// owned word the chunk's 32-row windows are dealt round-robin to the CTA's 8 warps; a warp tests its window with a ballot
// and adds the matching rows of dXe in ascending row order, four rows in flight (independent loads, ordered adds); the 8
// partial sums are added in warp order at the end.  The order of additions is fixed by (row, warp) alone: no
// atomics, no sort, bit-identical from run to run (index_add_ with float atomics is not).  A frequent word (<START> feeds
// every caption's first step) is spread over 8 warps x 4 loads in flight instead of one serial chain.  The <PAD> row stays
// zero (nn.Embedding(padding_idx), model.py:162).
constexpr int EG_TPC = 4;                // words per CTA
constexpr int EG_COLS = 256;             // columns per pass: 32 lanes x 2 float4
constexpr int EG_CHUNK = 4096;           // token ids staged per chunk (8 warps x 512 rows)
__global__ void __launch_bounds__(256)
embed_grad_kernel(const int32_t* __restrict__ tok, const float* __restrict__ dXe, int64_t ld_dxe, int M, int V0, int E0, int pad_idx,
                  float* __restrict__ dEmb) {
  __shared__ int32_t s_tok[EG_CHUNK];
  __shared__ __align__(16) float part[8][EG_COLS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = blockIdx.x * EG_TPC;
  for (int e0 = 0; e0 < E0; e0 += EG_COLS) {
    float4 acc[EG_TPC][2];
#pragma unroll
    for (int k = 0; k < EG_TPC; ++k) acc[k][0] = acc[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int mb = 0; mb < M; mb += EG_CHUNK) {
      const int nm = min(EG_CHUNK, M - mb);
      __syncthreads();
      for (int i = threadIdx.x; i < nm; i += 256) s_tok[i] = tok[mb + i];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < EG_TPC; ++k) {
        const int v = v0 + k;
        if (v >= V0 || v == pad_idx) continue;
        // 32-row windows are dealt round-robin to the warps: rows are time-major, so the rows that feed one word cluster
        // (every caption's first step feeds <START>: B consecutive rows) and contiguous ranges would leave them to one warp
        for (int i = warp * 32; i < nm; i += 8 * 32) {
          const int t = (i + lane) < nm ? s_tok[i + lane] : -1;
          unsigned hit = __ballot_sync(0xffffffffu, t == v);
          while (hit) {
            int r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              r[u] = hit ? mb + i + __ffs(hit) - 1 : -1;
              hit &= hit - 1;                                    // 0 stays 0
            }
            float4 x[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int e = e0 + lane * 4 + 128 * j;           // ld_dxe is a multiple of 4: whole float4s stay inside the row
                x[u][j] = (r[u] >= 0 && e < E0) ? *reinterpret_cast<const float4*>(dXe + (int64_t)r[u] * ld_dxe + e) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {                        // rows are added in ascending order
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                acc[k][j].x += x[u][j].x; acc[k][j].y += x[u][j].y; acc[k][j].z += x[u][j].z; acc[k][j].w += x[u][j].w;
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < EG_TPC; ++k) {
      __syncthreads();
      *reinterpret_cast<float4*>(&part[warp][lane * 4]) = acc[k][0];
      *reinterpret_cast<float4*>(&part[warp][lane * 4 + 128]) = acc[k][1];
      __syncthreads();
      const int v = v0 + k, c = threadIdx.x;
      if (v < V0 && e0 + c < E0) {
        float sum = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) sum += part[w2][c];
        dEmb[(int64_t)v * E0 + e0 + c] = sum;
      }
    }
  }
}

// ---- finalize: partial sums -> gradients in the reference's layout ---------------------------------------------------
struct FinJob {
  const float* src;        // partials: element (z, rs, c) at src[z * split_stride + rs * src_ld + src_c0 + c]
  float* dst;              // gradient: element (rd, c) at dst[rd * dst_ld + dst_c0 + c]
  int64_t split_stride, src_ld, dst_ld;
  int nsplit, src_c0, dst_c0;
  int rows, cols;          // true (unpadded) extent of the destination block
  int inter_h;             // > 0: destination row g*H0 + j comes from packed row 4*j + g (gate interleave, H0 = inter_h)
  int scale_g;             // multiply by the upstream loss gradient (operands computed before it was known)
  int accumulate;          // dst += (weight tying: dW_o is added to the embedding gradient)
  int blk0;                // first 256-element block of this job in the grid
};
constexpr int FIN_MAXJOBS = 40;
struct FinTable {
  FinJob job[FIN_MAXJOBS];
  int njobs;
};

__global__ void __launch_bounds__(256) param_grads_finalize_kernel(const __grid_constant__ FinTable tab, const float* __restrict__ gscale) {
  // a thread finishes 4 consecutive columns of one destination row: 16-byte loads from every split (the partials' pitches
  // are multiples of 8 floats), 32-bit index arithmetic
  int j = 0;
  while (j + 1 < tab.njobs && (int)blockIdx.x >= tab.job[j + 1].blk0) ++j;
  const FinJob& J = tab.job[j];
  const unsigned cols4 = (unsigned)(J.cols + 3) >> 2;
  const unsigned idx = (unsigned)(blockIdx.x - J.blk0) * 256u + threadIdx.x;
  if (idx >= (unsigned)J.rows * cols4) return;
  const int rd = (int)(idx / cols4), c = (int)(idx - (unsigned)rd * cols4) * 4;
  const int rs = J.inter_h > 0 ? 4 * (rd % J.inter_h) + rd / J.inter_h : rd;
  const float* p = J.src + (int64_t)rs * J.src_ld + J.src_c0 + c;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = c + 4 <= J.cols && (((uintptr_t)p | (uintptr_t)(J.split_stride * 4)) & 15) == 0;
  if (vec) {
    int z = 0;
    for (; z + 4 <= J.nsplit; z += 4) {      // four splits in flight; the sum order stays z = 0, 1, 2, ...
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const float4*>(p + (int64_t)(z + u) * J.split_stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s[0] += q[u].x; s[1] += q[u].y; s[2] += q[u].z; s[3] += q[u].w; }
    }
    for (; z < J.nsplit; ++z) {
      const float4 q = *reinterpret_cast<const float4*>(p + (int64_t)z * J.split_stride);
      s[0] += q.x; s[1] += q.y; s[2] += q.z; s[3] += q.w;
    }
  } else {
    for (int z = 0; z < J.nsplit; ++z)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < J.cols) s[u] += p[(int64_t)z * J.split_stride + u];
  }
  const float g = (J.scale_g && gscale) ? *gscale : 1.0f;
  float* o = J.dst + (int64_t)rd * J.dst_ld + J.dst_c0 + c;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (c + u < J.cols) o[u] = J.accumulate ? o[u] + s[u] * g : s[u] * g;
}

struct FinBuilder {
  FinTable t{};
  int blocks = 0;
  bool ok = true;
  void add(const float* src, int nsplit, int64_t split_stride, int64_t src_ld, int src_c0, float* dst, int64_t dst_ld, int dst_c0, int rows,
           int cols, int inter_h, int scale_g, int accumulate) {
    if (dst == nullptr || rows <= 0 || cols <= 0) return;
    if (t.njobs >= FIN_MAXJOBS) { ok = false; return; }
    FinJob& j = t.job[t.njobs++];
    j.src = src; j.dst = dst; j.split_stride = split_stride; j.src_ld = src_ld; j.dst_ld = dst_ld; j.nsplit = nsplit;
    j.src_c0 = src_c0; j.dst_c0 = dst_c0; j.rows = rows; j.cols = cols; j.inter_h = inter_h; j.scale_g = scale_g;
    j.accumulate = accumulate; j.blk0 = blocks;
    blocks += (int)(((int64_t)rows * ((cols + 3) / 4) + 255) / 256);
  }
};

// ---- plan: the weight-gradient GEMMs, their split factors and the workspace layout (a pure function of the dims) ------
// stacked LSTM layers l >= 1 add G_WHH0 (layer 0's recurrent weight, split from G_WH3 because q / beta read the top layer) and,
// per layer, G_WLI + l-1 (weight_ih_l{l}), G_WLH + l-1 (weight_hh_l{l}) and the bias column sum C_GL + l-1
enum { G_WO = 0, G_WHO, G_WZO, G_WH3, G_WIHE, G_WIHZ, G_WA, G_WINIT, G_WFACT, G_WHH0, G_WLI, G_WLH = G_WLI + SAT_MAX_LAYERS - 1,
       G_COUNT = G_WLH + SAT_MAX_LAYERS - 1 };
enum { C_BO = 0, C_DY, C_WF, C_INIT, C_FACT, C_GL, C_COUNT = C_GL + SAT_MAX_LAYERS - 1 };
struct Plan {
  int n1[G_COUNT], n2[G_COUNT], kr[G_COUNT], sk[G_COUNT];
  int64_t off[G_COUNT];                 // float offsets into the workspace
  int cs_r[C_COUNT], cs_c[C_COUNT];
  int64_t cs_off[C_COUNT];
  int64_t dpimg_off;                    // [Bi, L, A] fp32 / operand dtype: dP summed over the captions of an image (ncap > 1)
  int64_t total_floats;
};

static Plan make_plan(const SatDims& d) {
  Plan p{};
  const int B = d.B, Bi = d.Bi, L = d.L, D = d.D, A = d.A, E = d.E, H = d.H, V = d.V, T = d.T;
  const int NH3 = A + D + 4 * H, M = T * B;
  const bool tc = d.use_tc != 0 && d.dtype == SAT_BF16;
  auto set = [&](int g, int n1, int n2, int kr) { p.n1[g] = n1; p.n2[g] = n2; p.kr[g] = kr; };
  set(G_WO, V, E, M);
  set(G_WHO, E, H, M);
  set(G_WZO, E, D, M);
  const int nl = d.layers > 1 ? d.layers : 1;
  set(G_WH3, nl == 1 ? NH3 : A + D, H, M);
  set(G_WHH0, nl == 1 ? 0 : 4 * H, H, M);
  for (int l = 1; l < SAT_MAX_LAYERS; ++l) {
    set(G_WLI + l - 1, l < nl ? 4 * H : 0, H, M);
    set(G_WLH + l - 1, l < nl ? 4 * H : 0, H, M);
  }
  set(G_WIHE, 4 * H, E, M);
  set(G_WIHZ, 4 * H, D, M);
  set(G_WA, A, D, Bi * L);
  set(G_WINIT, 2 * nl * H, E, Bi);
  set(G_WFACT, E, D, Bi);
  int64_t o = 0;
  for (int g = 0; g < G_COUNT; ++g) {
    if (tc) {
      p.sk[g] = tc::pick_splitk_nt(p.n1[g], p.n2[g], p.kr[g]);
    } else {      // SIMT: 64x64 tiles, 16-row k steps
      const long tiles = (long)((p.n1[g] + 63) / 64) * ((p.n2[g] + 63) / 64);
      int s = 1;
      while (s < 32 && tiles * (s * 2) <= 600 && p.kr[g] / (s * 2) >= 64) s *= 2;
      p.sk[g] = s;
    }
    p.off[g] = o;
    o += (int64_t)p.sk[g] * p.n1[g] * p.n2[g];
  }
  auto cs = [&](int c, int r, int cols) { p.cs_r[c] = r; p.cs_c[c] = cols; p.cs_off[c] = o; o += (int64_t)((r + CS_ROWS - 1) / CS_ROWS) * cols; };
  cs(C_BO, M, V);
  cs(C_DY, M, NH3);
  cs(C_WF, M, A);
  cs(C_INIT, Bi, 2 * nl * H);
  cs(C_FACT, Bi, E);
  for (int l = 1; l < SAT_MAX_LAYERS; ++l) cs(C_GL + l - 1, l < nl ? M : 0, 4 * H);
  p.dpimg_off = o;
  if (d.ncap > 1) o += (int64_t)Bi * L * A;
  p.total_floats = o;
  return p;
}

template <typename TS>
int param_grads_impl(const SatDims& d, const SatTrainBuffers& b, const SatParamGrads& g, float* ws, cudaStream_t st) {
  const int B = d.B, Bi = d.Bi, L = d.L, D = d.D, A = d.A, E = d.E, H = d.H, V = d.V, T = d.T;
  const int D0 = d.D0 ? d.D0 : D, A0 = d.A0 ? d.A0 : A, E0 = d.E0 ? d.E0 : E, H0 = d.H0 ? d.H0 : H, V0 = d.V0 ? d.V0 : V;
  const int NH3 = A + D + 4 * H, M = T * B;
  const bool bf = std::is_same<TS, bf16>::value;
  const bool tc = d.use_tc != 0 && bf;
  const Plan p = make_plan(d);
  SAT_PROF(7, st);                 // 7 = the whole parameter-gradient stage (closing mark after the finalize kernel)
  SatNoPdlScope first_launch;      // the operands were written by the caller's previous launches (sat_train_backward)

  // one weight-gradient GEMM: partials [sk][n1][n2] at ws + off
  auto nt = [&](int gi, const void* Ap, int64_t lda, bool a_is_ts, const void* Bp, int64_t ldb) -> int {
    float* C = ws + p.off[gi];
    EpiStore<float> epi{C, p.n2[gi], nullptr, nullptr, 0, (int64_t)p.n1[gi] * p.n2[gi]};
    if (tc && a_is_ts && tc::nt_operands_ok(Ap, lda, Bp, ldb))
      return tc::launch_nt<EpiStore<float>>((const bf16*)Ap, lda, (const bf16*)Bp, ldb, p.kr[gi], p.n1[gi], p.n2[gi], epi, st, p.sk[gi]);
    if (a_is_ts)
      return launch_gemm_nt_simt<TS, TS, EpiStore<float>>((const TS*)Ap, lda, (const TS*)Bp, ldb, p.kr[gi], p.n1[gi], p.n2[gi], epi, st, p.sk[gi]);
    return launch_gemm_nt_simt<float, TS, EpiStore<float>>((const float*)Ap, lda, (const TS*)Bp, ldb, p.kr[gi], p.n1[gi], p.n2[gi], epi, st,
                                                           p.sk[gi]);
  };

  const TS* dlog = (const TS*)b.dlogits;
  const TS* DY = (const TS*)b.DY;
  const TS* Hs = (const TS*)b.Hs;
  const int nl = d.layers > 1 ? d.layers : 1;
  const int64_t LS = (int64_t)(T + 1) * B * H, GS = (int64_t)T * B * 4 * H;
  const TS* Hs_top = Hs + (nl - 1) * LS;
  SAT_TRY(nt(G_WO, dlog, V, true, b.Xo, E));
  SAT_TRY(nt(G_WHO, b.dpre, E, true, Hs_top + (int64_t)B * H, H));
  if (!d.plain_output) SAT_TRY(nt(G_WZO, b.dpre, E, true, b.Z, D));
  SAT_TRY(nt(G_WH3, DY, NH3, true, Hs_top, H));          // one layer: [dq | dbeta_pre | dG]^T h;  stacked: [dq | dbeta_pre]^T h_top
  if (nl > 1) {
    SAT_REQUIRE(b.dGl, "sat_train_param_grads: decoder_layers > 1 needs dGl");
    SAT_TRY(nt(G_WHH0, DY + A + D, NH3, true, Hs, H));
    for (int l = 1; l < nl; ++l) {
      const TS* dG_l = (const TS*)b.dGl + (l - 1) * GS;
      SAT_TRY(nt(G_WLI + l - 1, dG_l, 4 * H, true, Hs + (l - 1) * LS + (int64_t)B * H, H));   // input: the new state of the layer below
      SAT_TRY(nt(G_WLH + l - 1, dG_l, 4 * H, true, Hs + l * LS, H));
      SAT_TRY(launch_colsum<TS>(dG_l, 4 * H, M, 4 * H, ws + p.cs_off[C_GL + l - 1], st));
    }
  }
  SAT_TRY(nt(G_WIHE, DY + A + D, NH3, true, b.Xe, E));
  SAT_TRY(nt(G_WIHZ, DY + A + D, NH3, true, b.GZ, D));
  // dW_a = dP^T ann over the Bi*L image locations (dP summed over the captions of an image first when ncap > 1)
  {
    const bool have16 = bf && b.dP16 != nullptr && d.use_tc != 0;      // operand-dtype copy written by the backward kernels
    if (d.ncap > 1) {
      const int64_t per = (int64_t)L * A, n_out = (int64_t)Bi * per;
      float* dpi = ws + p.dpimg_off;
      const unsigned nb = (unsigned)((n_out / 4 + 255) / 256);
      if (tc) {
        ncap_sum_kernel<float, TS><<<nb, 256, 0, st>>>(b.dP, (TS*)dpi, per, d.ncap, n_out);
        SAT_COUNT_LAUNCH();
        SAT_TRY(nt(G_WA, dpi, A, true, b.ann, D));
      } else {
        ncap_sum_kernel<float, float><<<nb, 256, 0, st>>>(b.dP, dpi, per, d.ncap, n_out);
        SAT_COUNT_LAUNCH();
        SAT_TRY(nt(G_WA, dpi, A, std::is_same<TS, float>::value, b.ann, D));
      }
      SAT_LAUNCH_OK();
    } else if (have16) {
      SAT_TRY(nt(G_WA, b.dP16, A, true, b.ann, D));
    } else {
      SAT_TRY(nt(G_WA, b.dP, A, std::is_same<TS, float>::value, b.ann, D));
    }
  }
  {
    const bool have16 = tc && b.d_init_out16 != nullptr && b.df116 != nullptr;
    if (have16) {
      SAT_TRY(nt(G_WINIT, b.d_init_out16, 2 * nl * H, true, b.f1, E));
      SAT_TRY(nt(G_WFACT, b.df116, E, true, b.meanv, D));
    } else {
      SAT_TRY(nt(G_WINIT, b.d_init_out, 2 * nl * H, std::is_same<TS, float>::value, b.f1, E));
      SAT_TRY(nt(G_WFACT, b.df1, E, std::is_same<TS, float>::value, b.meanv, D));
    }
  }
  // column sums
  if (g.out_b) SAT_TRY(launch_colsum<TS>(dlog, V, M, V, ws + p.cs_off[C_BO], st));
  SAT_TRY(launch_colsum<TS>(DY, NH3, M, NH3, ws + p.cs_off[C_DY], st));
  SAT_TRY(launch_colsum<float>(b.dwf_part, A, M, A, ws + p.cs_off[C_WF], st));
  SAT_TRY(launch_colsum<float>(b.d_init_out, 2 * nl * H, Bi, 2 * nl * H, ws + p.cs_off[C_INIT], st));
  SAT_TRY(launch_colsum<float>(b.df1, E, Bi, E, ws + p.cs_off[C_FACT], st));
  // embedding: segment sum of dXe by the word that was fed
  if (g.embedding) {
    SAT_PROF(5, st);
    embed_grad_kernel<<<(V0 + EG_TPC - 1) / EG_TPC, 256, 0, st>>>(b.tok, b.dXe, E, M, V0, E0, g.pad_idx, g.embedding);
    SAT_PROF(5, st);
    SAT_COUNT_LAUNCH();
    SAT_LAUNCH_OK();
  }

  FinBuilder f;
  auto gj = [&](int gi, int src_r0, int src_c0, float* dst, int64_t dst_ld, int dst_c0, int rows, int cols, int inter, int scale, int acc) {
    f.add(ws + p.off[gi] + (int64_t)src_r0 * p.n2[gi], p.sk[gi], (int64_t)p.n1[gi] * p.n2[gi], p.n2[gi], src_c0, dst, dst_ld, dst_c0, rows, cols,
          inter, scale, acc);
  };
  auto cj = [&](int ci, int c0, float* dst, int n, int inter) {
    const int nch = (p.cs_r[ci] + CS_ROWS - 1) / CS_ROWS;
    if (inter > 0) f.add(ws + p.cs_off[ci] + c0, nch, p.cs_c[ci], 1, 0, dst, 1, 0, n, 1, inter, 0, 0);          // [4H0] as a column
    else f.add(ws + p.cs_off[ci], nch, p.cs_c[ci], 0, c0, dst, 0, 0, 1, n, 0, 0, 0);
  };
  if (g.weight_tying) gj(G_WO, 0, 0, g.embedding, E0, 0, V0, E0, 0, 1, 1);        // after embed_grad_kernel in the stream
  else gj(G_WO, 0, 0, g.out_w, E0, 0, V0, E0, 0, 1, 0);
  gj(G_WHO, 0, 0, g.out_hidden, H0, 0, E0, H0, 0, 0, 0);
  if (!d.plain_output) gj(G_WZO, 0, 0, g.out_context, D0, 0, E0, D0, 0, 0, 0);
  gj(G_WH3, 0, 0, g.dec_att, H0, 0, A0, H0, 0, 0, 0);
  gj(G_WH3, A, 0, g.beta_w, H0, 0, D0, H0, 0, 0, 0);
  if (nl == 1) gj(G_WH3, A + D, 0, g.w_hh, H0, 0, 4 * H0, H0, H0, 0, 0);
  else gj(G_WHH0, 0, 0, g.w_hh, H0, 0, 4 * H0, H0, H0, 0, 0);
  for (int l = 1; l < nl; ++l) {
    gj(G_WLI + l - 1, 0, 0, g.w_ih_l[l - 1], H0, 0, 4 * H0, H0, H0, 0, 0);
    gj(G_WLH + l - 1, 0, 0, g.w_hh_l[l - 1], H0, 0, 4 * H0, H0, H0, 0, 0);
    cj(C_GL + l - 1, 0, g.b_ih_l[l - 1], 4 * H0, H0);
    cj(C_GL + l - 1, 0, g.b_hh_l[l - 1], 4 * H0, H0);
  }
  gj(G_WIHE, 0, 0, g.w_ih, E0 + D0, 0, 4 * H0, E0, H0, 0, 0);
  gj(G_WIHZ, 0, 0, g.w_ih, E0 + D0, E0, 4 * H0, D0, H0, 0, 0);
  gj(G_WA, 0, 0, g.enc_att, D0, 0, A0, D0, 0, 0, 0);
  gj(G_WINIT, 0, 0, g.init_w, E0, 0, 2 * nl * H0, E0, 0, 0, 0);
  gj(G_WFACT, 0, 0, g.fact_w, D0, 0, E0, D0, 0, 0, 0);
  if (g.out_b) {
    const int nch = (M + CS_ROWS - 1) / CS_ROWS;
    f.add(ws + p.cs_off[C_BO], nch, V, 0, 0, g.out_b, 0, 0, 1, V0, 0, 1, 0);
  }
  cj(C_DY, A, g.beta_b, D0, 0);
  cj(C_DY, A + D, g.b_ih, 4 * H0, H0);
  cj(C_DY, A + D, g.b_hh, 4 * H0, H0);
  cj(C_WF, 0, g.f_att, A0, 0);
  cj(C_INIT, 0, g.init_b, 2 * nl * H0, 0);
  cj(C_FACT, 0, g.fact_b, E0, 0);
  SAT_REQUIRE(f.ok, "sat_train_param_grads: finalize job table overflow");
  if (f.blocks > 0) {
    SAT_PROF(6, st);
    param_grads_finalize_kernel<<<f.blocks, 256, 0, st>>>(f.t, b.gscale);
    SAT_PROF(6, st);
    SAT_COUNT_LAUNCH();
    SAT_LAUNCH_OK();
  }
  SAT_PROF(7, st);
  return 0;
}

}  // namespace

extern "C" {

int sat_linear_nt(const void* A, int64_t lda, const void* Bm, int64_t ldb, float* C, int64_t ldc, int32_t Krows, int32_t N1, int32_t N2,
                  int32_t dtype, int32_t use_tc, int32_t splitk, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SAT_REQUIRE(A && Bm && C, "sat_linear_nt: NULL operand");
  SAT_REQUIRE(splitk >= 1 && splitk <= 64, "sat_linear_nt: splitk %d out of range", splitk);
  SAT_REQUIRE(N1 % 4 == 0 && N2 % 4 == 0 && ldc % 4 == 0, "sat_linear_nt: N1, N2, ldc must be multiples of 4");
  SatNoPdlScope no_pdl;               // caller-supplied operands
  EpiStore<float> epi{C, ldc, nullptr, nullptr, 0, (int64_t)N1 * ldc};
  if (dtype == SAT_F32) return launch_gemm_nt_simt<float, float, EpiStore<float>>((const float*)A, lda, (const float*)Bm, ldb, Krows, N1, N2, epi, st, splitk);
  SAT_REQUIRE(dtype == SAT_BF16, "sat_linear_nt: unknown dtype %d", dtype);
  if (use_tc && tc::nt_operands_ok(A, lda, Bm, ldb))
    return tc::launch_nt<EpiStore<float>>((const bf16*)A, lda, (const bf16*)Bm, ldb, Krows, N1, N2, epi, st, splitk);
  return launch_gemm_nt_simt<bf16, bf16, EpiStore<float>>((const bf16*)A, lda, (const bf16*)Bm, ldb, Krows, N1, N2, epi, st, splitk);
}

int64_t sat_param_grads_workspace_bytes(const SatDims* d) {
  if (d == nullptr) return -1;
  return make_plan(*d).total_floats * (int64_t)sizeof(float) + 256;
}

int sat_train_param_grads(const SatDims* d, const SatTrainBuffers* b, const SatParamGrads* g, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  SAT_REQUIRE(d && b && g && workspace, "sat_train_param_grads: NULL argument");
  SAT_REQUIRE(d->dtype == SAT_F32 || d->dtype == SAT_BF16, "unknown dtype %d", d->dtype);
  SAT_REQUIRE(workspace_bytes >= sat_param_grads_workspace_bytes(d), "sat_train_param_grads: workspace of %lld bytes, need %lld",
              (long long)workspace_bytes, (long long)sat_param_grads_workspace_bytes(d));
  SAT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sat_train_param_grads: workspace must be 256-byte aligned");
  SAT_REQUIRE(b->dlogits && b->Xo && b->dpre && b->Hs && b->Z && b->DY && b->Xe && b->GZ && b->dP && b->ann && b->d_init_out && b->f1 &&
                  b->df1 && b->meanv && b->dwf_part && b->tok && b->dXe && b->gscale,
              "sat_train_param_grads: NULL training buffer (run sat_train_forward and sat_train_backward first)");
  SAT_REQUIRE(!g->weight_tying || g->embedding, "sat_train_param_grads: weight tying needs the embedding gradient");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == SAT_F32) return param_grads_impl<float>(*d, *b, *g, (float*)workspace, st);
  return param_grads_impl<bf16>(*d, *b, *g, (float*)workspace, st);
}

}  // extern "C"

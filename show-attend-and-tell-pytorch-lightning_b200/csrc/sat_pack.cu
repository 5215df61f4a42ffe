// sat_pack_weights: one kernel that converts the reference-named fp32 master parameters into every
// packed device layout of SatWeights (operand dtype, gate-interleaved LSTM rows, concatenated projections, transposed
// copies for the backward GEMMs).  Runs once per optimizer step (weights changed) instead of ~60 small framework ops.
#include "sat_common.cuh"

namespace {

struct PackJob {
  const float* src;     // [rows, cols] with leading dimension src_ld
  const float* src2;    // optional second addend (bias_ih + bias_hh)
  void* dst;
  int64_t src_ld, dst_ld;
  int rows, cols;
  int dst_r0, dst_c0;   // offset of the block inside dst (applied after the optional transpose)
  int inter_h;          // > 0: source row r goes to row 4*(r % H) + r / H (gate interleave, H = inter_h)
  int transpose;        // dst[c][r'] instead of dst[r'][c]
  int dst_f32;          // destination is fp32 (biases / w_f) instead of the operand dtype
  int tile0;            // first 32x32 tile of this job in the grid
};

constexpr int MAXJOBS = 48;
struct PackTable {
  PackJob job[MAXJOBS];
  int njobs;
};

template <typename TS>
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackTable tab) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < tab.njobs && (int)blockIdx.x >= tab.job[j + 1].tile0) ++j;
  const PackJob& J = tab.job[j];
  const int tiles_c = (J.cols + 31) / 32;
  const int tl = blockIdx.x - J.tile0;
  const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.0f;
    if (r < J.rows && c < J.cols) {
      v = J.src[(int64_t)r * J.src_ld + c];
      if (J.src2) v += J.src2[(int64_t)r * J.src_ld + c];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (!J.transpose) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      if (r < J.rows && c < J.cols) {
        const int rr = J.inter_h > 0 ? 4 * (r % J.inter_h) + r / J.inter_h : r;
        const int64_t o = (int64_t)(J.dst_r0 + rr) * J.dst_ld + J.dst_c0 + c;
        if (J.dst_f32) reinterpret_cast<float*>(J.dst)[o] = tile[i][tx];
        else reinterpret_cast<TS*>(J.dst)[o] = from_f<TS>(tile[i][tx]);
      }
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;                          // coalesced along the transposed destination row
      if (r < J.rows && c < J.cols) {
        const int rr = J.inter_h > 0 ? 4 * (r % J.inter_h) + r / J.inter_h : r;
        const int64_t o = (int64_t)(J.dst_r0 + c) * J.dst_ld + J.dst_c0 + rr;
        if (J.dst_f32) reinterpret_cast<float*>(J.dst)[o] = tile[tx][i];
        else reinterpret_cast<TS*>(J.dst)[o] = from_f<TS>(tile[tx][i]);
      }
    }
  }
}

struct Builder {
  PackTable t{};
  int tiles = 0;
  bool ok = true;
  void add(const float* src, const float* src2, const void* dst, int rows, int cols, int64_t src_ld, int64_t dst_ld, int r0, int c0,
           int inter_h, int transpose, int f32) {
    if (src == nullptr || dst == nullptr) return;
    if (t.njobs >= MAXJOBS) { ok = false; return; }
    PackJob& j = t.job[t.njobs++];
    j.src = src; j.src2 = src2; j.dst = const_cast<void*>(dst);
    j.src_ld = src_ld; j.dst_ld = dst_ld; j.rows = rows; j.cols = cols; j.dst_r0 = r0; j.dst_c0 = c0;
    j.inter_h = inter_h; j.transpose = transpose; j.dst_f32 = f32; j.tile0 = tiles;
    tiles += ((rows + 31) / 32) * ((cols + 31) / 32);
  }
};

}  // namespace

extern "C" int sat_pack_weights(const SatDims* d, const SatMasterWeights* m, const SatWeights* w, void* stream) {
  SAT_REQUIRE(d && m && w, "sat_pack_weights: NULL struct");
  SAT_REQUIRE(d->dtype == SAT_F32 || d->dtype == SAT_BF16, "unknown dtype %d", d->dtype);
  // storage (padded) dims address the destinations, the module's true dims (D0 ..) the fp32 sources; the destinations'
  // padding is zero-filled once by the owner of the buffers (and -inf for the padded entries of the vocabulary bias)
  const int D = d->D, A = d->A, E = d->E, H = d->H, V = d->V;
  const int D0 = d->D0 ? d->D0 : D, A0 = d->A0 ? d->A0 : A, E0 = d->E0 ? d->E0 : E, H0 = d->H0 ? d->H0 : H, V0 = d->V0 ? d->V0 : V;
  SAT_REQUIRE(D0 <= D && A0 <= A && E0 <= E && H0 <= H && V0 <= V, "sat_pack_weights: true dims exceed the storage dims");
  const int NH3 = A + D + 4 * H, NH4 = NH3 + E;
  const int nl = d->layers > 1 ? d->layers : 1;
  SAT_REQUIRE(nl <= SAT_MAX_LAYERS, "decoder_layers %d > %d", nl, SAT_MAX_LAYERS);
  SAT_REQUIRE(m->embedding && m->w_ih && m->w_hh && m->b_ih && m->b_hh && m->enc_att && m->dec_att && m->f_att && m->beta_w &&
                  m->beta_b && m->out_hidden && m->out_w && m->fact_w && m->fact_b && m->init_w && m->init_b,
              "sat_pack_weights: missing master parameter");
  Builder b;
  // forward layouts                     rows    cols  src_ld   dst_ld  r0       c0 inter tr f32
  b.add(m->enc_att, nullptr, w->Wa,      A0,     D0,   D0,      D,      0,       0, 0,  0, 0);
  b.add(m->dec_att, nullptr, w->Whcat,   A0,     H0,   H0,      H,      0,       0, 0,  0, 0);
  b.add(m->beta_w, nullptr, w->Whcat,    D0,     H0,   H0,      H,      A,       0, 0,  0, 0);
  b.add(m->w_hh, nullptr, w->Whcat,      4 * H0, H0,   H0,      H,      A + D,   0, H0, 0, 0);
  b.add(m->out_hidden, nullptr, w->Whcat, E0,    H0,   H0,      H,      NH3,     0, 0,  0, 0);
  b.add(m->beta_b, nullptr, w->bhcat,    1,      D0,   D0,      NH4,    0,       A, 0,  0, 1);
  b.add(m->w_ih + E0, nullptr, w->Wihz,  4 * H0, D0,   E0 + D0, D,      0,       0, H0, 0, 0);
  b.add(m->w_ih, nullptr, w->Wihe,       4 * H0, E0,   E0 + D0, E,      0,       0, H0, 0, 0);
  b.add(m->b_ih, m->b_hh, w->bg,         4 * H0, 1,    1,       1,      0,       0, H0, 0, 1);
  b.add(m->out_hidden, nullptr, w->Whozo, E0,    H0,   H0,      H + D,  0,       0, 0,  0, 0);
  b.add(m->out_context, nullptr, w->Whozo, E0,   D0,   D0,      H + D,  0,       H, 0,  0, 0);
  b.add(m->out_w, nullptr, w->Wo,        V0,     E0,   E0,      E,      0,       0, 0,  0, 0);
  b.add(m->out_b, nullptr, w->bo,        1,      V0,   V0,      V,      0,       0, 0,  0, 1);
  b.add(m->f_att, nullptr, w->wf,        1,      A0,   A0,      A,      0,       0, 0,  0, 1);
  if (w->Emb != w->Wo) b.add(m->embedding, nullptr, w->Emb, V0, E0, E0, E, 0, 0, 0, 0, 0);
  b.add(m->fact_w, nullptr, w->Wfact,    E0,     D0,   D0,      D,      0,       0, 0,  0, 0);
  b.add(m->fact_b, nullptr, w->bfact,    1,      E0,   E0,      E,      0,       0, 0,  0, 1);
  b.add(m->init_w, nullptr, w->Winit,    2 * nl * H0, E0, E0,   E,      0,       0, 0,  0, 0);
  b.add(m->init_b, nullptr, w->binit,    1,      2 * nl * H0, 2 * nl * H0, 2 * nl * H, 0, 0, 0, 0, 1);
  for (int l = 1; l < nl; ++l) {      // stacked layers: [W_ih_l | W_hh_l] gate-interleaved, summed biases, transposed copy
    SAT_REQUIRE(m->w_ih_l[l - 1] && m->w_hh_l[l - 1] && m->b_ih_l[l - 1] && m->b_hh_l[l - 1], "sat_pack_weights: parameters of LSTM layer %d missing", l);
    b.add(m->w_ih_l[l - 1], nullptr, w->Wl[l - 1], 4 * H0, H0, H0, 2 * H, 0, 0, H0, 0, 0);
    b.add(m->w_hh_l[l - 1], nullptr, w->Wl[l - 1], 4 * H0, H0, H0, 2 * H, 0, H, H0, 0, 0);
    b.add(m->b_ih_l[l - 1], m->b_hh_l[l - 1], w->bgl[l - 1], 4 * H0, 1, 1, 1, 0, 0, H0, 0, 1);
    b.add(m->w_ih_l[l - 1], nullptr, w->WlT[l - 1], 4 * H0, H0, H0, 4 * H, 0, 0, H0, 1, 0);
    b.add(m->w_hh_l[l - 1], nullptr, w->WlT[l - 1], 4 * H0, H0, H0, 4 * H, H, 0, H0, 1, 0);
  }
  // transposed copies for the backward GEMMs (skipped when the destination pointers are NULL)
  b.add(m->out_w, nullptr, w->WoT,       V0,     E0,   E0,      V,      0,       0, 0,  1, 0);
  b.add(m->out_hidden, nullptr, w->WhozoT, E0,   H0,   H0,      E,      0,       0, 0,  1, 0);
  b.add(m->out_context, nullptr, w->WhozoT, E0,  D0,   D0,      E,      H,       0, 0,  1, 0);
  b.add(m->w_ih + E0, nullptr, w->WihzT, 4 * H0, D0,   E0 + D0, 4 * H,  0,       0, H0, 1, 0);
  b.add(m->w_ih, nullptr, w->WiheT,      4 * H0, E0,   E0 + D0, 4 * H,  0,       0, H0, 1, 0);
  b.add(m->dec_att, nullptr, w->WhcatT,  A0,     H0,   H0,      NH3,    0,       0, 0,  1, 0);
  b.add(m->beta_w, nullptr, w->WhcatT,   D0,     H0,   H0,      NH3,    0,       A, 0,  1, 0);
  b.add(m->w_hh, nullptr, w->WhcatT,     4 * H0, H0,   H0,      NH3,    0,       A + D, H0, 1, 0);
  b.add(m->enc_att, nullptr, w->WaT,     A0,     D0,   D0,      A,      0,       0, 0,  1, 0);
  b.add(m->init_w, nullptr, w->WinitT,   2 * nl * H0, E0, E0,   2 * nl * H, 0,   0, 0,  1, 0);
  b.add(m->fact_w, nullptr, w->WfactT,   E0,     D0,   D0,      E,      0,       0, 0,  1, 0);
  SAT_REQUIRE(b.ok, "sat_pack_weights: job table overflow");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == SAT_F32) pack_kernel<float><<<b.tiles, 256, 0, st>>>(b.t);
  else pack_kernel<bf16><<<b.tiles, 256, 0, st>>>(b.t);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

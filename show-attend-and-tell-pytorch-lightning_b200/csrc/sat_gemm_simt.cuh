// FFMA (SIMT) GEMM core with fused epilogues:  C[M,N] = epi( sum_seg A_seg[M,K_seg] * W[N, Kbase_seg..]^T ).
//
// This is the fp32 parity path (Blackwell tensor cores have no fp32 MMA; TF32 is ~1e-3) and the
// shape fallback of the bf16 path when a GEMM does not meet the tcgen05/TMA tile constraints.
// Operands are K-major ("TN": nn.Linear weight layout), A may be split into up to three K
// segments living in different buffers (e.g. [h' ; z] for the deep-output projection).
// 64x64x16 CTA tile, 256 threads, 4x4 register micro-tile, register-prefetch double buffering.
#pragma once
#include "sat_common.cuh"

struct GemmOperandA {
  const void* p[3];
  int64_t ld[3];
  int k[3];
  int nseg;
};

static inline GemmOperandA gemm_a1(const void* p, int64_t ld, int k) {
  GemmOperandA a{};
  a.p[0] = p; a.ld[0] = ld; a.k[0] = k; a.nseg = 1;
  return a;
}
static inline GemmOperandA gemm_a2(const void* p0, int64_t ld0, int k0, const void* p1, int64_t ld1, int k1) {
  GemmOperandA a{};
  a.p[0] = p0; a.ld[0] = ld0; a.k[0] = k0;
  a.p[1] = p1; a.ld[1] = ld1; a.k[1] = k1;
  a.nseg = 2;
  return a;
}

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

// Epilogue concept:  __device__ void operator()(int m, int n, const float (&acc)[4]) const
// called once per (row m, 4 consecutive columns n..n+3), n % 4 == 0, m < M, n < N (N % 4 == 0).

template <typename TA, typename TW, typename Epi>
__global__ void __launch_bounds__(SG_THREADS)
gemm_tn_simt_kernel(GemmOperandA A, const TW* __restrict__ W, int64_t ldw, int M, int N, Epi epi) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[2][SG_BK][SG_BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int lrow = tid >> 2;          // 0..63
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // total number of k tiles over all segments
  int ntiles = 0;
  for (int s = 0; s < A.nseg; ++s) ntiles += (A.k[s] + SG_BK - 1) / SG_BK;

  int seg = 0, kt_in_seg = 0, kbase = 0;   // iterator state for the *next tile to load*
  float4 ra, rw;

  auto load_tile = [&]() {
    const int k = kt_in_seg * SG_BK + lk;
    const int am = m0 + lrow, wn = n0 + lrow;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < A.k[seg]) {
      if (am < M) ra = ld4(reinterpret_cast<const TA*>(A.p[seg]) + (int64_t)am * A.ld[seg] + k);
      if (wn < N) rw = ld4(W + (int64_t)wn * ldw + kbase + k);
    }
    // advance iterator
    ++kt_in_seg;
    if (kt_in_seg * SG_BK >= A.k[seg]) {
      kbase += A.k[seg];
      ++seg;
      kt_in_seg = 0;
    }
  };
  auto store_tile = [&](int buf) {
    As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y; As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
    Ws[buf][lk + 0][lrow] = rw.x; Ws[buf][lk + 1][lrow] = rw.y; Ws[buf][lk + 2][lrow] = rw.z; Ws[buf][lk + 3][lrow] = rw.w;
  };

  load_tile();
  store_tile(0);
  __syncthreads();
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) load_tile();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (it + 1 < ntiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  const int n = n0 + tx * 4;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M) epi(m, n, acc[i]);
    }
  }
}

template <typename TA, typename TW, typename Epi>
static int launch_gemm_tn_simt(const GemmOperandA& A, const TW* W, int64_t ldw, int M, int N, const Epi& epi,
                               cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  int ktot = 0;
  for (int s = 0; s < A.nseg; ++s) {
    SAT_REQUIRE(A.k[s] % 4 == 0 && A.ld[s] % 4 == 0, "gemm: K segment %d (k=%d ld=%lld) must be a multiple of 4", s, A.k[s],
                (long long)A.ld[s]);
    ktot += A.k[s];
  }
  SAT_REQUIRE(N % 4 == 0 && ldw % 4 == 0 && ktot > 0, "gemm: N=%d ldw=%lld must be multiples of 4", N, (long long)ldw);
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  gemm_tn_simt_kernel<TA, TW, Epi><<<grid, SG_THREADS, 0, stream>>>(A, W, ldw, M, N, epi);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

// ---- "NT" SIMT core: C[N1,N2] = sum_k A[k,n1] * B[k,n2] (weight gradients dW = dY^T X; fp32 parity mode and shapes
// off the TMA grid).  Both operands are read along their contiguous dimension, so the shared tiles need no transpose.
// gridDim.z splits the k range [0, Krows); partial z goes through the epilogue's zstride.
template <typename TA, typename TB, typename Epi>
__global__ void __launch_bounds__(SG_THREADS)
gemm_nt_simt_kernel(const TA* __restrict__ A, int64_t lda, const TB* __restrict__ B, int64_t ldb, int Krows, int N1, int N2, Epi epi) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int lk = tid >> 4;             // k row inside the tile, 0..15
  const int lc = (tid & 15) * 4;       // column inside the tile, 0..60
  const int ty = tid >> 4, tx = tid & 15;
  const int kper = (((Krows + gridDim.z - 1) / gridDim.z) + SG_BK - 1) / SG_BK * SG_BK;
  const int kbeg = blockIdx.z * kper, kend = min(Krows, kbeg + kper);
  const int ntiles = kend > kbeg ? (kend - kbeg + SG_BK - 1) / SG_BK : 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float4 ra, rb;
  auto load_tile = [&](int it) {
    const int k = kbeg + it * SG_BK + lk;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < kend) {
      if (m0 + lc < N1) ra = ld4(A + (int64_t)k * lda + m0 + lc);
      if (n0 + lc < N2) rb = ld4(B + (int64_t)k * ldb + n0 + lc);
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lk][lc]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lk][lc]) = rb;
  };
  if (ntiles > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) load_tile(it + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (it + 1 < ntiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
  const int n = n0 + tx * 4;
  if (n < N2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < N1) epi(m, n, acc[i]);
    }
  }
}

template <typename TA, typename TB, typename Epi>
static int launch_gemm_nt_simt(const TA* A, int64_t lda, const TB* B, int64_t ldb, int Krows, int N1, int N2, const Epi& epi,
                               cudaStream_t stream, int splitk) {
  if (N1 <= 0 || N2 <= 0) return 0;
  SAT_REQUIRE(N1 % 4 == 0 && N2 % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, "gemm_nt: N1=%d N2=%d lda=%lld ldb=%lld must be multiples of 4", N1,
              N2, (long long)lda, (long long)ldb);
  dim3 grid((N2 + SG_BN - 1) / SG_BN, (N1 + SG_BM - 1) / SG_BM, splitk);
  gemm_nt_simt_kernel<TA, TB, Epi><<<grid, SG_THREADS, 0, stream>>>(A, lda, B, ldb, Krows, N1, N2, epi);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

// ---- epilogues -----------------------------------------------------------------------------

// Epilogue concept (two-phase so that the tensor-core core can issue the global loads of several rows before
// any dependent math -- the epilogue is latency-bound otherwise):
//   Ctx  load(int m, int n) const                      -- global reads for (row m, columns n..n+3)
//   void apply(int m, int n, const float (&acc)[4], const Ctx&) const
//   void operator()(m, n, acc)                         -- load + apply (SIMT core)

// C = acc + bias (+ residual of type TR)
template <typename TO, typename TR = float>
struct EpiStore {
  TO* C;
  int64_t ldc;
  const float* bias;        // [N] or nullptr
  const TR* res;            // residual [M, ldr] or nullptr
  int64_t ldr;
  int64_t zstride;          // split-K: CTA z writes its partial at C + z*zstride (0 when unused)
  struct Ctx { float4 b, r; };
  __device__ __forceinline__ Ctx load(int m, int n) const {
    Ctx c;
    c.b = bias ? ld4(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    c.r = res ? ld4(res + (int64_t)m * ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    return c;
  }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx& c) const {
    st4(C + (int64_t)blockIdx.z * zstride + (int64_t)m * ldc + n,
        make_float4(acc[0] + c.b.x + c.r.x, acc[1] + c.b.y + c.r.y, acc[2] + c.b.z + c.r.z, acc[3] + c.b.w + c.r.w));
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, load(m, n)); }
};

// C (fp32) = acc, plus an operand-dtype copy C16 (feeds a following tensor-core GEMM)
template <typename TS>
struct EpiStoreDual {
  float* C; TS* C16; int64_t ldc;
  struct Ctx {};
  __device__ __forceinline__ Ctx load(int, int) const { return Ctx{}; }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx&) const {
    const float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
    st4(C + (int64_t)m * ldc + n, v);
    st4(C16 + (int64_t)m * ldc + n, v);
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, Ctx{}); }
};

// d_ann[b,l,:] = acc (= dP[b,l,:] * Wa) + sum_t alpha[b,t,l] * dZ[t,b,:] + dmean[img,:] * mean_scale
// rows m = b*L + l   (d_a += alpha (x) dz ; d_a += dP W_a ; mean path)
template <typename TS>
struct EpiDAnn {
  TS* d_ann; const float* alphas; const TS* dZ; const float* dmean;
  int B, T, L, D, ncap; float mean_scale;
  struct Ctx { float4 dm; };
  __device__ __forceinline__ Ctx load(int m, int n) const {
    Ctx c;
    c.dm = ld4(dmean + (int64_t)((m / L) / ncap) * D + n);
    return c;
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, load(m, n)); }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx& c) const {
    const int b = m / L, l = m - b * L;
    const float4 dm = c.dm;
    float4 v = make_float4(acc[0] + dm.x * mean_scale, acc[1] + dm.y * mean_scale, acc[2] + dm.z * mean_scale,
                           acc[3] + dm.w * mean_scale);
    const float* al = alphas + ((int64_t)b * T) * L + l;
    for (int t = 0; t < T; ++t) {
      const float a = al[(int64_t)t * L];
      const float4 dz = ld4(dZ + ((int64_t)t * B + b) * D + n);
      v.x = fmaf(a, dz.x, v.x); v.y = fmaf(a, dz.y, v.y); v.z = fmaf(a, dz.z, v.z); v.w = fmaf(a, dz.w, v.w);
    }
    st4(d_ann + (int64_t)m * D + n, v);
  }
};

// LSTM cell (torch.nn.LSTM, gate order i,f,g,o; model.py:544): the GEMM computes (beta*z) * Wihz^T over
// gate-interleaved columns, so acc[0..3] are the i,f,g,o pre-activations of hidden unit j = n/4 once
// Gx (embedding part + biases) and the W_hh*h part (from the h-projection GEMM) are added.
// Rows with t >= lens[m] are frozen (model.py:512,544: only incomplete rows are updated).
template <typename TS, bool kExact>
struct EpiLstm {
  const float* Gx;  int64_t ldgx;      // row m -> Gx + m*ldgx (already offset to step t)
  const float* Gh;  int64_t ldgh;      // W_hh h part: hp + (A+D), row stride ldgh
  const TS* h_prev;
  const float* c_prev;                 // [.., H] row stride ldst_c
  TS* h_next; float* c_next;           // row stride ldst
  int64_t ldst_h, ldst_c;
  TS* gates; int64_t ldgates;          // post-activation gates for backward (row m -> gates + m*ldgates)
  const int32_t* lens; int t;
  const int32_t* gx_row;               // optional: Gx row index per m (decode: token id into the [V,4H] table)
  struct Ctx { float4 gx, gh; float cp; bool active; };
  __device__ __forceinline__ Ctx load(int m, int n) const {
    Ctx c;
    c.active = t < lens[m];
    c.cp = c_prev[(int64_t)m * ldst_c + (n >> 2)];
    if (c.active) {
      c.gx = ld4(Gx + (int64_t)(gx_row ? gx_row[m] : m) * ldgx + n);      // ldgx = 0: one bias row for every m (stacked layers)
      c.gh = Gh ? ld4(Gh + (int64_t)m * ldgh + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      c.gx = c.gh = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return c;
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, load(m, n)); }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx& c) const {
    const int j = n >> 2;
    const float cp = c.cp;
    if (!c.active) {
      h_next[(int64_t)m * ldst_h + j] = h_prev[(int64_t)m * ldst_h + j];
      c_next[(int64_t)m * ldst_c + j] = cp;
      if (gates) st4(gates + (int64_t)m * ldgates + n, make_float4(0.f, 0.f, 0.f, 0.f));
      return;
    }
    const float4 gx = c.gx, gh = c.gh;
    const float gi = sat_sigmoid<kExact>(acc[0] + gx.x + gh.x);
    const float gf = sat_sigmoid<kExact>(acc[1] + gx.y + gh.y);
    const float gg = sat_tanh<kExact>(acc[2] + gx.z + gh.z);
    const float go = sat_sigmoid<kExact>(acc[3] + gx.w + gh.w);
    const float cn = gf * cp + gi * gg;
    const float hn = go * sat_tanh<kExact>(cn);
    c_next[(int64_t)m * ldst_c + j] = cn;
    h_next[(int64_t)m * ldst_h + j] = from_f<TS>(hn);
    if (gates) st4(gates + (int64_t)m * ldgates + n, make_float4(gi, gf, gg, go));
  }
};

// deep output pre-activation: Xo = tanh(acc + Xe)   (model.py:127)
template <typename TS, bool kExact>
struct EpiTanhAdd {
  const TS* Xe; TS* Xo; int64_t ld;
  const int32_t* xe_row;               // optional: Xe row index per m (decode: token id into the embedding table)
  int plain;                           // DeepOutput(deep=False): Xo = acc (model.py:129)
  float drop_p; uint64_t seed; int64_t m_base;   // dropout on the activations before the vocabulary layer (model.py:130)
  struct Ctx { float4 x; };
  __device__ __forceinline__ Ctx load(int m, int n) const {
    Ctx c;
    c.x = plain ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(Xe + (int64_t)(xe_row ? xe_row[m] : m) * ld + n);
    return c;
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, load(m, n)); }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx& c) const {
    const float4 x = c.x;
    float4 o;
    if (plain) o = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
      o = make_float4(sat_tanh<kExact>(acc[0] + x.x), sat_tanh<kExact>(acc[1] + x.y), sat_tanh<kExact>(acc[2] + x.z),
                      sat_tanh<kExact>(acc[3] + x.w));
    if (drop_p > 0.0f) {
      const uint64_t i0 = (uint64_t)(m_base + m) * ld + n;
      o.x *= sat_dropout_scale(drop_p, seed, 3u, i0);     o.y *= sat_dropout_scale(drop_p, seed, 3u, i0 + 1);
      o.z *= sat_dropout_scale(drop_p, seed, 3u, i0 + 2); o.w *= sat_dropout_scale(drop_p, seed, 3u, i0 + 3);
    }
    st4(Xo + (int64_t)m * ld + n, o);
  }
};

// dpre = g * acc * (1 - Xo^2)
template <typename TS>
struct EpiDpre {
  const TS* Xo; TS* dpre; int64_t ld; const float* gscale;
  int plain;                           // DeepOutput(deep=False): no tanh derivative
  float drop_p; uint64_t seed;         // Xo holds the DROPPED activations: x = Xo * (1-p) where kept; grad gets the mask
  struct Ctx { float4 x; float g; };
  __device__ __forceinline__ Ctx load(int m, int n) const {
    Ctx c;
    c.g = gscale ? *gscale : 1.0f;
    c.x = ld4(Xo + (int64_t)m * ld + n);
    return c;
  }
  __device__ __forceinline__ void operator()(int m, int n, const float (&acc)[4]) const { apply(m, n, acc, load(m, n)); }
  __device__ __forceinline__ void apply(int m, int n, const float (&acc)[4], const Ctx& c) const {
    const float g = c.g;
    float4 x = plain ? make_float4(0.f, 0.f, 0.f, 0.f) : c.x;
    float4 k4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (drop_p > 0.0f) {
      const uint64_t i0 = (uint64_t)m * ld + n;
      k4 = make_float4(sat_dropout_scale(drop_p, seed, 3u, i0), sat_dropout_scale(drop_p, seed, 3u, i0 + 1),
                       sat_dropout_scale(drop_p, seed, 3u, i0 + 2), sat_dropout_scale(drop_p, seed, 3u, i0 + 3));
      const float keep = 1.0f - drop_p;
      x.x *= keep; x.y *= keep; x.z *= keep; x.w *= keep;      // undo the forward scaling where kept (irrelevant where dropped)
    }
    st4(dpre + (int64_t)m * ld + n, make_float4(g * k4.x * acc[0] * (1.f - x.x * x.x), g * k4.y * acc[1] * (1.f - x.y * x.y),
                                                 g * k4.z * acc[2] * (1.f - x.z * x.z), g * k4.w * acc[3] * (1.f - x.w * x.w)));
  }
};

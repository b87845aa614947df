// fp32 GEMM core on CUDA cores with pluggable operand loaders and epilogues.
//
// One register-tiled kernel family serves every dense contraction of the fp32 path:
//   * nn.Linear forward / data-gradient        (PlainLoader)           model.py:15,23
//   * Conv2d as implicit GEMM (im2col on load) (Im2colLoader)          Unet.py:16,19  model.py:228-241
//   * ConvTranspose2d k2 s2                    (ConvtEpilogue / ConvtLoader)  Unet.py:53
//   * weight gradients: C = A^T B over the row dimension, deterministic split (gemm_tn)
// Tile: 128 x BN x 8, 256 threads, 8 x (BN/16) accumulators per thread, double-buffered smem,
// 128-bit loads wherever the operand layout allows (VEC template flags).
#pragma once
#include "tm_common.cuh"

namespace tmk {

// ---------------------------------------------------------------------------------------------
// operand loaders: element (r, c) of a logical row-major matrix
// ---------------------------------------------------------------------------------------------
struct PlainLoader {
  const float* A;
  int64_t ld;
  const int32_t* rows;  // optional gather
  struct Row {
    const float* p;
    bool valid;
  };
  __device__ __forceinline__ Row row(int64_t r, int64_t R) const {
    Row o;
    o.valid = r < R;
    int64_t rr = o.valid ? (rows ? (int64_t)rows[r] : r) : 0;
    o.p = A + rr * ld;
    return o;
  }
  template <bool VEC>
  __device__ __forceinline__ void load4(const Row& r, int64_t c, int64_t C, float (&v)[4]) const {
    if (VEC) {
      if (r.valid && c < C) {
        float4 t = ld4(r.p + c);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = v[1] = v[2] = v[3] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (r.valid && c + i < C) ? r.p[c + i] : 0.f;
    }
  }
};

// row = output pixel (b,y,x); column = (tap, ci) with ci fastest.  NHWC input, stride 1, pad ks/2.
struct Im2colLoader {
  const float* X;
  int64_t ldx;  // pixel stride
  int H, W, Cin, ks, pad;
  struct Row {
    int b, y, x;
    bool valid;
  };
  __device__ __forceinline__ Row row(int64_t r, int64_t R) const {
    Row o;
    o.valid = r < R;
    int64_t rr = o.valid ? r : 0;
    int hw = H * W;
    o.b = (int)(rr / hw);
    int rem = (int)(rr - (int64_t)o.b * hw);
    o.y = rem / W;
    o.x = rem - o.y * W;
    return o;
  }
  __device__ __forceinline__ const float* at(const Row& r, int tap, bool& inb) const {
    int ty = tap / ks, tx = tap - ty * ks;
    int yy = r.y + ty - pad, xx = r.x + tx - pad;
    inb = r.valid && yy >= 0 && yy < H && xx >= 0 && xx < W;
    return X + (((int64_t)r.b * H + yy) * W + xx) * ldx;
  }
  template <bool VEC>
  __device__ __forceinline__ void load4(const Row& r, int64_t c, int64_t C, float (&v)[4]) const {
    if (VEC) {  // Cin % 4 == 0: the 4 columns share a tap
      v[0] = v[1] = v[2] = v[3] = 0.f;
      if (c < C) {
        int tap = (int)(c / Cin), ci = (int)(c - (int64_t)tap * Cin);
        bool inb;
        const float* p = at(r, tap, inb);
        if (inb) {
          float4 t = ld4(p + ci);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = 0.f;
        if (c + i < C) {
          int tap = (int)((c + i) / Cin), ci = (int)((c + i) - (int64_t)tap * Cin);
          bool inb;
          const float* p = at(r, tap, inb);
          if (inb) v[i] = p[ci];
        }
      }
    }
  }
};

// row = input pixel (b,y,x) of a k2 s2 transposed conv; column = (dy,dx,co), co fastest:
// element = Y[b, 2y+dy+oy, 2x+dx+ox, co] of the upsampled tensor (gradient gather).
struct ConvtLoader {
  const float* Y;
  int64_t ldy;
  int H, W, Cout, Hy, Wy, oy, ox;
  struct Row {
    int b, y, x;
    bool valid;
  };
  __device__ __forceinline__ Row row(int64_t r, int64_t R) const {
    Row o;
    o.valid = r < R;
    int64_t rr = o.valid ? r : 0;
    int hw = H * W;
    o.b = (int)(rr / hw);
    int rem = (int)(rr - (int64_t)o.b * hw);
    o.y = rem / W;
    o.x = rem - o.y * W;
    return o;
  }
  __device__ __forceinline__ const float* at(const Row& r, int q) const {
    int dy = q >> 1, dx = q & 1;
    return Y + (((int64_t)r.b * Hy + 2 * r.y + dy + oy) * Wy + 2 * r.x + dx + ox) * ldy;
  }
  template <bool VEC>
  __device__ __forceinline__ void load4(const Row& r, int64_t c, int64_t C, float (&v)[4]) const {
    if (VEC) {
      v[0] = v[1] = v[2] = v[3] = 0.f;
      if (r.valid && c < C) {
        int q = (int)(c / Cout), co = (int)(c - (int64_t)q * Cout);
        float4 t = ld4(at(r, q) + co);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = 0.f;
        if (r.valid && c + i < C) {
          int q = (int)((c + i) / Cout), co = (int)((c + i) - (int64_t)q * Cout);
          v[i] = at(r, q)[co];
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
struct PlainEpilogue {
  float* C;
  int64_t ldc;
  const int32_t* c_rows;
  const float* bias;
  const float* mask;
  int64_t ldmask;
  int flags;
  struct Row {
    int64_t r;
  };
  __device__ __forceinline__ Row row(int64_t m) const { return Row{c_rows ? (int64_t)c_rows[m] : m}; }
  __device__ __forceinline__ void store(const Row& rw, int64_t n, float v) const {
    if (flags & TM_EPI_BIAS) v += bias[n];
    float* p = C + rw.r * ldc + n;
    if (flags & TM_EPI_ACCUM) v += *p;
    if (flags & TM_EPI_RELU) v = fmaxf(v, 0.f);
    if (flags & TM_EPI_MASK) v = (mask[rw.r * ldmask + n] > 0.f) ? v : 0.f;
    *p = v;
  }
};

// scatter of a k2 s2 transposed conv: row = input pixel, column = (dy,dx,co)
struct ConvtEpilogue {
  float* Y;
  int64_t ldy;
  const float* bias;
  int H, W, Cout, Hy, Wy, oy, ox;
  struct Row {
    int b, y, x;
  };
  __device__ __forceinline__ Row row(int64_t m) const {
    Row o;
    int hw = H * W;
    o.b = (int)(m / hw);
    int rem = (int)(m - (int64_t)o.b * hw);
    o.y = rem / W;
    o.x = rem - o.y * W;
    return o;
  }
  __device__ __forceinline__ void store(const Row& r, int64_t n, float v) const {
    int q = (int)(n / Cout), co = (int)(n - (int64_t)q * Cout);
    int dy = q >> 1, dx = q & 1;
    if (bias) v += bias[co];
    Y[(((int64_t)r.b * Hy + 2 * r.y + dy + oy) * Wy + 2 * r.x + dx + ox) * ldy + co] = v;
  }
};

// ---------------------------------------------------------------------------------------------
// C[M,N] = epi(A[M,K] @ B[K,N]);  B plain row-major
// ---------------------------------------------------------------------------------------------
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 8;
constexpr int GEMM_AS = GEMM_BM + 4;  // padded k-major A tile row

template <int TN>
__device__ __forceinline__ void load_bfrag(const float* bs, int tx, float (&b)[TN]) {
  if constexpr (TN == 8) {
    float4 u = *reinterpret_cast<const float4*>(bs + tx * 4);
    float4 w = *reinterpret_cast<const float4*>(bs + 64 + tx * 4);
    b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w; b[4] = w.x; b[5] = w.y; b[6] = w.z; b[7] = w.w;
  } else if constexpr (TN == 4) {
    float4 u = *reinterpret_cast<const float4*>(bs + tx * 4);
    b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w;
  } else if constexpr (TN == 2) {
    float2 u = *reinterpret_cast<const float2*>(bs + tx * 2);
    b[0] = u.x; b[1] = u.y;
  } else {
    b[0] = bs[tx];
  }
}
template <int TN>
__device__ __forceinline__ int frag_col(int tx, int j) {
  if constexpr (TN == 8) return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
  if constexpr (TN == 4) return tx * 4 + j;
  if constexpr (TN == 2) return tx * 2 + j;
  return tx;
}

template <class AL, class EP, int BN, bool VECA, bool VECB>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(AL al, const float* __restrict__ B, int64_t ldb, EP ep, int64_t M, int64_t N, int64_t K) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][GEMM_BK][GEMM_AS];
  __shared__ __align__(16) float Bs[2][GEMM_BK][BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * GEMM_BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int arow = tid >> 1, ak = (tid & 1) * 4;
  const typename AL::Row arw = al.row(m0 + arow, M);
  constexpr int B_F4 = GEMM_BK * BN / 4;  // float4 slots in a B tile
  const bool bact = tid < B_F4;
  const int bk = bact ? tid / (BN / 4) : 0;
  const int bc = bact ? (tid % (BN / 4)) * 4 : 0;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ar[4], br[4];
  auto fetch = [&](int64_t k0) {
    al.template load4<VECA>(arw, k0 + ak, K, ar);
    if (bact) {
      const int64_t kk = k0 + bk, nn = n0 + bc;
      if (VECB) {
        if (kk < K && nn < N) {
          float4 t = ld4(B + kk * ldb + nn);
          br[0] = t.x; br[1] = t.y; br[2] = t.z; br[3] = t.w;
        } else {
          br[0] = br[1] = br[2] = br[3] = 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) br[i] = (kk < K && nn + i < N) ? B[kk * ldb + nn + i] : 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][ak + i][arow] = ar[i];
    if (bact) *reinterpret_cast<float4*>(&Bs[buf][bk][bc]) = make_float4(br[0], br[1], br[2], br[3]);
  };

  const int64_t nk = (K + GEMM_BK - 1) / GEMM_BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int cur = (int)(kt & 1);
    if (kt + 1 < nk) fetch((kt + 1) * GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      load_bfrag<TN>(&Bs[cur][kk][0], tx, b);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(cur ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const typename EP::Row rw = ep.row(m);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + frag_col<TN>(tx, j);
      if (n < N) ep.store(rw, n, acc[i][j]);
    }
  }
}

template <class AL, class EP>
int launch_gemm_nn(const AL& al, bool veca, const float* B, int64_t ldb, const EP& ep, int64_t M,
                   int64_t N, int64_t K, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const bool vecb = (ldb % 4 == 0) && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : (N > 16 ? 32 : 16));
  dim3 grid((unsigned)cdiv(M, GEMM_BM), (unsigned)cdiv(N, bn));
#define TM_GEMM_CASE(BN_, VA_, VB_)                                                               \
  gemm_nn_kernel<AL, EP, BN_, VA_, VB_><<<grid, 256, 0, st>>>(al, B, ldb, ep, M, N, K)
#define TM_GEMM_BN(BN_)                                   \
  do {                                                    \
    if (veca && vecb) TM_GEMM_CASE(BN_, true, true);      \
    else if (veca) TM_GEMM_CASE(BN_, true, false);        \
    else if (vecb) TM_GEMM_CASE(BN_, false, true);        \
    else TM_GEMM_CASE(BN_, false, false);                 \
  } while (0)
  switch (bn) {
    case 128: TM_GEMM_BN(128); break;
    case 64: TM_GEMM_BN(64); break;
    case 32: TM_GEMM_BN(32); break;
    default: TM_GEMM_BN(16); break;
  }
#undef TM_GEMM_BN
#undef TM_GEMM_CASE
  return check_launch("gemm_nn");
}

// ---------------------------------------------------------------------------------------------
// P[split][M][N] = sum over the split's rows r of A[r][m] * B[r][n]   (+ column sums of B)
// ---------------------------------------------------------------------------------------------
template <class AL, class BL, int BN, bool VECA, bool VECB>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(AL al, BL bl, float* __restrict__ P, float* __restrict__ PcsA, float* __restrict__ Pcs,
               int64_t M, int64_t N, int64_t R, int64_t rows_per_split) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][GEMM_BK][GEMM_BM];
  __shared__ __align__(16) float Bs[2][GEMM_BK][BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * GEMM_BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
  const int akk = tid >> 5, am = (tid & 31) * 4;  // A tile: 8 rows x 128 cols, one float4 each
  constexpr int B_F4 = GEMM_BK * BN / 4;
  const bool bact = tid < B_F4;
  const int bk = bact ? tid / (BN / 4) : 0;
  const int bc = bact ? (tid % (BN / 4)) * 4 : 0;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float cs = 0.f;   // column sum of B for column n0 + tid (threads < BN of m-tile 0)
  float csa = 0.f;  // column sum of A for column m0 + tid (threads < 128 of n-tile 0)
  const bool do_cs = (Pcs != nullptr) && blockIdx.x == 0 && tid < BN;
  const bool do_csa = (PcsA != nullptr) && blockIdx.y == 0 && tid < GEMM_BM;

  float ar[4], br[4];
  auto fetch = [&](int64_t r0) {
    {
      const int64_t r = r0 + akk;
      typename AL::Row rw = al.row(r, r_end);
      al.template load4<VECA>(rw, m0 + am, M, ar);
    }
    if (bact) {
      const int64_t r = r0 + bk;
      typename BL::Row rw = bl.row(r, r_end);
      bl.template load4<VECB>(rw, n0 + bc, N, br);
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][akk][am]) = make_float4(ar[0], ar[1], ar[2], ar[3]);
    if (bact) *reinterpret_cast<float4*>(&Bs[buf][bk][bc]) = make_float4(br[0], br[1], br[2], br[3]);
  };

  const int64_t nk = (r_end > r_begin) ? (r_end - r_begin + GEMM_BK - 1) / GEMM_BK : 0;
  if (nk > 0) {
    fetch(r_begin);
    stash(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int cur = (int)(kt & 1);
    if (kt + 1 < nk) fetch(r_begin + (kt + 1) * GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      load_bfrag<TN>(&Bs[cur][kk][0], tx, b);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_cs) cs += Bs[cur][kk][tid];
      if (do_csa) csa += As[cur][kk][tid];
    }
    if (kt + 1 < nk) stash(cur ^ 1);
    __syncthreads();
  }
  float* Ps = P + (int64_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + frag_col<TN>(tx, j);
      if (n < N) Ps[m * N + n] = acc[i][j];
    }
  }
  if (do_cs && n0 + tid < N) Pcs[(int64_t)blockIdx.z * N + n0 + tid] = cs;
  if (do_csa && m0 + tid < M) PcsA[(int64_t)blockIdx.z * M + m0 + tid] = csa;
}

__global__ void split_reduce_kernel(const float* __restrict__ P, int64_t count, int splits,
                                    float* __restrict__ C, int64_t N, int64_t ldc, int accumulate);

inline int tn_splits(int64_t M, int64_t N, int64_t R, int bn) {
  int64_t tiles = cdiv(M, GEMM_BM) * cdiv(N, bn);
  int64_t want = cdiv(2 * (int64_t)sm_count(), tiles);
  int64_t cap = cdiv(R, 64);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}
inline int tn_bn(int64_t N) { return N > 64 ? 128 : (N > 32 ? 64 : (N > 16 ? 32 : 16)); }
inline size_t tn_ws_bytes(int64_t M, int64_t N, int64_t R) {
  int s = tn_splits(M, N, R, tn_bn(N));
  return (size_t)s * (size_t)(M * N + N + M) * sizeof(float) + 1024;
}

template <class AL, class BL>
int launch_gemm_tn(const AL& al, bool veca, const BL& bl, bool vecb, int64_t M, int64_t N, int64_t R,
                   float* C, int64_t ldc, float* colsum_a, float* colsum, int accumulate, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int bn = tn_bn(N);
  const int splits = tn_splits(M, N, R, bn);
  if (ws_bytes < tn_ws_bytes(M, N, R)) return fail(TM_EWORKSPACE, "gemm_tn: workspace too small");
  Carver c(ws);
  float* P = c.take<float>((size_t)splits * M * N);
  float* Pcs = colsum ? c.take<float>((size_t)splits * N) : nullptr;
  float* PcsA = colsum_a ? c.take<float>((size_t)splits * M) : nullptr;
  const int64_t rps = cdiv(cdiv(R, splits), GEMM_BK) * GEMM_BK;
  dim3 grid((unsigned)cdiv(M, GEMM_BM), (unsigned)cdiv(N, bn), (unsigned)splits);
#define TM_TN_CASE(BN_, VA_, VB_)                                                                  \
  gemm_tn_kernel<AL, BL, BN_, VA_, VB_><<<grid, 256, 0, st>>>(al, bl, P, PcsA, Pcs, M, N, R, rps)
#define TM_TN_BN(BN_)                                   \
  do {                                                  \
    if (veca && vecb) TM_TN_CASE(BN_, true, true);      \
    else if (veca) TM_TN_CASE(BN_, true, false);        \
    else if (vecb) TM_TN_CASE(BN_, false, true);        \
    else TM_TN_CASE(BN_, false, false);                 \
  } while (0)
  switch (bn) {
    case 128: TM_TN_BN(128); break;
    case 64: TM_TN_BN(64); break;
    case 32: TM_TN_BN(32); break;
    default: TM_TN_BN(16); break;
  }
#undef TM_TN_BN
#undef TM_TN_CASE
  TM_TRY(check_launch("gemm_tn"));
  {
    const int64_t count = M * N;
    unsigned blocks = (unsigned)cdiv(count, 64);
    split_reduce_kernel<<<blocks, 256, 0, st>>>(P, count, splits, C, N, ldc, accumulate);
    TM_TRY(check_launch("split_reduce"));
  }
  if (colsum) {
    split_reduce_kernel<<<(unsigned)cdiv(N, 64), 256, 0, st>>>(Pcs, N, splits, colsum, N, N, accumulate);
    TM_TRY(check_launch("split_reduce_cs"));
  }
  if (colsum_a) {
    split_reduce_kernel<<<(unsigned)cdiv(M, 64), 256, 0, st>>>(PcsA, M, splits, colsum_a, M, M, accumulate);
    TM_TRY(check_launch("split_reduce_csa"));
  }
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tmk

// Dense fp32 entry points: nn.Linear forward/backward building blocks (model.py:15,23 and the
// MLP call sites model.py:104,141-142,151,272,280,292), weight re-layout, split reduction.
#include "tm_gemm.cuh"

using namespace tmk;

namespace tmk {
// C[m][n] (+)= sum_k P[k][m*N + n].  Block = 64 elements x 4 split groups: group y adds splits y, y+4, ... (eight
// independent loads in flight), the four groups are folded in a fixed order -- deterministic, and four times fewer
// dependent additions per thread than one thread per element (the reductions sit between the weight-gradient GEMMs of
// the backward chain, where their latency is exposed).  Launch with split_reduce_grid(count) blocks of 256 threads.
__global__ void __launch_bounds__(256)
split_reduce_kernel(const float* __restrict__ P, int64_t count, int splits,
                    float* __restrict__ C, int64_t N, int64_t ldc, int accumulate) {
  __shared__ float sm[4][64];
  const int x = threadIdx.x & 63, y = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + x;
  float s = 0.f;
  if (i < count) {
    int k = y;
    for (; k + 28 < splits; k += 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = P[(int64_t)(k + 4 * u) * count + i];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < splits; k += 4) s += P[(int64_t)k * count + i];
  }
  sm[y][x] = s;
  __syncthreads();
  if (y == 0 && i < count) {
    s = ((sm[0][x] + sm[1][x]) + sm[2][x]) + sm[3][x];
    const int64_t m = i / N, n = i - m * N;
    float* p = C + m * ldc + n;
    *p = accumulate ? (*p + s) : s;
  }
}
}  // namespace tmk

namespace {
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                 int64_t cols) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
  }
}
// K <= 4 (fc_net_self layer 1: K = 2, model.py:50): pure streaming -- one thread per 4 outputs,
// 128-bit stores, the K x N weight stays in L1
__global__ void __launch_bounds__(256)
gemm_small_k_kernel(int64_t M, int64_t N, int K, const float* __restrict__ A, int64_t lda,
                    const int32_t* __restrict__ a_rows, const float* __restrict__ B, int64_t ldb,
                    float* __restrict__ C, int64_t ldc, const int32_t* __restrict__ c_rows,
                    const float* __restrict__ bias, int flags) {
  const int64_t n4s = N >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * n4s) return;
  const int64_t m = i / n4s, n = (i - m * n4s) << 2;
  const float* a = A + (a_rows ? (int64_t)a_rows[m] : m) * lda;
  float4 acc = (flags & TM_EPI_BIAS) ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < K; ++k) {
    const float av = a[k];
    const float4 b = __ldg(reinterpret_cast<const float4*>(B + (int64_t)k * ldb + n));
    acc.x = fmaf(av, b.x, acc.x); acc.y = fmaf(av, b.y, acc.y); acc.z = fmaf(av, b.z, acc.z); acc.w = fmaf(av, b.w, acc.w);
  }
  if (flags & TM_EPI_RELU) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
  st4(C + (c_rows ? (int64_t)c_rows[m] : m) * ldc + n, acc);
}
}  // namespace

extern "C" int tm_gemm_nn(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                          const int32_t* a_rows, const float* B, int64_t ldb, float* C, int64_t ldc,
                          const int32_t* c_rows, const float* bias, const float* mask,
                          int64_t ldmask, int flags, void* stream) {
  TM_REQUIRE(M >= 0 && N >= 0 && K >= 0, "tm_gemm_nn: negative size");
  TM_REQUIRE(!(flags & TM_EPI_BIAS) || bias, "tm_gemm_nn: TM_EPI_BIAS without bias");
  TM_REQUIRE(!(flags & TM_EPI_MASK) || mask, "tm_gemm_nn: TM_EPI_MASK without mask");
  if (K >= 1 && K <= 4 && M > 0 && N > 0 && (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && aligned16(B) && aligned16(C) &&
      (!bias || aligned16(bias)) && !(flags & (TM_EPI_MASK | TM_EPI_ACCUM))) {
    gemm_small_k_kernel<<<(unsigned)cdiv(M * (N / 4), 256), 256, 0, (cudaStream_t)stream>>>(M, N, (int)K, A, lda, a_rows, B, ldb, C,
                                                                                         ldc, c_rows, bias, flags);
    return check_launch("gemm_small_k");
  }
  PlainLoader al{A, lda, a_rows};
  PlainEpilogue ep{C, ldc, c_rows, bias, mask, ldmask, flags};
  const bool veca = (lda % 4 == 0) && (K % 4 == 0) && aligned16(A);
  return launch_gemm_nn(al, veca, B, ldb, ep, M, N, K, (cudaStream_t)stream);
}

extern "C" size_t tm_gemm_tn_ws(int64_t M, int64_t N, int64_t R) { return tn_ws_bytes(M, N, R); }

extern "C" int tm_gemm_tn(int64_t M, int64_t N, int64_t R, const float* A, int64_t lda,
                          const int32_t* a_rows, const float* B, int64_t ldb, const int32_t* b_rows,
                          float* C, int64_t ldc, float* colsum_a, float* colsum_b, int accumulate,
                          void* ws, size_t ws_bytes, void* stream) {
  TM_REQUIRE(M >= 0 && N >= 0 && R >= 0, "tm_gemm_tn: negative size");
  PlainLoader al{A, lda, a_rows};
  PlainLoader bl{B, ldb, b_rows};
  const bool veca = (lda % 4 == 0) && (M % 4 == 0) && aligned16(A);
  const bool vecb = (ldb % 4 == 0) && (N % 4 == 0) && aligned16(B);
  return launch_gemm_tn(al, veca, bl, vecb, M, N, R, C, ldc, colsum_a, colsum_b, accumulate, ws, ws_bytes,
                        (cudaStream_t)stream);
}

extern "C" int tm_transpose(int64_t rows, int64_t cols, const float* in, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, rows, cols);
  return check_launch("transpose");
}

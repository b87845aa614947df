// Tensor-core (tcgen05 / TMEM) entry points: dense layers, weight gradients and NHWC convolutions.
//
// Same contracts as the fp32 CUDA-core entry points in tm_gemm.cu / tm_conv.cu (which they replace on
// the hot path); `precision` selects 0 = bf16 operands (U-Net bf16 mode, rtol 2e-2) or 1 = split-bf16
// x3 (fp32-class accuracy, rtol 1e-3).  `err` is an optional device int set to 1 if a tensor-core
// barrier ever times out (never expected; it turns a would-be hang into a detectable failure).
#include "tm_tc.cuh"

using namespace tmk;

namespace {
// widest vector load the operand rows allow: 2 = 256-bit, 1 = 128-bit, 0 = scalar
inline int vec_mode(const void* p, int64_t ld) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((ld % 8 == 0) && (a % 32 == 0)) return 2;
  if ((ld % 4 == 0) && (a % 16 == 0)) return 1;
  return 0;
}
// splits of the reduction dimension so that (tiles x splits) fills the persistent grid once
inline void tn_split(int64_t M, int64_t N, int64_t R, int bn, int* splits, int64_t* k_per_split) {
  const int64_t tiles = cdiv(M, tc::BM) * cdiv(N, bn);
  int64_t want = (int64_t)sm_count() / tiles;
  const int64_t cap = cdiv(R, 4 * tc::BK);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  int64_t kps = cdiv(cdiv(R, want), tc::BK) * tc::BK;
  *splits = (int)cdiv(R, kps);
  *k_per_split = kps;
}
}  // namespace

extern "C" int tm_tc_gemm_nn(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                             const int32_t* a_rows, const float* B, int64_t ldb, int b_is_nk, float* C,
                             int64_t ldc, const int32_t* c_rows, const float* bias, const float* mask,
                             int64_t ldmask, int flags, int precision, int* err, void* stream) {
  TM_REQUIRE(M >= 0 && N >= 0 && K >= 0, "tm_tc_gemm_nn: negative size");
  TM_REQUIRE(!(flags & TM_EPI_BIAS) || bias, "tm_tc_gemm_nn: TM_EPI_BIAS without bias");
  TM_REQUIRE(!(flags & TM_EPI_MASK) || mask, "tm_tc_gemm_nn: TM_EPI_MASK without mask");
  cudaStream_t st = (cudaStream_t)stream;
  tc::RowLoader al{A, lda, a_rows, M, K, vec_mode(A, lda)};
  PlainEpilogue ep{C, ldc, c_rows, bias, mask, ldmask, flags};
  if (b_is_nk) {
    tc::RowLoader bl{B, ldb, nullptr, N, K, vec_mode(B, ldb)};
    return tc::launch(al, bl, ep, M, N, K, 1, K > 0 ? cdiv(K, tc::BK) * tc::BK : tc::BK, precision, err, st);
  }
  tc::ColLoader bl{B, ldb, nullptr, N, K, vec_mode(B, ldb)};
  return tc::launch(al, bl, ep, M, N, K, 1, K > 0 ? cdiv(K, tc::BK) * tc::BK : tc::BK, precision, err, st);
}

extern "C" size_t tm_tc_gemm_tn_ws(int64_t M, int64_t N, int64_t R) {
  size_t worst = 0;                       // the N tile depends on the precision chosen at launch
  for (int prec = 0; prec < 5; ++prec) {
    int splits;
    int64_t kps;
    tn_split(M, N, R, tc::pick_bn(N, prec), &splits, &kps);
    if ((size_t)splits > worst) worst = (size_t)splits;
  }
  return worst * M * N * sizeof(float) + 256;
}

extern "C" int tm_tc_gemm_tn(int64_t M, int64_t N, int64_t R, const float* A, int64_t lda,
                             const int32_t* a_rows, const float* B, int64_t ldb, const int32_t* b_rows,
                             float* C, int64_t ldc, int accumulate, int precision, void* ws, size_t ws_bytes,
                             int* err, void* stream) {
  TM_REQUIRE(M >= 0 && N >= 0 && R >= 0, "tm_tc_gemm_tn: negative size");
  if (M == 0 || N == 0) return 0;
  TM_REQUIRE(ws_bytes >= tm_tc_gemm_tn_ws(M, N, R), "tm_tc_gemm_tn: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int splits;
  int64_t kps;
  tn_split(M, N, R, tc::pick_bn(N, precision), &splits, &kps);
  tc::ColLoader al{A, lda, a_rows, M, R, vec_mode(A, lda)};
  tc::ColLoader bl{B, ldb, b_rows, N, R, vec_mode(B, ldb)};
  tc::PartialEpilogue ep{(float*)ws, M, N};
  TM_TRY(tc::launch(al, bl, ep, M, N, R, splits, kps, precision, err, st));
  split_reduce_kernel<<<(unsigned)cdiv(M * N, 256), 256, 0, st>>>((const float*)ws, M * N, splits, C, N, ldc, accumulate);
  return check_launch("split_reduce(tc)");
}

extern "C" int tm_tc_conv2d_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                                 const float* x, int64_t ldx, const float* wf, const float* bias, float* y,
                                 int64_t ldy, int flags, int precision, int* err, void* stream) {
  TM_REQUIRE(k >= 1 && (k & 1), "tm_tc_conv2d_nhwc: odd kernel sizes only");
  const int64_t M = B * H * W, K = k * k * Cin;
  tc::Im2colLoader8 al{x, ldx, (int)H, (int)W, (int)Cin, (int)k, (int)(k / 2), M, K,
                       (Cin % 4 == 0) ? vec_mode(x, ldx) : 0};
  tc::ColLoader bl{wf, Cout, nullptr, Cout, K, vec_mode(wf, Cout)};
  PlainEpilogue ep{y, ldy, nullptr, bias, nullptr, 0, (flags & TM_EPI_RELU) | (bias ? TM_EPI_BIAS : 0)};
  return tc::launch(al, bl, ep, M, Cout, K, 1, cdiv(K, tc::BK) * tc::BK, precision, err, (cudaStream_t)stream);
}

extern "C" size_t tm_tc_conv2d_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k) {
  return tm_tc_gemm_tn_ws(k * k * Cin, Cout, B * H * W);
}

extern "C" int tm_tc_conv2d_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                                       const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dwf,
                                       int precision, void* ws, size_t ws_bytes, int* err, void* stream) {
  const int64_t M = k * k * Cin, N = Cout, R = B * H * W;
  TM_REQUIRE(ws_bytes >= tm_tc_gemm_tn_ws(M, N, R), "tm_tc_conv2d_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int splits;
  int64_t kps;
  tn_split(M, N, R, tc::pick_bn(N, precision), &splits, &kps);
  tc::Im2colColLoader al{x, ldx, (int)H, (int)W, (int)Cin, (int)k, (int)(k / 2), M, R, (Cin % 4 == 0) ? vec_mode(x, ldx) : 0};
  tc::ColLoader bl{dy, lddy, nullptr, N, R, vec_mode(dy, lddy)};
  tc::PartialEpilogue ep{(float*)ws, M, N};
  TM_TRY(tc::launch(al, bl, ep, M, N, R, splits, kps, precision, err, st));
  split_reduce_kernel<<<(unsigned)cdiv(M * N, 256), 256, 0, st>>>((const float*)ws, M * N, splits, dwf, N, N, 0);
  return check_launch("split_reduce(tc conv)");
}

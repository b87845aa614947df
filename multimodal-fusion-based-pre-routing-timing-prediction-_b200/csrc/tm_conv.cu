// G4 (fp32 path): U-Net / LayoutNet image branch on NHWC activations.
//
// Replaces the cuDNN / ATen calls behind src/Unet.py:16-21 (Conv2d 3x3 + BatchNorm2d + ReLU),
// :53 (ConvTranspose2d k2 s2), :75-77 (1x1 conv, pool, ReLU), :89-91 (pooling) and
// src/model.py:227-243 (LayoutNet 9x9 / 7x7 convs).  Convolutions are implicit GEMMs on the shared
// fp32 GEMM core (tm_gemm.cuh) with im2col performed by the operand loader; every activation
// carries an explicit pixel stride so torch.cat (Unet.py:67) is free: producers write straight
// into the halves of the concat buffer.  The bf16 tcgen05 path lives in tm_conv_tc.cu.
#include "tm_gemm.cuh"

using namespace tmk;

namespace {
// ------------------------------------------------------------------------------ layout kernels
__global__ void nchw_to_nhwc_kernel(int64_t B, int64_t C, int64_t HW, const float* __restrict__ in,
                                    float* __restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C * HW) return;
  const int64_t c = i % C, p = (i / C) % HW, b = i / (C * HW);
  out[(b * HW + p) * ld + c] = in[(b * C + c) * HW + p];
}
__global__ void nhwc_to_nchw_kernel(int64_t B, int64_t C, int64_t HW, const float* __restrict__ in,
                                    int64_t ld, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C * HW) return;
  const int64_t p = i % HW, c = (i / HW) % C, b = i / (C * HW);
  out[i] = in[(b * HW + p) * ld + c];
}
__global__ void conv_pack_kernel(int64_t Cout, int64_t Cin, int64_t k, const float* __restrict__ w,
                                 float* __restrict__ wf, float* __restrict__ wb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * k * k) return;
  const int64_t tx = i % k, ty = (i / k) % k, ci = (i / (k * k)) % Cin, co = i / (k * k * Cin);
  const float v = w[i];
  if (wf) wf[((ty * k + tx) * Cin + ci) * Cout + co] = v;
  if (wb) wb[(((k - 1 - ty) * k + (k - 1 - tx)) * Cout + co) * Cin + ci] = v;
}
__global__ void conv_unpack_kernel(int64_t Cout, int64_t Cin, int64_t k, const float* __restrict__ dwf,
                                   float* __restrict__ dw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * k * k) return;
  const int64_t tx = i % k, ty = (i / k) % k, ci = (i / (k * k)) % Cin, co = i / (k * k * Cin);
  dw[i] = dwf[((ty * k + tx) * Cin + ci) * Cout + co];
}
__global__ void convt_pack_kernel(int64_t Cin, int64_t Cout, const float* __restrict__ w,
                                  float* __restrict__ wt, float* __restrict__ wtT) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin * Cout * 4) return;
  const int64_t q = i % 4, co = (i / 4) % Cout, ci = i / (4 * Cout);  // w[ci][co][dy][dx], q = dy*2+dx
  const float v = w[i];
  if (wt) wt[ci * 4 * Cout + q * Cout + co] = v;
  if (wtT) wtT[(q * Cout + co) * Cin + ci] = v;
}
__global__ void convt_unpack_kernel(int64_t Cin, int64_t Cout, const float* __restrict__ dwt,
                                    float* __restrict__ dw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin * Cout * 4) return;
  const int64_t q = i % 4, co = (i / 4) % Cout, ci = i / (4 * Cout);
  dw[i] = dwt[ci * 4 * Cout + q * Cout + co];
}
__global__ void fold4_kernel(int64_t Cout, const float* __restrict__ cs, float* __restrict__ dbias) {
  const int64_t co = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (co < Cout) dbias[co] = cs[co] + cs[Cout + co] + cs[2 * Cout + co] + cs[3 * Cout + co];
}

// ------------------------------------------------------------------------------ batch norm
// block = 32 channels x 8 pixel lanes; grid.x = pixel chunks, grid.y = channel groups of 32
constexpr int BN_PIX_PER_BLOCK = 128;

__global__ void __launch_bounds__(256)
bn_stats_kernel(int64_t npix, int64_t C, const float* __restrict__ x, int64_t ldx, double* __restrict__ part) {
  __shared__ double s1[8][32], s2[8][32];
  const int c = blockIdx.y * 32 + threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * BN_PIX_PER_BLOCK;
  const int64_t p1 = (p0 + BN_PIX_PER_BLOCK < npix) ? p0 + BN_PIX_PER_BLOCK : npix;
  double a = 0.0, b = 0.0;
  if (c < C)
    for (int64_t p = p0 + threadIdx.y; p < p1; p += 8) {
      const double v = (double)x[p * ldx + c];
      a += v;
      b += v * v;
    }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    part[((int64_t)blockIdx.x * C + c) * 2 + 0] = a;
    part[((int64_t)blockIdx.x * C + c) * 2 + 1] = b;
  }
}

__global__ void bn_finalize_kernel(int64_t npix, int64_t C, int nblk, const double* __restrict__ part,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per channel
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < nblk; i += 32) { a += part[((int64_t)i * C + c) * 2]; b += part[((int64_t)i * C + c) * 2 + 1]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane != 0) return;
  const double mean = a / (double)npix;
  double var = b / (double)npix - mean * mean;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_relu_apply_kernel(int64_t npix, int64_t C, const float* __restrict__ x, int64_t ldx,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     float* __restrict__ y, int64_t ldy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const int64_t p = i / C, c = i - p * C;
  const float v = (x[p * ldx + c] - mean[c]) * invstd[c] * gamma[c] + beta[c];
  y[p * ldy + c] = fmaxf(v, 0.f);
}

__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(int64_t npix, int64_t C, const float* __restrict__ x, int64_t ldx,
                    const float* __restrict__ y, int64_t ldy, const float* __restrict__ dy, int64_t lddy,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    double* __restrict__ part) {
  __shared__ double s1[8][32], s2[8][32];
  const int c = blockIdx.y * 32 + threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * BN_PIX_PER_BLOCK;
  const int64_t p1 = (p0 + BN_PIX_PER_BLOCK < npix) ? p0 + BN_PIX_PER_BLOCK : npix;
  double a = 0.0, b = 0.0;
  if (c < C) {
    const float mu = mean[c], is = invstd[c];
    for (int64_t p = p0 + threadIdx.y; p < p1; p += 8) {
      const float g = (y[p * ldy + c] > 0.f) ? dy[p * lddy + c] : 0.f;
      a += (double)g;
      b += (double)g * (double)((x[p * ldx + c] - mu) * is);
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    part[((int64_t)blockIdx.x * C + c) * 2 + 0] = a;
    part[((int64_t)blockIdx.x * C + c) * 2 + 1] = b;
  }
}

__global__ void bn_bwd_finalize_kernel(int64_t C, int nblk, const double* __restrict__ part,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per channel
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < nblk; i += 32) { a += part[((int64_t)i * C + c) * 2]; b += part[((int64_t)i * C + c) * 2 + 1]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { dbeta[c] = (float)a; dgamma[c] = (float)b; }
}

__global__ void bn_bwd_apply_kernel(int64_t npix, int64_t C, const float* __restrict__ x, int64_t ldx,
                                    const float* __restrict__ y, int64_t ldy, const float* __restrict__ dy,
                                    int64_t lddy, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                    float* __restrict__ dx, int64_t lddx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const int64_t p = i / C, c = i - p * C;
  const float g = (y[p * ldy + c] > 0.f) ? dy[p * lddy + c] : 0.f;
  const float xh = (x[p * ldx + c] - mean[c]) * invstd[c];
  const float inv_n = 1.f / (float)npix;
  dx[p * lddx + c] = gamma[c] * invstd[c] * (g - dbeta[c] * inv_n - xh * dgamma[c] * inv_n);
}

// ------------------------------------------------------------------------------ pooling etc.
__global__ void pool_fwd_kernel(int64_t B, int64_t H, int64_t W, int64_t C, int mode,
                                const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
                                uint8_t* __restrict__ idx, int relu) {
  const int64_t Ho = H / 2, Wo = W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ho * Wo * C) return;
  const int64_t c = i % C, xo = (i / C) % Wo, yo = (i / (C * Wo)) % Ho, b = i / (C * Wo * Ho);
  const float* base = x + ((b * H + 2 * yo) * W + 2 * xo) * ldx + c;
  const float v0 = base[0], v1 = base[ldx], v2 = base[W * ldx], v3 = base[W * ldx + ldx];
  float out;
  if (mode == 0) {
    int am = 0;
    out = v0;
    if (v1 > out || v1 != v1) { out = v1; am = 1; }   // first maximum wins, NaN propagates (ATen rule)
    if (v2 > out || v2 != v2) { out = v2; am = 2; }
    if (v3 > out || v3 != v3) { out = v3; am = 3; }
    if (idx) idx[i] = (uint8_t)am;
  } else {
    out = 0.25f * (v0 + v1 + v2 + v3);
  }
  if (relu) out = fmaxf(out, 0.f);
  y[((b * Ho + yo) * Wo + xo) * ldy + c] = out;
}

__global__ void pool_bwd_kernel(int64_t B, int64_t H, int64_t W, int64_t C, int mode,
                                const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y,
                                int64_t ldy, const uint8_t* __restrict__ idx, float* __restrict__ dx,
                                int64_t lddx, int relu) {
  const int64_t Ho = H / 2, Wo = W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H * W * C) return;
  const int64_t c = i % C, xi = (i / C) % W, yi = (i / (C * W)) % H, b = i / (C * W * H);
  const int64_t yo = yi / 2, xo = xi / 2;
  float g = 0.f;
  if (yo < Ho && xo < Wo) {
    const int64_t o = ((b * Ho + yo) * Wo + xo);
    g = dy[o * lddy + c];
    if (relu && !(y[o * ldy + c] > 0.f)) g = 0.f;
    if (mode == 0) {
      const int am = idx[o * C + c];
      if (am != (int)((yi & 1) * 2 + (xi & 1))) g = 0.f;
    } else {
      g *= 0.25f;
    }
  }
  dx[((b * H + yi) * W + xi) * lddx + c] = g;
}

__global__ void lrelu_fwd_kernel(int64_t n, const float* __restrict__ x, float slope, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = v > 0.f ? v : v * slope; }
}
__global__ void lrelu_bwd_kernel(int64_t n, const float* __restrict__ y, const float* __restrict__ dy,
                                 float slope, float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.f ? dy[i] : dy[i] * slope;
}
__global__ void add_strided_kernel(int64_t npix, int64_t C, const float* __restrict__ src, int64_t lds,
                                   float* __restrict__ dst, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const int64_t p = i / C, c = i - p * C;
  dst[p * ldd + c] += src[p * lds + c];
}

inline unsigned blocks_for(int64_t n) { return (unsigned)cdiv(n, 256); }
}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int tm_nchw_to_nhwc(int64_t B, int64_t C, int64_t H, int64_t W, const float* in, float* out,
                               int64_t ld, void* stream) {
  if (B * C * H * W <= 0) return 0;
  nchw_to_nhwc_kernel<<<blocks_for(B * C * H * W), 256, 0, ST>>>(B, C, H * W, in, out, ld);
  return check_launch("nchw_to_nhwc");
}
extern "C" int tm_nhwc_to_nchw(int64_t B, int64_t C, int64_t H, int64_t W, const float* in, int64_t ld,
                               float* out, void* stream) {
  if (B * C * H * W <= 0) return 0;
  nhwc_to_nchw_kernel<<<blocks_for(B * C * H * W), 256, 0, ST>>>(B, C, H * W, in, ld, out);
  return check_launch("nhwc_to_nchw");
}
extern "C" int tm_conv_pack_weight(int64_t Cout, int64_t Cin, int64_t k, const float* w, float* wf,
                                   float* wb, void* stream) {
  conv_pack_kernel<<<blocks_for(Cout * Cin * k * k), 256, 0, ST>>>(Cout, Cin, k, w, wf, wb);
  return check_launch("conv_pack");
}
extern "C" int tm_conv_unpack_wgrad(int64_t Cout, int64_t Cin, int64_t k, const float* dwf, float* dw,
                                    void* stream) {
  conv_unpack_kernel<<<blocks_for(Cout * Cin * k * k), 256, 0, ST>>>(Cout, Cin, k, dwf, dw);
  return check_launch("conv_unpack");
}

extern "C" int tm_conv2d_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                              const float* x, int64_t ldx, const float* wf, const float* bias, float* y,
                              int64_t ldy, int flags, void* stream) {
  TM_REQUIRE(k >= 1 && (k & 1), "tm_conv2d_nhwc: odd kernel sizes only");
  Im2colLoader al{x, ldx, (int)H, (int)W, (int)Cin, (int)k, (int)(k / 2)};
  PlainEpilogue ep{y, ldy, nullptr, bias, nullptr, 0, (flags & TM_EPI_RELU) | (bias ? TM_EPI_BIAS : 0)};
  const bool veca = (Cin % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  return launch_gemm_nn(al, veca, wf, Cout, ep, B * H * W, Cout, k * k * Cin, ST);
}

extern "C" size_t tm_conv2d_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k) {
  return tn_ws_bytes(k * k * Cin, Cout, B * H * W);
}
extern "C" int tm_conv2d_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                                    const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dwf,
                                    float* dbias, void* ws, size_t ws_bytes, void* stream) {
  Im2colLoader al{x, ldx, (int)H, (int)W, (int)Cin, (int)k, (int)(k / 2)};
  PlainLoader bl{dy, lddy, nullptr};
  const bool veca = (Cin % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  const bool vecb = (Cout % 4 == 0) && (lddy % 4 == 0) && aligned16(dy);
  return launch_gemm_tn(al, veca, bl, vecb, k * k * Cin, Cout, B * H * W, dwf, Cout, nullptr, dbias, 0, ws, ws_bytes, ST);
}

extern "C" int tm_convt_pack_weight(int64_t Cin, int64_t Cout, const float* w, float* wt, float* wtT,
                                    void* stream) {
  convt_pack_kernel<<<blocks_for(Cin * Cout * 4), 256, 0, ST>>>(Cin, Cout, w, wt, wtT);
  return check_launch("convt_pack");
}
extern "C" int tm_convt_unpack_wgrad(int64_t Cin, int64_t Cout, const float* dwt, float* dw, void* stream) {
  convt_unpack_kernel<<<blocks_for(Cin * Cout * 4), 256, 0, ST>>>(Cin, Cout, dwt, dw);
  return check_launch("convt_unpack");
}
extern "C" int tm_convt2x2_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, const float* x,
                                int64_t ldx, const float* wt, const float* bias, float* y, int64_t ldy,
                                int64_t Hy, int64_t Wy, int64_t oy, int64_t ox, void* stream) {
  TM_REQUIRE(2 * H + oy <= Hy && 2 * W + ox <= Wy && oy >= 0 && ox >= 0, "tm_convt2x2: window outside y");
  PlainLoader al{x, ldx, nullptr};
  ConvtEpilogue ep{y, ldy, bias, (int)H, (int)W, (int)Cout, (int)Hy, (int)Wy, (int)oy, (int)ox};
  const bool veca = (Cin % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  return launch_gemm_nn(al, veca, wt, 4 * Cout, ep, B * H * W, 4 * Cout, Cin, ST);
}
extern "C" int tm_convt2x2_dgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                                      const float* dyo, int64_t lddy, int64_t Hy, int64_t Wy, int64_t oy,
                                      int64_t ox, const float* wtT, float* dx, int64_t lddx, void* stream) {
  ConvtLoader al{dyo, lddy, (int)H, (int)W, (int)Cout, (int)Hy, (int)Wy, (int)oy, (int)ox};
  PlainEpilogue ep{dx, lddx, nullptr, nullptr, nullptr, 0, 0};
  const bool veca = (Cout % 4 == 0) && (lddy % 4 == 0) && aligned16(dyo);
  return launch_gemm_nn(al, veca, wtT, Cin, ep, B * H * W, Cin, 4 * Cout, ST);
}
extern "C" size_t tm_convt2x2_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout) {
  return tn_ws_bytes(Cin, 4 * Cout, B * H * W) + (size_t)4 * Cout * sizeof(float) + 256;
}
extern "C" int tm_convt2x2_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                                      const float* x, int64_t ldx, const float* dyo, int64_t lddy,
                                      int64_t Hy, int64_t Wy, int64_t oy, int64_t ox, float* dwt,
                                      float* dbias, void* ws, size_t ws_bytes, void* stream) {
  TM_REQUIRE(ws_bytes >= tm_convt2x2_wgrad_ws(B, H, W, Cin, Cout), "tm_convt2x2_wgrad: workspace too small");
  PlainLoader al{x, ldx, nullptr};
  ConvtLoader bl{dyo, lddy, (int)H, (int)W, (int)Cout, (int)Hy, (int)Wy, (int)oy, (int)ox};
  const bool veca = (Cin % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  const bool vecb = (Cout % 4 == 0) && (lddy % 4 == 0) && aligned16(dyo);
  float* cs = (float*)ws;                                   // [4*Cout] column sums
  char* rest = (char*)ws + align_up((size_t)4 * Cout * sizeof(float), 256);
  TM_TRY(launch_gemm_tn(al, veca, bl, vecb, Cin, 4 * Cout, B * H * W, dwt, 4 * Cout, nullptr, dbias ? cs : nullptr, 0,
                        rest, ws_bytes - (size_t)(rest - (char*)ws), ST));
  if (dbias) {
    fold4_kernel<<<blocks_for(Cout), 256, 0, ST>>>(Cout, cs, dbias);
    TM_TRY(check_launch("fold4"));
  }
  return 0;
}

extern "C" size_t tm_bn_ws(int64_t npix, int64_t C) {
  return (size_t)cdiv(npix, BN_PIX_PER_BLOCK) * C * 2 * sizeof(double) + 256;
}

extern "C" int tm_bn_relu_forward(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var,
                                  float momentum, float eps, float* y, int64_t ldy, float* save_mean,
                                  float* save_invstd, void* ws, size_t ws_bytes, void* stream) {
  TM_REQUIRE(npix > 0 && C > 0, "tm_bn_relu_forward: bad sizes");
  TM_REQUIRE(ws_bytes >= tm_bn_ws(npix, C), "tm_bn_relu_forward: workspace too small");
  const int nblk = (int)cdiv(npix, BN_PIX_PER_BLOCK);
  double* part = (double*)ws;
  bn_stats_kernel<<<dim3(nblk, (unsigned)cdiv(C, 32)), dim3(32, 8), 0, ST>>>(npix, C, x, ldx, part);
  TM_TRY(check_launch("bn_stats"));
  bn_finalize_kernel<<<(unsigned)cdiv(C * 32, 128), 128, 0, ST>>>(npix, C, nblk, part, running_mean, running_var,
                                                             momentum, eps, save_mean, save_invstd);
  TM_TRY(check_launch("bn_finalize"));
  bn_relu_apply_kernel<<<blocks_for(npix * C), 256, 0, ST>>>(npix, C, x, ldx, gamma, beta, save_mean,
                                                             save_invstd, y, ldy);
  return check_launch("bn_relu_apply");
}

extern "C" int tm_bn_relu_backward(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* y,
                                   int64_t ldy, const float* dy, int64_t lddy, const float* gamma,
                                   const float* save_mean, const float* save_invstd, float* dx,
                                   int64_t lddx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                                   void* stream) {
  TM_REQUIRE(npix > 0 && C > 0, "tm_bn_relu_backward: bad sizes");
  TM_REQUIRE(ws_bytes >= tm_bn_ws(npix, C), "tm_bn_relu_backward: workspace too small");
  const int nblk = (int)cdiv(npix, BN_PIX_PER_BLOCK);
  double* part = (double*)ws;
  bn_bwd_stats_kernel<<<dim3(nblk, (unsigned)cdiv(C, 32)), dim3(32, 8), 0, ST>>>(npix, C, x, ldx, y, ldy, dy,
                                                                                 lddy, save_mean, save_invstd,
                                                                                 part);
  TM_TRY(check_launch("bn_bwd_stats"));
  bn_bwd_finalize_kernel<<<(unsigned)cdiv(C * 32, 128), 128, 0, ST>>>(C, nblk, part, dgamma, dbeta);
  TM_TRY(check_launch("bn_bwd_finalize"));
  bn_bwd_apply_kernel<<<blocks_for(npix * C), 256, 0, ST>>>(npix, C, x, ldx, y, ldy, dy, lddy, gamma, save_mean,
                                                            save_invstd, dgamma, dbeta, dx, lddx);
  return check_launch("bn_bwd_apply");
}

extern "C" int tm_pool2x2_forward(int64_t B, int64_t H, int64_t W, int64_t C, int mode, const float* x,
                                  int64_t ldx, float* y, int64_t ldy, uint8_t* idx, int flags, void* stream) {
  TM_REQUIRE(mode == 0 || mode == 1, "tm_pool2x2: mode must be 0 (max) or 1 (avg)");
  const int64_t n = B * (H / 2) * (W / 2) * C;
  if (n <= 0) return 0;
  pool_fwd_kernel<<<blocks_for(n), 256, 0, ST>>>(B, H, W, C, mode, x, ldx, y, ldy, idx, (flags & TM_EPI_RELU) != 0);
  return check_launch("pool_fwd");
}
extern "C" int tm_pool2x2_backward(int64_t B, int64_t H, int64_t W, int64_t C, int mode, const float* dy,
                                   int64_t lddy, const float* y, int64_t ldy, const uint8_t* idx, float* dx,
                                   int64_t lddx, int flags, void* stream) {
  TM_REQUIRE(mode == 0 || mode == 1, "tm_pool2x2: mode must be 0 (max) or 1 (avg)");
  TM_REQUIRE(mode == 1 || idx, "tm_pool2x2_backward: max pooling needs idx");
  const int64_t n = B * H * W * C;
  if (n <= 0) return 0;
  pool_bwd_kernel<<<blocks_for(n), 256, 0, ST>>>(B, H, W, C, mode, dy, lddy, y, ldy, idx, dx, lddx,
                                                 (flags & TM_EPI_RELU) != 0);
  return check_launch("pool_bwd");
}
extern "C" int tm_leaky_relu_forward(int64_t n, const float* x, float slope, float* y, void* stream) {
  if (n <= 0) return 0;
  lrelu_fwd_kernel<<<blocks_for(n), 256, 0, ST>>>(n, x, slope, y);
  return check_launch("lrelu_fwd");
}
extern "C" int tm_leaky_relu_backward(int64_t n, const float* y, const float* dy, float slope, float* dx,
                                      void* stream) {
  if (n <= 0) return 0;
  lrelu_bwd_kernel<<<blocks_for(n), 256, 0, ST>>>(n, y, dy, slope, dx);
  return check_launch("lrelu_bwd");
}
extern "C" int tm_add_strided(int64_t npix, int64_t C, const float* src, int64_t lds, float* dst, int64_t ldd,
                              void* stream) {
  if (npix * C <= 0) return 0;
  add_strided_kernel<<<blocks_for(npix * C), 256, 0, ST>>>(npix, C, src, lds, dst, ldd);
  return check_launch("add_strided");
}

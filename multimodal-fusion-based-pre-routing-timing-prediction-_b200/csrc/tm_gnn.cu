// G2/G3: level-wise timing propagation, forward and backward.
//
// Replaces PathConv.forward and its UDFs (src/model.py:88-116,138-153,158-213), which the reference
// runs as one DGL `pull` + several cuBLAS calls + a full (N,128) index_copy PER LEVEL, and the
// autograd backward through them (src/train.py:553).
//
// Data layout (HBM): H, S, G are [N][128] fp32 row-major indexed by pin id (one pin = one 512 B
// row = four full 128 B lines, so scattered rows still move at full sector efficiency); the
// buffers saved for backward (A, LSE, HID, GA, GHID, GZC) are compact [n_cell_rows][...] in
// schedule order, so each cell level owns a contiguous slab.
//
// Level schedule: pins sorted by (level, id); level l occupies order[level_ptr[l]..level_ptr[l+1]).
//   level 0        h[v] = relu(S[v])                                   (model.py:148-153,200-208)
//   odd  level     h[v] = relu(S[v] + mean_{u->v in net} h[u])         (model.py:103-108,186-187)
//   even level >0  a = sum_e m_e * softmax_e(m)_e (per channel), m_e = h[src e]   (model.py:113-116)
//                  h[v] = relu(S[v] + W2 relu(W1 a + b1) + b2)         (model.py:138-146)
// A warp owns a pin (lane = 4 channels, 128-bit loads); cell levels are processed in 16-pin tiles:
// the aggregated rows are staged in shared memory and pushed through the 128->256->128 MLP by the
// same CTA, so `a` and the hidden layer never round-trip through HBM on the forward critical path.
#include "tm_common.cuh"

using namespace tmk;

namespace {
constexpr int D = 128;     // out_feat_dim (model.py:43, options.py:10)
constexpr int HID = 256;   // MLP hidden width (model.py:48)
constexpr int TILE = 16;   // pins per CTA on a cell level
constexpr int CT = 256;    // threads per CTA

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4relu(float4 a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }

// ---------------------------------------------------------------------------------------------
// forward, level 0 and odd (net) levels: one warp per pin
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gnn_net_fwd_kernel(const int* __restrict__ order, int p0, int cnt, const int* __restrict__ iptr,
                   const int* __restrict__ isrc, const float* __restrict__ S, float* H, int level0) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= cnt) return;
  const int v = order[p0 + w];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!level0) {
    const int s = iptr[v], e = iptr[v + 1];
    for (int i = s; i < e; i += 4) {
      float4 m[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        m[q] = (i + q < e) ? ld4(H + (int64_t)isrc[i + q] * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc = f4add(acc, m[q]);
    }
    if (e > s) acc = f4scale(acc, 1.f / (float)(e - s));
  }
  const float4 sv = ld4_stream(S + (int64_t)v * D + lane * 4);
  st4(H + (int64_t)v * D + lane * 4, f4relu(f4add(sv, acc)));
}

// ---------------------------------------------------------------------------------------------
// tile MLP pieces shared by forward and backward (fp32 FFMA, weights streamed through L1/L2)
//   gemm_128x256: out[16][256] = in[16][128] @ Wk[128][256]     thread -> 4 rows x 4 cols
//   gemm_256x128: out[16][128] = in[16][256] @ Wk[256][128]     thread -> 2 rows x 4 cols
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_128x256(const float (*in)[D], const float* __restrict__ Wk,
                                             int rg, int cg, float (&acc)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
  for (int k = 0; k < D; k += 4) {
    float4 a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&in[rg * 4 + r][k]);
    float4 wv[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wv[kk] = __ldg(reinterpret_cast<const float4*>(Wk + (int64_t)(k + kk) * HID + cg * 4));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float av[4] = {a[r].x, a[r].y, a[r].z, a[r].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[r][0] = fmaf(av[kk], wv[kk].x, acc[r][0]);
        acc[r][1] = fmaf(av[kk], wv[kk].y, acc[r][1]);
        acc[r][2] = fmaf(av[kk], wv[kk].z, acc[r][2]);
        acc[r][3] = fmaf(av[kk], wv[kk].w, acc[r][3]);
      }
    }
  }
}

__device__ __forceinline__ void gemm_256x128(const float (*in)[HID], const float* __restrict__ Wk,
                                             int rg, int cg, float (&acc)[2][4]) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
  for (int k = 0; k < HID; k += 4) {
    float4 a[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) a[r] = *reinterpret_cast<const float4*>(&in[rg * 2 + r][k]);
    float4 wv[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wv[kk] = __ldg(reinterpret_cast<const float4*>(Wk + (int64_t)(k + kk) * D + cg * 4));
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float av[4] = {a[r].x, a[r].y, a[r].z, a[r].w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[r][0] = fmaf(av[kk], wv[kk].x, acc[r][0]);
        acc[r][1] = fmaf(av[kk], wv[kk].y, acc[r][1]);
        acc[r][2] = fmaf(av[kk], wv[kk].z, acc[r][2]);
        acc[r][3] = fmaf(av[kk], wv[kk].w, acc[r][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward, even level > 0: gather + per-channel softmax-weighted sum + MLP, 16 pins per CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT)
gnn_cell_fwd_kernel(const int* __restrict__ order, int p0, int cnt, int crow0,
                    const int* __restrict__ iptr, const int* __restrict__ isrc,
                    const float* __restrict__ S, float* H, const float* __restrict__ W1t,
                    const float* __restrict__ b1, const float* __restrict__ W2t,
                    const float* __restrict__ b2, float* __restrict__ A, float* __restrict__ LSE,
                    float* __restrict__ HIDb) {
  __shared__ __align__(16) float a_s[TILE][D];
  __shared__ __align__(16) float hid_s[TILE][HID];
  __shared__ int v_s[TILE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * TILE;

  // phase 1: each warp aggregates two pins (online softmax per channel, single pass over edges)
#pragma unroll
  for (int rr = 0; rr < TILE / (CT / 32); ++rr) {
    const int r = warp * (TILE / (CT / 32)) + rr;
    const int p = t0 + r;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), lse = av;
    int v = -1;
    if (p < cnt) {
      v = order[p0 + p];
      const int s = iptr[v], e = iptr[v + 1];
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      float sm[4] = {0.f, 0.f, 0.f, 0.f}, tw[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = s; i < e; i += 4) {
        float4 m[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i + q < e) m[q] = ld4(H + (int64_t)isrc[i + q] * D + lane * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (i + q < e) {
            const float mv[4] = {m[q].x, m[q].y, m[q].z, m[q].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float nm = fmaxf(mx[c], mv[c]);
              const float sc = expf(mx[c] - nm);      // exp(-inf) = 0 on the first edge
              const float ex = expf(mv[c] - nm);
              sm[c] = sm[c] * sc + ex;
              tw[c] = tw[c] * sc + mv[c] * ex;
              mx[c] = nm;
            }
          }
        }
      }
      if (e > s) {
        av = make_float4(tw[0] / sm[0], tw[1] / sm[1], tw[2] / sm[2], tw[3] / sm[3]);
        lse = make_float4(mx[0] + logf(sm[0]), mx[1] + logf(sm[1]), mx[2] + logf(sm[2]), mx[3] + logf(sm[3]));
      }
      if (A) {
        st4(A + (int64_t)(crow0 + p) * D + lane * 4, av);
        st4(LSE + (int64_t)(crow0 + p) * D + lane * 4, lse);
      }
    }
    *reinterpret_cast<float4*>(&a_s[r][lane * 4]) = av;
    if (lane == 0) v_s[r] = v;
  }
  __syncthreads();

  // phase 2: hidden = relu(a @ W1t + b1)
  {
    const int cg = tid & 63, rg = tid >> 6;
    float acc[4][4];
    gemm_128x256(a_s, W1t, rg, cg, acc);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + cg * 4));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 o = f4relu(make_float4(acc[r][0] + bb.x, acc[r][1] + bb.y, acc[r][2] + bb.z, acc[r][3] + bb.w));
      *reinterpret_cast<float4*>(&hid_s[rg * 4 + r][cg * 4]) = o;
      const int p = t0 + rg * 4 + r;
      if (HIDb && p < cnt) st4(HIDb + (int64_t)(crow0 + p) * HID + cg * 4, o);
    }
  }
  __syncthreads();

  // phase 3: h = relu(S + hidden @ W2t + b2)
  {
    const int cg = tid & 31, rg = tid >> 5;
    float acc[2][4];
    gemm_256x128(hid_s, W2t, rg, cg, acc);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + cg * 4));
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int v = v_s[rg * 2 + r];
      if (v >= 0) {
        const float4 sv = ld4_stream(S + (int64_t)v * D + cg * 4);
        float4 o = make_float4(acc[r][0] + bb.x + sv.x, acc[r][1] + bb.y + sv.y, acc[r][2] + bb.z + sv.z,
                               acc[r][3] + bb.w + sv.w);
        st4(H + (int64_t)v * D + cg * 4, f4relu(o));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: every pin PULLS its gradient from its successors over the out-edge CSRs (no atomics,
// deterministic).  Appendix B of SURVEY.md:
//   net  edge v->u (u odd level):   g_h[v] += g_z[u] / indeg_net(u)
//   cell edge v->u (u even level):  g_h[v] += g_a[u] * w_e * (1 + h[v] - a[u]),  w_e = exp(h[v]-lse[u])
// then g_z[v] = g_h[v] * (h[v] > 0) is written back into G[v].
// ---------------------------------------------------------------------------------------------
struct SchedDev {
  const int* order;
  const int* level;
  const int* crow;
  const int* net_iptr;
  const int* net_optr;
  const int* net_odst;
  const int* cell_optr;
  const int* cell_odst;
};

__device__ __forceinline__ float4 pull_gz(const SchedDev& s, int v, int lane, const float* __restrict__ H,
                                          const float* G, const float* __restrict__ GA,
                                          const float* __restrict__ A, const float* __restrict__ LSE) {
  const int64_t off = (int64_t)v * D + lane * 4;
  float4 g = ld4(G + off);
  const float4 hv = ld4(H + off);
  const int lv = s.level[v];
  for (int e = s.net_optr[v], e1 = s.net_optr[v + 1]; e < e1; ++e) {
    const int u = s.net_odst[e];
    const int lu = s.level[u];
    if ((lu & 1) && lu > lv) {
      const float inv = 1.f / (float)(s.net_iptr[u + 1] - s.net_iptr[u]);
      g = f4add(g, f4scale(ld4(G + (int64_t)u * D + lane * 4), inv));
    }
  }
  for (int e = s.cell_optr[v], e1 = s.cell_optr[v + 1]; e < e1; ++e) {
    const int u = s.cell_odst[e];
    const int lu = s.level[u];
    if (lu > 0 && !(lu & 1) && lu > lv) {
      const int64_t co = (int64_t)s.crow[u] * D + lane * 4;
      const float4 ga = ld4(GA + co), ls = ld4(LSE + co), aa = ld4(A + co);
      g.x += ga.x * expf(hv.x - ls.x) * (1.f + hv.x - aa.x);
      g.y += ga.y * expf(hv.y - ls.y) * (1.f + hv.y - aa.y);
      g.z += ga.z * expf(hv.z - ls.z) * (1.f + hv.z - aa.z);
      g.w += ga.w * expf(hv.w - ls.w) * (1.f + hv.w - aa.w);
    }
  }
  return make_float4(hv.x > 0.f ? g.x : 0.f, hv.y > 0.f ? g.y : 0.f, hv.z > 0.f ? g.z : 0.f,
                     hv.w > 0.f ? g.w : 0.f);
}

__global__ void __launch_bounds__(256)
gnn_net_bwd_kernel(SchedDev s, int p0, int cnt, const float* __restrict__ H, float* G,
                   const float* __restrict__ GA, const float* __restrict__ A,
                   const float* __restrict__ LSE) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= cnt) return;
  const int v = s.order[p0 + w];
  const float4 gz = pull_gz(s, v, lane, H, G, GA, A, LSE);
  st4(G + (int64_t)v * D + lane * 4, gz);
}

__global__ void __launch_bounds__(CT)
gnn_cell_bwd_kernel(SchedDev s, int p0, int cnt, int crow0, const float* __restrict__ H, float* G,
                    const float* __restrict__ W1, const float* __restrict__ W2, float* GA,
                    const float* __restrict__ A, const float* __restrict__ LSE,
                    const float* __restrict__ HIDb, float* __restrict__ GHID, float* __restrict__ GZC) {
  __shared__ __align__(16) float gz_s[TILE][D];
  __shared__ __align__(16) float gh_s[TILE][HID];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * TILE;
#pragma unroll
  for (int rr = 0; rr < TILE / (CT / 32); ++rr) {
    const int r = warp * (TILE / (CT / 32)) + rr;
    const int p = t0 + r;
    float4 gz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < cnt) {
      const int v = s.order[p0 + p];
      gz = pull_gz(s, v, lane, H, G, GA, A, LSE);
      st4(G + (int64_t)v * D + lane * 4, gz);
      st4(GZC + (int64_t)(crow0 + p) * D + lane * 4, gz);
    }
    *reinterpret_cast<float4*>(&gz_s[r][lane * 4]) = gz;
  }
  __syncthreads();
  {  // g_hid = (g_z @ W2) * (hid > 0);  W2 is [128][256] as stored by nn.Linear(256,128)
    const int cg = tid & 63, rg = tid >> 6;
    float acc[4][4];
    gemm_128x256(gz_s, W2, rg, cg, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int p = t0 + rg * 4 + r;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < cnt) {
        const float4 hd = ld4(HIDb + (int64_t)(crow0 + p) * HID + cg * 4);
        o = make_float4(hd.x > 0.f ? acc[r][0] : 0.f, hd.y > 0.f ? acc[r][1] : 0.f,
                        hd.z > 0.f ? acc[r][2] : 0.f, hd.w > 0.f ? acc[r][3] : 0.f);
        st4(GHID + (int64_t)(crow0 + p) * HID + cg * 4, o);
      }
      *reinterpret_cast<float4*>(&gh_s[rg * 4 + r][cg * 4]) = o;
    }
  }
  __syncthreads();
  {  // g_a = g_hid @ W1;  W1 is [256][128] as stored by nn.Linear(128,256)
    const int cg = tid & 31, rg = tid >> 5;
    float acc[2][4];
    gemm_256x128(gh_s, W1, rg, cg, acc);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int p = t0 + rg * 2 + r;
      if (p < cnt)
        st4(GA + (int64_t)(crow0 + p) * D + cg * 4, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
    }
  }
}

int cell_base_of(const tm_schedule* s, int level) {
  int acc = 0;
  for (int l = 2; l < level; l += 2) acc += s->h_level_ptr[l + 1] - s->h_level_ptr[l];
  return acc;
}
}  // namespace

extern "C" int tm_gnn_forward(const tm_schedule* s, int32_t lb, int32_t le, float* H, const float* S,
                              const float* W1t, const float* b1, const float* W2t, const float* b2,
                              float* A, float* LSE, float* HIDb, void* stream) {
  TM_REQUIRE(s && s->h_level_ptr && lb >= 0 && le <= s->num_levels && lb <= le, "tm_gnn_forward: bad level range");
  TM_REQUIRE((A == nullptr) == (LSE == nullptr) && (A == nullptr) == (HIDb == nullptr),
             "tm_gnn_forward: A, LSE, HID must be all set or all NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int crow0 = cell_base_of(s, lb + (lb & 1));
  for (int l = lb; l < le; ++l) {
    const int p0 = s->h_level_ptr[l], cnt = s->h_level_ptr[l + 1] - p0;
    const bool cell = (l > 0) && !(l & 1);
    if (cnt > 0) {
      if (!cell) {
        gnn_net_fwd_kernel<<<(unsigned)cdiv(cnt, 8), 256, 0, st>>>(s->order, p0, cnt, s->net_iptr, s->net_isrc,
                                                                  S, H, l == 0);
        TM_TRY(check_launch("gnn_net_fwd"));
      } else {
        gnn_cell_fwd_kernel<<<(unsigned)cdiv(cnt, TILE), CT, 0, st>>>(s->order, p0, cnt, crow0, s->cell_iptr,
                                                                     s->cell_isrc, S, H, W1t, b1, W2t, b2, A,
                                                                     LSE, HIDb);
        TM_TRY(check_launch("gnn_cell_fwd"));
      }
    }
    if (cell) crow0 += cnt;
  }
  return 0;
}

extern "C" int tm_gnn_backward(const tm_schedule* s, const float* H, float* G, const float* W1,
                               const float* W2, const float* A, const float* LSE, const float* HIDb,
                               float* GA, float* GHID, float* GZC, void* stream) {
  TM_REQUIRE(s && s->h_level_ptr, "tm_gnn_backward: bad schedule");
  cudaStream_t st = (cudaStream_t)stream;
  SchedDev d{s->order, s->level, s->crow, s->net_iptr, s->net_optr, s->net_odst, s->cell_optr, s->cell_odst};
  int crow_end = cell_base_of(s, s->num_levels + (s->num_levels & 1));  // total cell rows
  for (int l = s->num_levels - 1; l >= 0; --l) {
    const int p0 = s->h_level_ptr[l], cnt = s->h_level_ptr[l + 1] - p0;
    const bool cell = (l > 0) && !(l & 1);
    if (cell) crow_end -= cnt;
    if (cnt <= 0) continue;
    if (!cell) {
      gnn_net_bwd_kernel<<<(unsigned)cdiv(cnt, 8), 256, 0, st>>>(d, p0, cnt, H, G, GA, A, LSE);
      TM_TRY(check_launch("gnn_net_bwd"));
    } else {
      gnn_cell_bwd_kernel<<<(unsigned)cdiv(cnt, TILE), CT, 0, st>>>(d, p0, cnt, crow_end, H, G, W1, W2, GA, A,
                                                                   LSE, HIDb, GHID, GZC);
      TM_TRY(check_launch("gnn_cell_bwd"));
    }
  }
  return 0;
}

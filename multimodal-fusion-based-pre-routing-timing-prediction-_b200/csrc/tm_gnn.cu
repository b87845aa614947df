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
constexpr int CT = 128;    // threads per CTA on a cell level

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4relu(float4 a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }

// ---------------------------------------------------------------------------------------------
// forward, level 0 and odd (net) levels: one warp per pin
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gnn_net_fwd_kernel(const int* __restrict__ order, int p0, int cnt, const int* __restrict__ f_ptr,
                   const int* __restrict__ f_src, const float* __restrict__ S, float* H) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= cnt) return;
  const int p = p0 + w;
  const int v = order[p];
  const int s = f_ptr[p], e = f_ptr[p + 1];          // level 0: empty range
  const float4 sv = ld4_stream(S + (int64_t)v * D + lane * 4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = s; i < e; i += 4) {
    float4 m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      m[q] = (i + q < e) ? ld4(H + (int64_t)f_src[i + q] * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = f4add(acc, m[q]);
  }
  if (e > s) acc = f4scale(acc, 1.f / (float)(e - s));
  st4(H + (int64_t)v * D + lane * 4, f4relu(f4add(sv, acc)));
}

// ---------------------------------------------------------------------------------------------
// tile MLP shared by forward and backward: out = epi2( epi1(in @ Wa) @ Wb ) for a 16-row tile,
//   Wa: [128][256] row-major (k-major), Wb: [256][128] row-major.
// The 256 KB of weights do not fit in shared memory next to the tiles, so they are STREAMED:
// 16 chunks of 16 KB (8 of Wa, 8 of Wb; each chunk is a contiguous run of k-rows) flow through a
// 3-stage cp.async ring, so the L2->SM weight traffic is bandwidth- not latency-bound and overlaps
// the FFMA work.  128 threads; GEMM1 thread tile 8 rows x 4 cols, GEMM2 thread tile 4 rows x 4 cols;
// tile rows are read as warp-wide broadcasts, weights as conflict-free 128-bit LDS.
// ---------------------------------------------------------------------------------------------
constexpr int CHUNK = 4096;       // floats per weight chunk (16 KB)
constexpr int NSTAGE = 3;
constexpr int NCHUNK = 16;
constexpr size_t CELL_SMEM = (size_t)(TILE * D + TILE * HID + NSTAGE * CHUNK) * sizeof(float) + TILE * sizeof(int);

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void issue_chunk(float* wbuf, const float* __restrict__ Wa,
                                            const float* __restrict__ Wb, int c, int tid) {
  if (c < NCHUNK) {
    const float* src = (c < 8) ? Wa + (size_t)c * CHUNK : Wb + (size_t)(c - 8) * CHUNK;
    float* dst = wbuf + (c % NSTAGE) * CHUNK;
#pragma unroll
    for (int i = 0; i < CHUNK / 4 / CT; ++i) cp_async16(dst + (tid + i * CT) * 4, src + (tid + i * CT) * 4);
  }
  cp_async_commit();   // always commit (possibly empty) so that wait_group<1> means "chunk c landed"
}

// Call order inside a kernel:  mlp_prologue() -> fill in_s -> mlp_tile()
__device__ __forceinline__ void mlp_prologue(float* wbuf, const float* Wa, const float* Wb, int tid) {
  issue_chunk(wbuf, Wa, Wb, 0, tid);
  issue_chunk(wbuf, Wa, Wb, 1, tid);
}

template <class Epi1, class Epi2>
__device__ __forceinline__ void mlp_tile(const float (*in_s)[D], float (*mid_s)[HID], float* wbuf,
                                         const float* Wa, const float* Wb, int tid, Epi1 epi1, Epi2 epi2) {
  const int cg1 = tid & 63, rg1 = tid >> 6;   // GEMM1: cols 4*cg1.., rows 8*rg1..
  const int cg2 = tid & 31, rg2 = tid >> 5;   // GEMM2: cols 4*cg2.., rows 4*rg2..
  float acc1[8][4], acc2[4][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc1[r][c] = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc2[r][c] = 0.f;

  for (int c = 0; c < NCHUNK; ++c) {
    cp_async_wait<1>();
    __syncthreads();                      // chunk c visible; buffer of chunk c-1 free; tiles visible
    issue_chunk(wbuf, Wa, Wb, c + 2, tid);
    const float* w = wbuf + (c % NSTAGE) * CHUNK;
    if (c < 8) {                          // GEMM1: k rows 16c .. 16c+15 of Wa, 256 columns
      const int kb = c * 16;
#pragma unroll
      for (int kk = 0; kk < 16; kk += 4) {
        float4 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(&in_s[rg1 * 8 + r][kb + kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(w + (kk + j) * HID + cg1 * 4);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float av = (j == 0) ? a[r].x : (j == 1) ? a[r].y : (j == 2) ? a[r].z : a[r].w;
            acc1[r][0] = fmaf(av, wv.x, acc1[r][0]);
            acc1[r][1] = fmaf(av, wv.y, acc1[r][1]);
            acc1[r][2] = fmaf(av, wv.z, acc1[r][2]);
            acc1[r][3] = fmaf(av, wv.w, acc1[r][3]);
          }
        }
      }
      if (c == 7) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 o = epi1(rg1 * 8 + r, cg1 * 4, make_float4(acc1[r][0], acc1[r][1], acc1[r][2], acc1[r][3]));
          *reinterpret_cast<float4*>(&mid_s[rg1 * 8 + r][cg1 * 4]) = o;
        }
      }
    } else {                              // GEMM2: k rows 32(c-8) .. +31 of Wb, 128 columns
      const int kb = (c - 8) * 32;
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        float4 a[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&mid_s[rg2 * 4 + r][kb + kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(w + (kk + j) * D + cg2 * 4);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float av = (j == 0) ? a[r].x : (j == 1) ? a[r].y : (j == 2) ? a[r].z : a[r].w;
            acc2[r][0] = fmaf(av, wv.x, acc2[r][0]);
            acc2[r][1] = fmaf(av, wv.y, acc2[r][1]);
            acc2[r][2] = fmaf(av, wv.z, acc2[r][2]);
            acc2[r][3] = fmaf(av, wv.w, acc2[r][3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    epi2(rg2 * 4 + r, cg2 * 4, make_float4(acc2[r][0], acc2[r][1], acc2[r][2], acc2[r][3]));
}

// ---------------------------------------------------------------------------------------------
// forward, even level > 0: gather + per-channel softmax-weighted sum + MLP, 16 pins per CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT)
gnn_cell_fwd_kernel(const int* __restrict__ order, int p0, int cnt, int crow0,
                    const int* __restrict__ f_ptr, const int* __restrict__ f_src,
                    const float* __restrict__ S, float* H, const float* __restrict__ W1t,
                    const float* __restrict__ b1, const float* __restrict__ W2t,
                    const float* __restrict__ b2, float* __restrict__ A, float* __restrict__ LSE,
                    float* __restrict__ HIDb) {
  extern __shared__ __align__(16) float smem[];
  float (*a_s)[D] = reinterpret_cast<float (*)[D]>(smem);
  float (*hid_s)[HID] = reinterpret_cast<float (*)[HID]>(smem + TILE * D);
  float* wbuf = smem + TILE * D + TILE * HID;
  int* v_s = reinterpret_cast<int*>(wbuf + NSTAGE * CHUNK);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * TILE;
  mlp_prologue(wbuf, W1t, W2t, tid);

  // phase 1: each warp aggregates PPW consecutive pins.  Their in-edges are one contiguous range of
  // the level-ordered edge list, walked once with warp-uniform pin boundaries; eight source rows
  // are in flight per step (online softmax per channel, lane = 4 channels).
  constexpr int PPW = TILE / (CT / 32);
  {
    const int r0 = warp * PPW;
    const int npin = min(PPW, cnt - (t0 + r0));                  // may be <= 0 on the tail tile
    const int pb = p0 + t0 + r0;
    int ptrs = 0, vv = -1;
    if (npin > 0) {
      if (lane <= npin) ptrs = f_ptr[pb + lane];
      if (lane < npin) vv = order[pb + lane];
    }
    if (lane < PPW) v_s[r0 + lane] = vv;
    int q = 0;
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float sm[4] = {0.f, 0.f, 0.f, 0.f}, tw[4] = {0.f, 0.f, 0.f, 0.f};
    auto finish_pin = [&]() {                                    // writes pin q, resets the state
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), lse = av;
      if (sm[0] > 0.f) {                                         // at least one in-edge
        av = make_float4(tw[0] / sm[0], tw[1] / sm[1], tw[2] / sm[2], tw[3] / sm[3]);
        lse = make_float4(mx[0] + logf(sm[0]), mx[1] + logf(sm[1]), mx[2] + logf(sm[2]), mx[3] + logf(sm[3]));
      }
      *reinterpret_cast<float4*>(&a_s[r0 + q][lane * 4]) = av;
      if (A) {
        st4(A + (int64_t)(crow0 + t0 + r0 + q) * D + lane * 4, av);
        st4(LSE + (int64_t)(crow0 + t0 + r0 + q) * D + lane * 4, lse);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { mx[c] = -INFINITY; sm[c] = 0.f; tw[c] = 0.f; }
      ++q;
    };
    if (npin > 0) {
      const int E0 = __shfl_sync(0xffffffffu, ptrs, 0), E1 = __shfl_sync(0xffffffffu, ptrs, npin);
      int bnd = __shfl_sync(0xffffffffu, ptrs, 1);               // end of pin q's range
      for (int i = E0; i < E1; i += 8) {
        const int idx = (i + (lane & 7) < E1) ? f_src[i + (lane & 7)] : 0;
        float4 m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int src = __shfl_sync(0xffffffffu, idx, u);
          if (i + u < E1) m[u] = ld4(H + (int64_t)src * D + lane * 4);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i + u < E1) {
            while (i + u >= bnd) { finish_pin(); bnd = __shfl_sync(0xffffffffu, ptrs, q + 1); }
            const float mv[4] = {m[u].x, m[u].y, m[u].z, m[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float nm = fmaxf(mx[c], mv[c]);
              const float sc = expf(mx[c] - nm);                 // exp(-inf) = 0 on the first edge
              const float ex = expf(mv[c] - nm);
              sm[c] = sm[c] * sc + ex;
              tw[c] = tw[c] * sc + mv[c] * ex;
              mx[c] = nm;
            }
          }
        }
      }
      while (q < npin) finish_pin();
    }
    for (int r = max(npin, 0); r < PPW; ++r)                     // rows past the end of the level
      *reinterpret_cast<float4*>(&a_s[r0 + r][lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // phases 2+3: hidden = relu(a @ W1t + b1);  h = relu(S + hidden @ W2t + b2)
  auto epi1 = [&](int row, int col, float4 acc) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + col));
    const float4 o = f4relu(make_float4(acc.x + bb.x, acc.y + bb.y, acc.z + bb.z, acc.w + bb.w));
    const int p = t0 + row;
    if (HIDb && p < cnt) st4(HIDb + (int64_t)(crow0 + p) * HID + col, o);
    return o;
  };
  auto epi2 = [&](int row, int col, float4 acc) {
    const int v = v_s[row];
    if (v >= 0) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + col));
      const float4 sv = ld4_stream(S + (int64_t)v * D + col);
      st4(H + (int64_t)v * D + col,
          f4relu(make_float4(acc.x + bb.x + sv.x, acc.y + bb.y + sv.y, acc.z + bb.z + sv.z, acc.w + bb.w + sv.w)));
    }
  };
  mlp_tile(a_s, hid_s, wbuf, W1t, W2t, tid, epi1, epi2);
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
  const int* bn_ptr;
  const int* bn_dst;
  const float* bn_w;
  const int* bc_ptr;
  const int* bc_row;
};

__device__ __forceinline__ float4 cell_edge_grad(float4 hv, float4 ga, float4 ls, float4 aa) {
  return make_float4(ga.x * expf(hv.x - ls.x) * (1.f + hv.x - aa.x), ga.y * expf(hv.y - ls.y) * (1.f + hv.y - aa.y),
                     ga.z * expf(hv.z - ls.z) * (1.f + hv.z - aa.z), ga.w * expf(hv.w - ls.w) * (1.f + hv.w - aa.w));
}
__device__ __forceinline__ float4 relu_mask(float4 hv, float4 g) {
  return make_float4(hv.x > 0.f ? g.x : 0.f, hv.y > 0.f ? g.y : 0.f, hv.z > 0.f ? g.z : 0.f, hv.w > 0.f ? g.w : 0.f);
}

// one warp per pin (level 0 and odd levels)
__global__ void __launch_bounds__(256)
gnn_net_bwd_kernel(SchedDev s, int p0, int cnt, const float* __restrict__ H, float* G,
                   const float* __restrict__ GA, const float* __restrict__ A,
                   const float* __restrict__ LSE) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= cnt) return;
  const int p = p0 + w;
  const int v = s.order[p];
  const int ns = s.bn_ptr[p], ne = s.bn_ptr[p + 1], cs = s.bc_ptr[p], ce = s.bc_ptr[p + 1];
  const int64_t off = (int64_t)v * D + lane * 4;
  float4 g = ld4(G + off);
  const float4 hv = ld4(H + off);
  for (int i = ns; i < ne; i += 4) {
    float4 m[4];
    float wq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      wq[q] = 0.f;
      m[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i + q < ne) { wq[q] = s.bn_w[i + q]; m[q] = ld4(G + (int64_t)s.bn_dst[i + q] * D + lane * 4); }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) g = f4add(g, f4scale(m[q], wq[q]));
  }
  for (int i = cs; i < ce; i += 2) {
    float4 ga[2], ls[2], aa[2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (i + q < ce) {
        const int64_t co = (int64_t)s.bc_row[i + q] * D + lane * 4;
        ga[q] = ld4(GA + co); ls[q] = ld4(LSE + co); aa[q] = ld4(A + co);
      }
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (i + q < ce) g = f4add(g, cell_edge_grad(hv, ga[q], ls[q], aa[q]));
  }
  st4(G + off, relu_mask(hv, g));
}

__global__ void __launch_bounds__(CT)
gnn_cell_bwd_kernel(SchedDev s, int p0, int cnt, int crow0, const float* __restrict__ H, float* G,
                    const float* __restrict__ W1, const float* __restrict__ W2, float* GA,
                    const float* __restrict__ A, const float* __restrict__ LSE,
                    const float* __restrict__ HIDb, float* __restrict__ GHID, float* __restrict__ GZC) {
  extern __shared__ __align__(16) float smem[];
  float (*gz_s)[D] = reinterpret_cast<float (*)[D]>(smem);
  float (*gh_s)[HID] = reinterpret_cast<float (*)[HID]>(smem + TILE * D);
  float* wbuf = smem + TILE * D + TILE * HID;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * TILE;
  // g_hid = (g_z @ W2) * (hid > 0): W2 is [128][256] as stored by nn.Linear(256,128)  -> "Wa"
  // g_a   =  g_hid @ W1:            W1 is [256][128] as stored by nn.Linear(128,256)  -> "Wb"
  mlp_prologue(wbuf, W2, W1, tid);

  // phase 1: each warp pulls the gradient of PPW consecutive pins.  gz_s accumulates, the first
  // 128 columns of gh_s hold the pins' own h rows until the MLP overwrites them.
  constexpr int PPW = TILE / (CT / 32);
  {
    const int r0 = warp * PPW;
    const int npin = min(PPW, cnt - (t0 + r0));
    const int pb = p0 + t0 + r0;
    int nptr = 0, cptr = 0, vv = 0;
    if (npin > 0) {
      if (lane <= npin) { nptr = s.bn_ptr[pb + lane]; cptr = s.bc_ptr[pb + lane]; }
      if (lane < npin) vv = s.order[pb + lane];
    }
#pragma unroll
    for (int r = 0; r < PPW; ++r) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), hv = g;
      if (r < npin) {
        const int64_t off = (int64_t)__shfl_sync(0xffffffffu, vv, r) * D + lane * 4;
        g = ld4(G + off);
        hv = ld4(H + off);
      }
      *reinterpret_cast<float4*>(&gz_s[r0 + r][lane * 4]) = g;
      *reinterpret_cast<float4*>(&gh_s[r0 + r][lane * 4]) = hv;
    }
    if (npin > 0) {
      {  // net out-edges: one contiguous range for the warp's pins
        const int E0 = __shfl_sync(0xffffffffu, nptr, 0), E1 = __shfl_sync(0xffffffffu, nptr, npin);
        int q = 0, bnd = __shfl_sync(0xffffffffu, nptr, 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        auto flush = [&]() {
          float4* dst = reinterpret_cast<float4*>(&gz_s[r0 + q][lane * 4]);
          *dst = f4add(*dst, acc);
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
          ++q;
        };
        for (int i = E0; i < E1; i += 8) {
          const bool ok = i + (lane & 7) < E1;
          const int idx = ok ? s.bn_dst[i + (lane & 7)] : 0;
          const float wt = ok ? s.bn_w[i + (lane & 7)] : 0.f;
          float4 m[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int dst = __shfl_sync(0xffffffffu, idx, u);
            if (i + u < E1) m[u] = ld4(G + (int64_t)dst * D + lane * 4);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float wu = __shfl_sync(0xffffffffu, wt, u);
            if (i + u < E1) {
              while (i + u >= bnd) { flush(); bnd = __shfl_sync(0xffffffffu, nptr, q + 1); }
              acc = f4add(acc, f4scale(m[u], wu));
            }
          }
        }
        if (q < npin) flush();
      }
      {  // cell out-edges (rare on cell levels)
        const int E0 = __shfl_sync(0xffffffffu, cptr, 0), E1 = __shfl_sync(0xffffffffu, cptr, npin);
        int q = 0, bnd = __shfl_sync(0xffffffffu, cptr, 1);
        for (int i = E0; i < E1; ++i) {
          while (i >= bnd) { ++q; bnd = __shfl_sync(0xffffffffu, cptr, q + 1); }
          const int64_t co = (int64_t)s.bc_row[i] * D + lane * 4;
          const float4 hv = *reinterpret_cast<const float4*>(&gh_s[r0 + q][lane * 4]);
          float4* dst = reinterpret_cast<float4*>(&gz_s[r0 + q][lane * 4]);
          *dst = f4add(*dst, cell_edge_grad(hv, ld4(GA + co), ld4(LSE + co), ld4(A + co)));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < PPW; ++r) {
      if (r < npin) {
        const float4 hv = *reinterpret_cast<const float4*>(&gh_s[r0 + r][lane * 4]);
        const float4 gz = relu_mask(hv, *reinterpret_cast<const float4*>(&gz_s[r0 + r][lane * 4]));
        *reinterpret_cast<float4*>(&gz_s[r0 + r][lane * 4]) = gz;
        st4(G + (int64_t)__shfl_sync(0xffffffffu, vv, r) * D + lane * 4, gz);
        st4(GZC + (int64_t)(crow0 + t0 + r0 + r) * D + lane * 4, gz);
      }
    }
  }
  auto epi1 = [&](int row, int col, float4 acc) {
    const int p = t0 + row;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < cnt) {
      const float4 hd = ld4(HIDb + (int64_t)(crow0 + p) * HID + col);
      o = make_float4(hd.x > 0.f ? acc.x : 0.f, hd.y > 0.f ? acc.y : 0.f, hd.z > 0.f ? acc.z : 0.f,
                      hd.w > 0.f ? acc.w : 0.f);
      st4(GHID + (int64_t)(crow0 + p) * HID + col, o);
    }
    return o;
  };
  auto epi2 = [&](int row, int col, float4 acc) {
    const int p = t0 + row;
    if (p < cnt) st4(GA + (int64_t)(crow0 + p) * D + col, acc);
  };
  mlp_tile(gz_s, gh_s, wbuf, W2, W1, tid, epi1, epi2);
}

int cell_smem_optin() {
  static bool done = false;   // per process; the attribute is per function and device
  if (!done) {
    TM_CUDA(cudaFuncSetAttribute(gnn_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CELL_SMEM));
    TM_CUDA(cudaFuncSetAttribute(gnn_cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CELL_SMEM));
    done = true;
  }
  return 0;
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
  TM_REQUIRE(s && s->f_ptr && s->f_src, "tm_gnn_forward: schedule lacks the level-ordered edge lists");
  TM_REQUIRE(s && s->h_level_ptr && lb >= 0 && le <= s->num_levels && lb <= le, "tm_gnn_forward: bad level range");
  TM_REQUIRE((A == nullptr) == (LSE == nullptr) && (A == nullptr) == (HIDb == nullptr),
             "tm_gnn_forward: A, LSE, HID must be all set or all NULL");
  cudaStream_t st = (cudaStream_t)stream;
  TM_TRY(cell_smem_optin());
  int crow0 = cell_base_of(s, lb + (lb & 1));
  for (int l = lb; l < le; ++l) {
    const int p0 = s->h_level_ptr[l], cnt = s->h_level_ptr[l + 1] - p0;
    const bool cell = (l > 0) && !(l & 1);
    if (cnt > 0) {
      if (!cell) {
        gnn_net_fwd_kernel<<<(unsigned)cdiv(cnt, 8), 256, 0, st>>>(s->order, p0, cnt, s->f_ptr, s->f_src, S, H);
        TM_TRY(check_launch("gnn_net_fwd"));
      } else {
        gnn_cell_fwd_kernel<<<(unsigned)cdiv(cnt, TILE), CT, CELL_SMEM, st>>>(s->order, p0, cnt, crow0, s->f_ptr, s->f_src, S,
                                                                             H, W1t, b1, W2t, b2, A, LSE, HIDb);
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
  TM_REQUIRE(s && s->h_level_ptr && s->bn_ptr && s->bc_ptr, "tm_gnn_backward: bad schedule");
  cudaStream_t st = (cudaStream_t)stream;
  TM_TRY(cell_smem_optin());
  SchedDev d{s->order, s->bn_ptr, s->bn_dst, s->bn_w, s->bc_ptr, s->bc_row};
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
      gnn_cell_bwd_kernel<<<(unsigned)cdiv(cnt, TILE), CT, CELL_SMEM, st>>>(d, p0, cnt, crow_end, H, G, W1, W2, GA, A,
                                                                   LSE, HIDb, GHID, GZC);
      TM_TRY(check_launch("gnn_cell_bwd"));
    }
  }
  return 0;
}

// N2: path-mask rasteriser on the GPU.
//
// Replaces the host preprocessing that builds `path_masks` for a batch of timing endpoints:
//   * critical-path back-trace (verilog_parser_asap7.py:1433-1450): from the endpoint, repeatedly step
//     to the FIRST predecessor (edge-insertion order of the pin graph) that sits exactly one
//     topological level lower, until the level drops below 2;
//   * rasterisation (verilog_parser_asap7.py:1302-1369): union of the bin bounding boxes of consecutive
//     pins of that path, column index = x * map_size + y, de-duplicated, ascending.
// Output is the CSR the fusion kernels consume (tm_fuse_forward*), bit-identical to the reference's
// sparse rows.  Two calls because nnz is data dependent: tm_mask_count() leaves one bitmap per
// endpoint in the workspace and returns the row sizes, the caller scans them and sizes `cols`,
// tm_mask_fill() expands the bitmaps.
#include "tm_common.cuh"

using namespace tmk;

namespace {
constexpr int MAX_BOXES = 2048;      // longest critical path handled (levels)
constexpr int MT = 128;              // threads per endpoint

// earliest in-edge (insertion order) whose source is exactly one level below its destination
__global__ void first_pred_kernel(int64_t E, const int* __restrict__ src, const int* __restrict__ dst,
                                  const int* __restrict__ level, int* __restrict__ best) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int s = src[e], d = dst[e];
  if (level[s] >= 0 && level[s] == level[d] - 1) atomicMin(&best[d], (int)e);
}

__global__ void fill_int_kernel(int64_t n, int* __restrict__ p, int v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(MT)
mask_bitmap_kernel(const int* __restrict__ src, const int* __restrict__ level, const int* __restrict__ best,
                   const int* __restrict__ endpoints, const int* __restrict__ pin_xy, int map_size, int words,
                   unsigned* __restrict__ bitmaps, int* __restrict__ counts) {
  extern __shared__ unsigned sm[];
  unsigned* bits = sm;                                       // [words]
  short* box = reinterpret_cast<short*>(sm + words);         // [MAX_BOXES][4]: x1, x2, y1, y2
  __shared__ int nbox_s, cnt_s;
  const int t = blockIdx.x, tid = threadIdx.x;
  for (int w = tid; w < words; w += MT) bits[w] = 0u;
  if (tid == 0) {
    int cur = endpoints[t], lv = level[cur], nb = 0;
    bool ok = lv >= 0;
    while (ok && lv >= 2) {
      const int e = best[cur];
      if (e == 0x7fffffff || nb >= MAX_BOXES) { ok = false; break; }   // the reference would spin forever here
      const int nxt = src[e];
      const int ax = pin_xy[2 * cur], ay = pin_xy[2 * cur + 1], bx = pin_xy[2 * nxt], by = pin_xy[2 * nxt + 1];
      box[4 * nb + 0] = (short)min(ax, bx); box[4 * nb + 1] = (short)max(ax, bx);
      box[4 * nb + 2] = (short)min(ay, by); box[4 * nb + 3] = (short)max(ay, by);
      ++nb;
      cur = nxt;
      --lv;
    }
    nbox_s = ok ? nb : -1;
    cnt_s = 0;
  }
  __syncthreads();
  const int nb = nbox_s;
  if (nb < 0) {
    if (tid == 0) counts[t] = -1;
    return;
  }
  // one work item = one bin row x of one box: set bits [x*map + y1, x*map + y2]
  for (int b = 0; b < nb; ++b) {
    const int x1 = box[4 * b], x2 = box[4 * b + 1], y1 = box[4 * b + 2], y2 = box[4 * b + 3];
    for (int x = x1 + tid; x <= x2; x += MT) {
      const int lo = x * map_size + y1, hi = x * map_size + y2;
      for (int w = lo >> 5; w <= (hi >> 5); ++w) {
        const int a = max(lo, w << 5) & 31, z = min(hi, (w << 5) + 31) & 31;
        const unsigned m = (z == 31 ? 0xffffffffu : ((1u << (z + 1)) - 1u)) & ~((1u << a) - 1u);
        atomicOr(&bits[w], m);
      }
    }
  }
  __syncthreads();
  int c = 0;
  for (int w = tid; w < words; w += MT) {
    const unsigned v = bits[w];
    bitmaps[(int64_t)t * words + w] = v;
    c += __popc(v);
  }
  c = (int)warp_sum((float)c);        // exact: c <= 65536 per endpoint
  if ((tid & 31) == 0) atomicAdd(&cnt_s, c);
  __syncthreads();
  if (tid == 0) counts[t] = cnt_s;
}

__global__ void __launch_bounds__(MT)
mask_fill_kernel(const unsigned* __restrict__ bitmaps, int words, const int* __restrict__ indptr,
                 int* __restrict__ cols) {
  __shared__ int wsum[MT];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int per = (words + MT - 1) / MT;                     // consecutive words per thread
  const int w0 = tid * per, w1 = min(words, w0 + per);
  int c = 0;
  for (int w = w0; w < w1; ++w) c += __popc(bitmaps[(int64_t)t * words + w]);
  wsum[tid] = c;
  __syncthreads();
  int off = indptr[t];
  for (int i = 0; i < tid; ++i) off += wsum[i];              // MT = 128: short, deterministic
  for (int w = w0; w < w1; ++w) {
    unsigned v = bitmaps[(int64_t)t * words + w];
    while (v) {
      const int b = __ffs(v) - 1;
      cols[off++] = (w << 5) + b;
      v &= v - 1;
    }
  }
}
}  // namespace

extern "C" size_t tm_mask_ws_bytes(int64_t n, int64_t T, int64_t map_size) {
  const int64_t words = (map_size * map_size + 31) / 32;
  return (size_t)(n * 4 + T * words * 4) + 512;
}

extern "C" int tm_mask_count(int64_t n, int64_t E, const int32_t* src, const int32_t* dst, const int32_t* level,
                             int64_t T, const int32_t* endpoints, const int32_t* pin_xy, int64_t map_size,
                             int32_t* counts, void* ws, size_t ws_bytes, void* stream) {
  TM_REQUIRE(map_size >= 1 && map_size <= 256, "tm_mask_count: map_size must be in [1,256]");
  TM_REQUIRE(ws && ws_bytes >= tm_mask_ws_bytes(n, T, map_size), "tm_mask_count: workspace too small");
  if (T <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (int)((map_size * map_size + 31) / 32);
  int* best = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  unsigned* bitmaps = reinterpret_cast<unsigned*>(best + n);
  fill_int_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(n, best, 0x7fffffff);
  TM_TRY(check_launch("mask_init"));
  if (E > 0) {
    first_pred_kernel<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(E, src, dst, level, best);
    TM_TRY(check_launch("mask_first_pred"));
  }
  const size_t sm = (size_t)words * 4 + (size_t)MAX_BOXES * 4 * sizeof(short);
  mask_bitmap_kernel<<<(unsigned)T, MT, sm, st>>>(src, level, best, endpoints, pin_xy, (int)map_size, words, bitmaps, counts);
  return check_launch("mask_bitmap");
}

extern "C" int tm_mask_fill(int64_t n, int64_t T, int64_t map_size, const int32_t* indptr, int32_t* cols, void* ws,
                            size_t ws_bytes, void* stream) {
  TM_REQUIRE(ws && ws_bytes >= tm_mask_ws_bytes(n, T, map_size), "tm_mask_fill: workspace too small");
  if (T <= 0) return 0;
  const int words = (int)((map_size * map_size + 31) / 32);
  int* best = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const unsigned* bitmaps = reinterpret_cast<const unsigned*>(best + n);
  mask_fill_kernel<<<(unsigned)T, MT, 0, (cudaStream_t)stream>>>(bitmaps, words, indptr, cols);
  return check_launch("mask_fill");
}

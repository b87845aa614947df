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

// =============================================================================================
// Row selection of a path-mask CSR for one endpoint batch, on the device, with no host round trip:
//   th.index_select(path_masks, 0, th.tensor(paths))        (src/train.py:500, once per level and batch)
// rows[T] picks mask rows (a shuffled DataLoader batch; repeats allowed: oversampled critical paths,
// train.py:377-380).  Produces the two forms the fusion kernels read:
//   * run-length form of the selected rows (forward, tm_fuse_forward_runs);
//   * their column-major transpose csc_ptr[J+1], csc_t[] with t ascending inside every column
//     (backward pull, tm_fuse_backward) -- built as a STABLE counting sort (per-block histograms, a
//     column-wise scan over the blocks, then a row-by-row fill), so the result and hence the order of the
//     backward's floating-point sums is deterministic without sorting anything.
// Every size the host needs is an upper bound it already knows (cap >= selected nnz), so the whole
// selection can sit inside a captured CUDA graph with `rows` as a graph input.
// =============================================================================================
namespace {
constexpr int SEL_R = 8;            // rows per block in the counting sort

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += y;
  }
  return v;
}

// one warp per selected row: number of runs of consecutive columns
__global__ void sel_count_runs_kernel(int T, const int* __restrict__ indptr, const int* __restrict__ cols,
                                      const int* __restrict__ rows, int* __restrict__ nruns) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const int r = rows[t], s = indptr[r], e = indptr[r + 1];
  int cnt = 0;
  for (int i = s + lane; i < e; i += 32) cnt += (i == s || cols[i] != cols[i - 1] + 1) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) nruns[t] = cnt;
}

// single block: out[0] = 0, out[i+1] = in[0] + ... + in[i]
__global__ void __launch_bounds__(1024) sel_scan_kernel(int n, const int* __restrict__ in, int* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry = 0; out[0] = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int v = i < n ? in[i] : 0;
    int x = warp_incl_scan(v, lane);
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const int w = wsum[lane];
      wsum[lane] = warp_incl_scan(w, lane) - w;
    }
    __syncthreads();
    x += wsum[warp] + carry;
    if (i < n) out[i + 1] = x;
    __syncthreads();
    if (tid == 1023) carry = x;
    __syncthreads();
  }
}

// one warp per selected row: write its runs at run_ptr[t]
__global__ void sel_fill_runs_kernel(int T, const int* __restrict__ indptr, const int* __restrict__ cols,
                                     const int* __restrict__ rows, const int* __restrict__ run_ptr,
                                     int* __restrict__ run_lo, int* __restrict__ run_hi) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const int r = rows[t], s = indptr[r], e = indptr[r + 1];
  int out = run_ptr[t];
  for (int i0 = s; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < e;
    const int c = in ? cols[i] : 0;
    const bool start = in && (i == s || cols[i - 1] + 1 != c);
    const bool last = in && (i == e - 1 || cols[i + 1] != c + 1);
    const unsigned ms = __ballot_sync(0xffffffffu, start), ml = __ballot_sync(0xffffffffu, last);
    // the run an entry starts / ends: runs opened before this chunk are counted in `out`
    if (start) run_lo[out + __popc(ms & ((1u << lane) - 1u))] = c;
    if (last) {
      // index of the run this entry closes = (#starts at or before this lane) - 1, relative to `out`, or the run
      // carried in from the previous chunk (then no start precedes it here and the index is out - 1)
      const int k = __popc(ms & ((2u << lane) - 1u)) - 1;
      run_hi[out + k] = c + 1;
    }
    out += __popc(ms);
    (void)ml;
  }
}

// counting sort by column, pass 1: per-block histogram of the columns of SEL_R consecutive selected rows
__global__ void __launch_bounds__(1024) sel_hist_kernel(int T, int J, const int* __restrict__ indptr, const int* __restrict__ cols,
                                                        const int* __restrict__ rows, int* __restrict__ hist) {
  extern __shared__ int sh[];                                    // [J]
  for (int j = threadIdx.x; j < J; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  const int t0 = blockIdx.x * SEL_R;
  for (int q = 0; q < SEL_R; ++q) {
    const int t = t0 + q;
    if (t >= T) break;
    const int r = rows[t], s = indptr[r], e = indptr[r + 1];
    for (int i = s + threadIdx.x; i < e; i += blockDim.x) sh[cols[i]] += 1;      // columns are unique inside a row
    __syncthreads();
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) hist[(size_t)blockIdx.x * J + j] = sh[j];
}

// pass 2a: per column, exclusive scan over the blocks (in place) and the column total
__global__ void sel_colscan_kernel(int nb, int J, int* __restrict__ hist, int* __restrict__ total) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  int acc = 0;
  for (int b = 0; b < nb; ++b) {
    const int v = hist[(size_t)b * J + j];
    hist[(size_t)b * J + j] = acc;
    acc += v;
  }
  total[j] = acc;
}

// pass 3: row by row inside a block (rows ascending), every entry goes to its column's next slot
__global__ void __launch_bounds__(1024) sel_scatter_kernel(int T, int J, const int* __restrict__ indptr, const int* __restrict__ cols,
                                                           const int* __restrict__ rows, const int* __restrict__ hist,
                                                           const int* __restrict__ csc_ptr, int* __restrict__ csc_t) {
  extern __shared__ int cur[];                                   // [J] next free slot of every column for this block
  for (int j = threadIdx.x; j < J; j += blockDim.x) cur[j] = csc_ptr[j] + hist[(size_t)blockIdx.x * J + j];
  __syncthreads();
  const int t0 = blockIdx.x * SEL_R;
  for (int q = 0; q < SEL_R; ++q) {
    const int t = t0 + q;
    if (t >= T) break;
    const int r = rows[t], s = indptr[r], e = indptr[r + 1];
    for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
      const int j = cols[i];
      csc_t[cur[j]] = t;
      cur[j] += 1;
    }
    __syncthreads();
  }
}
}  // namespace

extern "C" size_t tm_mask_select_ws(int64_t T, int64_t J) {
  const int64_t nb = cdiv(T > 0 ? T : 1, SEL_R);
  return (size_t)(nb * J + J + T + 64) * sizeof(int) + 1024;
}

extern "C" int tm_mask_select(int64_t T, int64_t J, const int32_t* indptr, const int32_t* cols, const int32_t* rows,
                              int32_t* run_ptr, int32_t* run_lo, int32_t* run_hi, int32_t* csc_ptr, int32_t* csc_t,
                              void* ws, size_t ws_bytes, void* stream) {
  TM_REQUIRE(T >= 0 && J > 0 && J * sizeof(int) <= 200 * 1024, "tm_mask_select: bad sizes (J = %lld)", (long long)J);
  TM_REQUIRE(ws && ws_bytes >= tm_mask_select_ws(T, J), "tm_mask_select: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)cdiv(T > 0 ? T : 1, SEL_R);
  Carver c(ws);
  int* hist = c.take<int>((size_t)nb * J);
  int* total = c.take<int>(J);
  int* nruns = c.take<int>(T + 1);
  if (T == 0) {
    TM_CUDA(cudaMemsetAsync(run_ptr, 0, sizeof(int), st));
    TM_CUDA(cudaMemsetAsync(csc_ptr, 0, sizeof(int) * (J + 1), st));
    return 0;
  }
  const unsigned wblocks = (unsigned)cdiv(T * 32, 256);
  sel_count_runs_kernel<<<wblocks, 256, 0, st>>>((int)T, indptr, cols, rows, nruns);
  TM_TRY(check_launch("sel_count_runs"));
  sel_scan_kernel<<<1, 1024, 0, st>>>((int)T, nruns, run_ptr);
  TM_TRY(check_launch("sel_scan(runs)"));
  sel_fill_runs_kernel<<<wblocks, 256, 0, st>>>((int)T, indptr, cols, rows, run_ptr, run_lo, run_hi);
  TM_TRY(check_launch("sel_fill_runs"));
  const size_t sm = (size_t)J * sizeof(int);
  TM_CUDA(cudaFuncSetAttribute(sel_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  TM_CUDA(cudaFuncSetAttribute(sel_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  sel_hist_kernel<<<nb, 1024, sm, st>>>((int)T, (int)J, indptr, cols, rows, hist);
  TM_TRY(check_launch("sel_hist"));
  sel_colscan_kernel<<<(unsigned)cdiv(J, 256), 256, 0, st>>>(nb, (int)J, hist, total);
  TM_TRY(check_launch("sel_colscan"));
  sel_scan_kernel<<<1, 1024, 0, st>>>((int)J, total, csc_ptr);
  TM_TRY(check_launch("sel_scan(columns)"));
  sel_scatter_kernel<<<nb, 1024, sm, st>>>((int)T, (int)J, indptr, cols, rows, hist, csc_ptr, csc_t);
  TM_TRY(check_launch("sel_scatter"));
  return 0;
}

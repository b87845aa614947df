// tcgen05 / TMEM GEMM core (sm_100a): C[M,N] = epi( A[M,K] . B[N,K]^T ), fp32 in / fp32 out.
//
// Operands are fp32 in HBM.  They are staged into shared memory by the CTA's threads, converted
// on the fly to bf16 in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
// matrices), and multiplied by tcgen05.mma.kind::f16 with the fp32 accumulator in tensor memory:
//   SPLIT == 1 : plain bf16 operands                          (rtol 2e-2 class, conv bf16 path)
//   SPLIT == 3 : x = hi + mid (two bf16 terms); hi*hi + hi*mid + mid*hi -- ~16-17 bit products
//   SPLIT == 6 : x = hi + mid + lo (three bf16 terms = 24 bits); the six products down to 2^-16
//                relative size are accumulated in fp32: fp32-class accuracy for the rtol 1e-3 path
// Staging through threads (instead of TMA) keeps row gathers, transposed operands (weight
// gradients) and im2col (convolutions) one template parameter away.
//
// Tile: BM = 128 rows, BN in {32,64,128} columns, BK = 32; 2-stage smem ring fed one k-block
// ahead through registers; one elected thread issues the MMAs, tcgen05.commit arrives on an
// mbarrier per stage; the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 bit),
// transposes it through shared memory and writes 128-bit row-contiguous stores.  Two or three
// CTAs are resident per SM so staging, MMA and epilogue of different tiles overlap.
#pragma once
#include <cuda_bf16.h>

#include "tm_gemm.cuh"

namespace tmk {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a tensor-core fault must not turn into a hung GPU.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (between the two 8-element k
//   groups of one MMA) | [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
// operand loaders: v[0..7] = operand(row, k0 .. k0+7), zero outside the matrix
// kTransposed selects the lane mapping of the staging loop (which index is contiguous in HBM)
// ---------------------------------------------------------------------------------------------
struct RowLoader {      // element(row, k) = A[rows ? rows[row] : row][k]   (k contiguous)
  static constexpr bool kTransposed = false;
  const float* A;
  int64_t ld;
  const int32_t* rows;
  int64_t nrows, K;
  bool vec;             // ld % 4 == 0 && 16-byte aligned base
  __device__ __forceinline__ void load8(int64_t row, int64_t k0, float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row >= nrows || k0 >= K) return;
    const float* p = A + (rows ? (int64_t)rows[row] : row) * ld + k0;
    if (vec && k0 + 8 <= K) {
      const float4 a = ld4(p), b = ld4(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k0 + i < K) v[i] = p[i];
    }
  }
};

struct ColLoader {      // element(row, k) = A[krows ? krows[k] : k][row]   (row contiguous)
  static constexpr bool kTransposed = true;
  const float* A;
  int64_t ld;
  const int32_t* krows;
  int64_t nrows, K;
  __device__ __forceinline__ void load8(int64_t row, int64_t k0, float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = 0.f;
      if (row < nrows && k0 + i < K) v[i] = A[(krows ? (int64_t)krows[k0 + i] : k0 + i) * ld + row];
    }
  }
};

// im2col view of an NHWC tensor: row = output pixel, k = tap * Cin + ci (stride 1, pad ks/2)
struct Im2colLoader8 {
  static constexpr bool kTransposed = false;
  const float* X;
  int64_t ldx;
  int H, W, Cin, ks, pad;
  int64_t nrows, K;
  bool vec;             // Cin % 8 == 0 && ldx % 4 == 0 && aligned
  __device__ __forceinline__ void load8(int64_t row, int64_t k0, float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row >= nrows || k0 >= K) return;
    const int hw = H * W;
    const int b = (int)(row / hw);
    const int rem = (int)(row - (int64_t)b * hw);
    const int y = rem / W, x = rem - y * W;
    if (vec) {
      const int tap = (int)(k0 / Cin), ci = (int)(k0 - (int64_t)tap * Cin);
      const int ty = tap / ks, tx = tap - ty * ks;
      const int yy = y + ty - pad, xx = x + tx - pad;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float* p = X + (((int64_t)b * H + yy) * W + xx) * ldx + ci;
        const float4 a = ld4(p), c = ld4(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t k = k0 + i;
        if (k < K) {
          const int tap = (int)(k / Cin), ci = (int)(k - (int64_t)tap * Cin);
          const int ty = tap / ks, tx = tap - ty * ks;
          const int yy = y + ty - pad, xx = x + tx - pad;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[i] = X[(((int64_t)b * H + yy) * W + xx) * ldx + ci];
        }
      }
    }
  }
};

// transposed im2col (weight gradients): row = k index (tap, ci), reduction index = pixel
struct Im2colColLoader {
  static constexpr bool kTransposed = true;
  const float* X;
  int64_t ldx;
  int H, W, Cin, ks, pad;
  int64_t nrows, K;     // nrows = ks*ks*Cin, K = number of pixels
  __device__ __forceinline__ void load8(int64_t row, int64_t k0, float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row >= nrows) return;
    const int tap = (int)(row / Cin), ci = (int)(row - (int64_t)tap * Cin);
    const int ty = tap / ks, tx = tap - ty * ks;
    const int hw = H * W;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t px = k0 + i;
      if (px < K) {
        const int b = (int)(px / hw);
        const int rem = (int)(px - (int64_t)b * hw);
        const int y = rem / W, x = rem - y * W;
        const int yy = y + ty - pad, xx = x + tx - pad;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[i] = X[(((int64_t)b * H + yy) * W + xx) * ldx + ci];
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// staging: fp32 -> bf16 parts into the UMMA K-major no-swizzle layout (BK = 32: 4 k-groups of 8)
//   byte offset of (row, kgroup) = (row / 8) * 512 + kgroup * 128 + (row % 8) * 16
// A warp-level work unit covers 32 (row, kgroup) items; global loads are issued one k-block ahead
// into registers (prefetch) and converted / stored after the previous MMAs released the stage.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool TRANSPOSED>
__device__ __forceinline__ void unit_coords(int u, int lane, int& row, int& kg) {
  if (TRANSPOSED) { row = (u >> 2) * 32 + lane; kg = u & 3; }
  else { row = u * 8 + (lane & 7); kg = lane >> 3; }
}

// PARTS bf16 terms of v[0..7] -> 16-byte stores at `off` of each part buffer (part stride PSTRIDE)
template <int PARTS>
__device__ __forceinline__ void split_store(float (&v)[8], uint8_t* base, uint32_t off, uint32_t pstride) {
#pragma unroll
  for (int part = 0; part < PARTS; ++part) {
    uint4 h;
    h.x = pack_bf16(v[0], v[1]); h.y = pack_bf16(v[2], v[3]); h.z = pack_bf16(v[4], v[5]); h.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(base + (size_t)part * pstride + off) = h;
    if (part + 1 < PARTS) {
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(hp[i]);
        v[2 * i] -= f.x;                 // exact: the residual is representable in fp32
        v[2 * i + 1] -= f.y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// epilogues write 4 consecutive columns of one row (vector path when the layout allows)
// ---------------------------------------------------------------------------------------------
struct PartialEpilogue {
  float* P;
  int64_t M, N;
  struct Row { int64_t r; };
  __device__ __forceinline__ Row row(int64_t m) const { return Row{m}; }
  __device__ __forceinline__ void store(const Row& rw, int64_t n, float v) const {
    P[((int64_t)blockIdx.z * M + rw.r) * N + n] = v;
  }
};

template <class EP>
__device__ __forceinline__ void store4(const EP& ep, const typename EP::Row& rw, int64_t n, int64_t N, const float (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (n + i < N) ep.store(rw, n + i, v[i]);
}
// PlainEpilogue fast path: one 128-bit store per lane when rows are 16-byte aligned
template <>
__device__ __forceinline__ void store4<PlainEpilogue>(const PlainEpilogue& ep, const PlainEpilogue::Row& rw, int64_t n,
                                                      int64_t N, const float (&v)[4]) {
  const bool vec = (n + 4 <= N) && ((ep.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) &&
                   (!(ep.flags & TM_EPI_MASK) || (((ep.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask) & 15) == 0)));
  if (!vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n + i < N) ep.store(rw, n + i, v[i]);
    return;
  }
  float4 o = make_float4(v[0], v[1], v[2], v[3]);
  if (ep.flags & TM_EPI_BIAS) { o.x += ep.bias[n]; o.y += ep.bias[n + 1]; o.z += ep.bias[n + 2]; o.w += ep.bias[n + 3]; }
  float* p = ep.C + rw.r * ep.ldc + n;
  if (ep.flags & TM_EPI_ACCUM) { const float4 c = ld4(p); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
  if (ep.flags & TM_EPI_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  if (ep.flags & TM_EPI_MASK) {
    const float4 mk = ld4(ep.mask + rw.r * ep.ldmask + n);
    o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f; o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
  }
  st4(p, o);
}

// ---------------------------------------------------------------------------------------------
// the kernel.  grid = (m tiles, n tiles, k splits)
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int parts_of(int split) { return split == 1 ? 1 : (split == 3 ? 2 : 3); }
constexpr int EPI_SCRATCH = 8 * 32 * 33 * 4;            // one 32 x 33 fp32 tile per warp

template <int BN, int SPLIT>
constexpr size_t smem_bytes() {
  size_t ring = (size_t)2 * (BM + BN) * BK * 2 * parts_of(SPLIT);
  return (ring > EPI_SCRATCH ? ring : EPI_SCRATCH) + 1024 /*alignment slack*/;
}

template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  if (CW == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
}

template <class AL, class BL, class EP, int BN, int SPLIT>
__global__ void __launch_bounds__(THREADS)
tc_gemm_kernel(AL al, BL bl, EP ep, int64_t M, int64_t N, int64_t K, int64_t k_per_split, int* __restrict__ err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int PARTS = parts_of(SPLIT);
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE = (A_BYTES + B_BYTES) * PARTS;
  constexpr int A_UNITS = BM / 8, B_UNITS = BN / 8;                 // 32-item work units per tile
  constexpr int NWARP = THREADS / 32;
  constexpr int A_PER = A_UNITS / NWARP;                            // units per warp (2)
  constexpr int B_PER = (B_UNITS + NWARP - 1) / NWARP;              // 1..4
  __shared__ __align__(8) uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
  const int64_t k_end = (k_begin + k_per_split < K) ? k_begin + k_per_split : K;
  const int nkb = (k_end > k_begin) ? (int)((k_end - k_begin + BK - 1) / BK) : 0;

  if (warp == 0) tmem_alloc(&tmem_base_s, BN);
  if (tid == 32) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t idesc = make_idesc(BN);
  bool ok = true;

  float pa[A_PER][8], pb[B_PER][8];
  auto prefetch = [&](int kb) {
    const int64_t k0 = k_begin + (int64_t)kb * BK;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int row, kg;
      unit_coords<AL::kTransposed>(warp * A_PER + i, lane, row, kg);
      al.load8(m0 + row, k0 + kg * 8, pa[i]);
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int u = warp * B_PER + i;
      if (u < B_UNITS) {
        int row, kg;
        unit_coords<BL::kTransposed>(u, lane, row, kg);
        bl.load8(n0 + row, k0 + kg * 8, pb[i]);
      }
    }
  };
  auto commit_stage = [&](uint8_t* a_base, uint8_t* b_base) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int row, kg;
      unit_coords<AL::kTransposed>(warp * A_PER + i, lane, row, kg);
      split_store<PARTS>(pa[i], a_base, (uint32_t)(row >> 3) * 512u + (uint32_t)kg * 128u + (uint32_t)(row & 7) * 16u, A_BYTES);
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int u = warp * B_PER + i;
      if (u < B_UNITS) {
        int row, kg;
        unit_coords<BL::kTransposed>(u, lane, row, kg);
        split_store<PARTS>(pb[i], b_base, (uint32_t)(row >> 3) * 512u + (uint32_t)kg * 128u + (uint32_t)(row & 7) * 16u, B_BYTES);
      }
    }
  };

  if (nkb > 0) prefetch(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= 2) ok = mbar_wait(&mma_done[s], (uint32_t)((kb >> 1) - 1) & 1u) && ok;   // stage free again
    uint8_t* a_base = smem + (size_t)s * STAGE;
    uint8_t* b_base = a_base + A_BYTES * PARTS;
    commit_stage(a_base, b_base);
    if (kb + 1 < nkb) prefetch(kb + 1);                 // in flight while the MMAs below are issued
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(a_base), sb = smem_u32(b_base);
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        const uint32_t koff = (uint32_t)ks * 256u;      // two 128-byte core matrices per K = 16
        uint64_t da[PARTS], db[PARTS];
#pragma unroll
        for (int p = 0; p < PARTS; ++p) {
          da[p] = make_desc(sa + p * A_BYTES + koff, 128, 512);
          db[p] = make_desc(sb + p * B_BYTES + koff, 128, 512);
        }
        // products in decreasing magnitude; parts: 0 = hi, 1 = mid, 2 = lo
        umma(tmem_d, da[0], db[0], idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        if (PARTS >= 2) { umma(tmem_d, da[0], db[1], idesc, 1u); umma(tmem_d, da[1], db[0], idesc, 1u); }
        if (PARTS >= 3) {
          umma(tmem_d, da[1], db[1], idesc, 1u);
          umma(tmem_d, da[0], db[2], idesc, 1u);
          umma(tmem_d, da[2], db[0], idesc, 1u);
        }
      }
      umma_commit(&mma_done[s]);
    }
  }
  if (nkb > 0) {
    const int last = nkb - 1;                           // the last commit covers every earlier MMA
    ok = mbar_wait(&mma_done[last & 1], (uint32_t)(last >> 1) & 1u) && ok;
  }
  tc_fence_after();
  if (!ok && err) atomicExch(err, 1);
  __syncthreads();                                      // every warp is done with the smem ring

  // epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (= tile rows) and one half of the columns.
  // TMEM -> registers -> 32x33 smem tile (transpose) -> 128-bit row-contiguous global stores.
  {
    constexpr int CW = (BN / 2 >= 32) ? 32 : 16;        // columns per chunk
    constexpr int LPR = CW / 4;                         // lanes per row when storing
    constexpr int RPI = 32 / LPR;                       // rows per store instruction
    float* scr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const int lane_grp = warp & 3;
    const int c_half = (warp >> 2) * (BN / 2);
#pragma unroll 1
    for (int c = c_half; c < c_half + BN / 2; c += CW) {
      float v[32];
      if (nkb > 0) {
        tmem_ld_cols<CW>(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) scr[lane * 33 + i] = v[i];
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 32 / RPI; ++it) {
        const int r = it * RPI + lane / LPR, cq = (lane % LPR) * 4;
        const int64_t m = m0 + lane_grp * 32 + r;
        const int64_t n = n0 + c + cq;
        if (m < M && n < N) {
          const float o[4] = {scr[r * 33 + cq], scr[r * 33 + cq + 1], scr[r * 33 + cq + 2], scr[r * 33 + cq + 3]};
          const typename EP::Row rw = ep.row(m);
          store4<EP>(ep, rw, n, N, o);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
}

// N tile: at most 128 so that two or three CTAs share an SM (one CTA's staging / epilogue overlaps
// another's MMAs); wider outputs use grid.y.
inline int pick_bn(int64_t N) { return N > 64 ? 128 : (N > 32 ? 64 : 32); }

template <class AL, class BL, class EP, int BN, int SPLIT>
int launch_one(const AL& al, const BL& bl, const EP& ep, int64_t M, int64_t N, int64_t K, int splits,
               int64_t k_per_split, int* err, cudaStream_t st) {
  auto kern = tc_gemm_kernel<AL, BL, EP, BN, SPLIT>;
  constexpr size_t sm = smem_bytes<BN, SPLIT>();
  static bool optin = false;
  if (!optin) {
    TM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    optin = true;
  }
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(N, BN), (unsigned)splits);
  kern<<<grid, THREADS, sm, st>>>(al, bl, ep, M, N, K, k_per_split, err);
  return check_launch("tc_gemm");
}

template <class AL, class BL, class EP>
int launch(const AL& al, const BL& bl, const EP& ep, int64_t M, int64_t N, int64_t K, int splits,
           int64_t k_per_split, int precision, int* err, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int bn = pick_bn(N);
#define TM_TC_CASE(BN_)                                                                                 \
  return precision == 2   ? launch_one<AL, BL, EP, BN_, 6>(al, bl, ep, M, N, K, splits, k_per_split, err, st) \
         : precision == 1 ? launch_one<AL, BL, EP, BN_, 3>(al, bl, ep, M, N, K, splits, k_per_split, err, st) \
                          : launch_one<AL, BL, EP, BN_, 1>(al, bl, ep, M, N, K, splits, k_per_split, err, st)
  switch (bn) {
    case 128: TM_TC_CASE(128);
    case 64: TM_TC_CASE(64);
    default: TM_TC_CASE(32);
  }
#undef TM_TC_CASE
}

}  // namespace tc
}  // namespace tmk

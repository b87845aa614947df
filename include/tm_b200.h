/*
 * tm_b200.h -- C ABI of libtm_b200.so: the B200 (sm_100a) hot path of the multimodal
 * pre-routing timing predictor (level-wise netlist GNN + layout U-Net + mask fusion).
 *
 * The reference (ZeayW/Multimodal-fusion-based-Pre-routing-Timing-Prediction-) has no
 * FFI/plugin layer: its hot path is Python calling DGL and ATen.  Each entry point below
 * therefore cites the reference *call site* (src/<file>:<line>) whose library work it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - no allocation and no ownership transfer inside: the caller passes workspaces;
 *   - all launches go to the given cudaStream_t (passed as void*), nothing synchronises
 *     unless stated;
 *   - return 0 on success, a negative TM_E* for argument errors or a positive cudaError_t;
 *     tm_last_error() returns a thread-local message for the last failure.
 * Python binding: multimodal-fusion-based-pre-routing-timing-prediction-_b200/tm_lib.py (ctypes).
 */
#ifndef TM_B200_H
#define TM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_OK 0
#define TM_EINVAL (-1)   /* bad argument                      */
#define TM_EWORKSPACE (-2) /* workspace too small             */
#define TM_EUNSUPPORTED (-3)

int tm_version(void);
const char* tm_last_error(void);
/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
long long tm_launch_count(void);
int tm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* The level schedule of one graph (all device arrays int32 unless noted).  Built by the G1
 * entry points below, consumed by the G2/G3 propagation kernels. */
typedef struct {
  int64_t n;                 /* pins                                                   */
  int32_t num_levels;
  int32_t n_cell_rows;       /* pins on even levels > 0                                */
  const int32_t* h_level_ptr;/* HOST copy of level_ptr[num_levels+1]                   */
  const int32_t* order;      /* scheduled pins by (level, id)                          */
  const int32_t* level;      /* pin -> level                                           */
  const int32_t* crow;       /* pin -> compact cell row or -1                          */
  const int32_t* net_iptr;  const int32_t* net_isrc;   /* in-edge CSRs                 */
  const int32_t* cell_iptr; const int32_t* cell_isrc;
  const int32_t* net_optr;  const int32_t* net_odst;   /* out-edge CSRs                */
  const int32_t* cell_optr; const int32_t* cell_odst;
  /* level-ordered edge lists (tm_schedule_edges), indexed by schedule position            */
  const int32_t* f_ptr;  const int32_t* f_src;         /* forward gather sources           */
  const int32_t* bn_ptr; const int32_t* bn_dst; const float* bn_w;   /* backward, net edges */
  const int32_t* bc_ptr; const int32_t* bc_row;        /* backward, cell edges (compact row) */
  /* device copies the persistent propagation kernels read (one launch loops over all levels) */
  const int32_t* level_ptr;  /* [num_levels+1] first schedule position of each level       */
  const int32_t* cell_base;  /* [num_levels+1] first compact cell row of each level        */
  int32_t* sync_flags;       /* [n + n_cell_rows] scratch: per-pin / per-cell-row ready flags of the
                              * dataflow-synchronised persistent kernels (zeroed by every call)   */
  int32_t single_driver;     /* 1: every pin on an odd level has exactly one net in-edge and its source is on an
                              * even level -- the persistent forward may then write net-level rows from their
                              * driver's producer ("push" fusion) and skip the odd levels                */
  int32_t reserved0;
} tm_schedule;

/* ------------------------------------------------------------------------------------
 * G1  graph structure: CSR build and level schedule
 *     replaces dgl.heterograph(...) + the lazy in-edge CSR DGL builds inside graph.pull
 *     (src/dataset.py:274-278, src/model.py:186-204) and Parser.cal_topo_level
 *     (src/verilog_parser_asap7.py:1452-1517).
 * ---------------------------------------------------------------------------------- */

/* Bytes of workspace tm_csr_build needs. */
size_t tm_csr_build_ws(int64_t n, int64_t e);
/* CSR keyed on `key` (dst for an in-edge CSR, src for an out-edge CSR):
 * indptr[n+1], indices[e] = `val` of every edge, ascending inside each row (duplicates kept).
 * key/val are the int64 edge lists of one edge type as DGL stores them. */
int tm_csr_build(int64_t n, int64_t e, const int64_t* key, const int64_t* val,
                 int32_t* indptr, int32_t* indices, void* ws, size_t ws_bytes, void* stream);

size_t tm_levelize_ws(int64_t n);
/* level[v] = length of the longest walk from the PI set to v (pins not reachable from a PI:
 * -1), i.e. the last frontier v appears in (verilog_parser_asap7.py:1494-1511).
 * optr/oidx: out-edge CSR over the union of both edge types.  num_levels: device int32.
 * Cooperative persistent kernel (frontier BFS + Kahn peeling with grid-wide barriers). */
int tm_levelize(int64_t n, const int32_t* optr, const int32_t* oidx, const int64_t* pis,
                int64_t n_pi, int32_t* level, int32_t* num_levels, void* ws, size_t ws_bytes,
                void* stream);

size_t tm_level_order_ws(int64_t n, int32_t num_levels);
/* order[] = scheduled pins sorted by (level, pin id); level_ptr[num_levels+1].
 * Deterministic stable counting sort. */
int tm_level_order(int64_t n, int32_t num_levels, const int32_t* level, int32_t* order,
                   int32_t* level_ptr, void* ws, size_t ws_bytes, void* stream);

/* Per-pin auxiliaries of a schedule:
 * crow[v]   = row of pin v in the compact "cell pin" buffers (pins on even levels > 0), else -1,
 *             numbered in schedule order;
 * cell_base[num_levels+1] = first compact row of each level (cell_base[num_levels] = row count);
 * n_violations (device int32) += number of in-edges whose source is not on an earlier level
 *             (such a schedule is not a topological one; the caller must refuse it). */
int tm_schedule_aux(int64_t n, int32_t num_levels, const int32_t* level, const int32_t* order,
                    const int32_t* level_ptr, const int32_t* net_iptr, const int32_t* net_isrc,
                    const int32_t* cell_iptr, const int32_t* cell_isrc, int32_t* crow,
                    int32_t* cell_base, int32_t* n_violations, void* stream);

size_t tm_schedule_edges_ws(int64_t n_sched);
/* Level-ordered edge lists for the propagation kernels (see tm_schedule): f_ptr/bn_ptr/bc_ptr
 * [n_sched+1]; f_src [E_net+E_cell]; bn_dst, bn_w [E_net]; bc_row [E_cell] (upper bounds).
 * `s` needs order, level, crow and the four CSRs filled in. */
int tm_schedule_edges(const tm_schedule* s, int64_t n_sched, int32_t* f_ptr, int32_t* f_src,
                      int32_t* bn_ptr, int32_t* bn_dst, float* bn_w, int32_t* bc_ptr,
                      int32_t* bc_row, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * dense fp32 building blocks (replace the cuBLAS SGEMMs behind nn.Linear,
 * src/model.py:15,23 and every MLP call site model.py:104,141-142,151,272,280,292)
 * ---------------------------------------------------------------------------------- */

/* epilogue flags for tm_gemm_nn */
#define TM_EPI_BIAS 1      /* C += bias[col]                                    */
#define TM_EPI_RELU 2      /* C = max(C,0)                                      */
#define TM_EPI_MASK 4      /* C *= (mask[row][col] > 0)   (ReLU backward)       */
#define TM_EPI_ACCUM 8     /* C += old C                                        */

/* C[M,N] = epi( A[M,K] @ B[K,N] ).  Row-major, leading dimensions in elements.
 * a_rows (optional, int32[M]): gather  A row i from A[a_rows[i]];
 * c_rows (optional, int32[M]): scatter C row i to   C[c_rows[i]] (mask is indexed like C). */
int tm_gemm_nn(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const int32_t* a_rows,
               const float* B, int64_t ldb, float* C, int64_t ldc, const int32_t* c_rows,
               const float* bias, const float* mask, int64_t ldmask, int flags, void* stream);

size_t tm_gemm_tn_ws(int64_t M, int64_t N, int64_t R);
/* C[M,N] (+)= A[R,M]^T @ B[R,N]  (weight gradients; reduction over the R rows, deterministic
 * split + fixed-order second pass).  a_rows/b_rows optional row gathers.  Optional bias
 * gradients: colsum_a[M] (+)= sum_r A[r,:], colsum_b[N] (+)= sum_r B[r,:].
 * accumulate != 0 adds to C / colsum_*. */
int tm_gemm_tn(int64_t M, int64_t N, int64_t R, const float* A, int64_t lda, const int32_t* a_rows,
               const float* B, int64_t ldb, const int32_t* b_rows, float* C, int64_t ldc,
               float* colsum_a, float* colsum_b, int accumulate, void* ws, size_t ws_bytes,
               void* stream);

/* out[c][r] = in[r][c]  (weight re-layout; in is rows x cols row-major) */
int tm_transpose(int64_t rows, int64_t cols, const float* in, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * G2/G3  level-wise timing propagation (replaces PathConv.forward, src/model.py:158-213,
 *        its UDFs :88-116,:138-153 and the autograd backward train.py:553)
 * ---------------------------------------------------------------------------------- */

/* Forward over levels [level_begin, level_end).  D = 128, hidden = 256 (model.py:48).
 *   H[n,128]   in/out: rows of the processed levels are written, sources are read
 *              (caller zero-fills before level 0, like train.py:342,559);
 *   S[n,128]   hoisted self terms fc_cell_self(cell_feat) / fc_net_self(net_feat) incl. bias;
 *   W1t[128,256], b1[256], W2t[256,128], b2[128]: fc_cell_neigh, weights TRANSPOSED;
 *   A[n_cell_rows,128], LSE[n_cell_rows,128], HID[n_cell_rows,256]: saved for backward
 *              (may be NULL for inference). */
size_t tm_gnn_ws_bytes(void);   /* workspace of tm_gnn_forward / tm_gnn_backward: the grid-barrier counter of the
                                 * persistent kernels (default) / the re-packed weights of the per-level kernels */
/* Implementation selector (process-wide; also env TM_GNN_IMPL=<bits>|persist|levels).  Bit 0 selects the forward
 * pass, bit 1 the backward pass; 4 (default) = auto: persistent forward, persistent backward only when the schedule
 * is wide (>= 6 000 pins per level on average) -- what measures fastest on config 2 and config 3:
 *   bit set   = ONE persistent kernel for the whole pass: 2-CTA clusters, fc_cell_neigh resident in shared memory,
 *               the tile MLP transposed on tcgen05 (fp16 two-term split, fp32 accumulate in TMEM), one grid-wide
 *               barrier per level (or, with TM_GNN_SYNC=flow, per-pin ready flags and no barrier);
 *   bit clear = one launch per level (mma.sync 3xTF32 tile MLP, weights streamed per level, PDL-chained);
 *   8         = one launch per level with the CELL levels on the cluster tile kernel (tcgen05, fp16 split); measured
 *               slower than both.
 * Returns the previous value; impl < 0 only queries. */
int tm_gnn_set_impl(int impl);
/* Level ordering inside the persistent kernels: 0 = grid barrier per level (default), 1 = per-pin ready flags
 * (dataflow; also env TM_GNN_SYNC=flow), 2 = grid barrier + net-level "push" fusion in the forward (rows of
 * single-driver net pins written by their driver's producer, odd levels skipped; also env TM_GNN_FUSE=1), 3 = sentinel
 * dataflow in the forward (H pre-filled with 0xFFFFFFFF words, every gathered 16-byte piece validates itself: no flag,
 * no fence; the backward uses the flags).  All alternatives measured slower than 0 on B200.
 * Returns the previous value; flow < 0 only queries. */
int tm_gnn_set_sync(int flow);
/* Grid-wide barriers inside the last tm_gnn_forward / tm_gnn_backward call of this thread (persistent kernel:
 * one per non-empty level; per-level kernels: 0, they synchronise by kernel boundaries). */
int tm_gnn_last_barriers(void);
/* Diagnostics: with clocks != NULL (device int64 [grid CTAs <= 148][slots], zero-filled by the caller) thread 0 of
 * every CTA of the persistent kernels adds the SM cycles it spent per phase (setup, barrier wait, gather, MMA 1,
 * epilogue 1, MMA 2, exchange, tail; net-level wait / body / tail; slots = 16).  NULL disables. */
int tm_gnn_set_profile(void* clocks);
int tm_gnn_forward(const tm_schedule* s, int32_t level_begin, int32_t level_end, float* H,
                   const float* S, const float* W1t, const float* b1, const float* W2t,
                   const float* b2, float* A, float* LSE, float* HID, void* ws, size_t ws_bytes,
                   void* stream);

/* Backward over all levels in reverse.
 *   G[n,128]   in: dLoss/dH contributions from the head (zero elsewhere);
 *              out: dLoss/d(pre-activation) of every scheduled pin == dLoss/dS;
 *   W1[256,128], W2[128,256]: fc_cell_neigh weights as stored by nn.Linear;
 *   GA[n_cell_rows,128] scratch; GHID[n_cell_rows,256], GZC[n_cell_rows,128] out: operands of
 *   the hoisted weight-gradient GEMMs (dW1 = GHID^T A, dW2 = GZC^T HID). */
int tm_gnn_backward(const tm_schedule* s, const float* H, float* G, const float* W1,
                    const float* W2, const float* A, const float* LSE, const float* HID,
                    float* GA, float* GHID, float* GZC, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * G5  mask fusion (replaces  path_mask.to_dense()*feat_map  and  fcn(path_map),
 *     src/train.py:500-501 + src/model.py:272)
 * ---------------------------------------------------------------------------------- */
/* out[t,0:D] = bias + sum_{j in mask row rows[t]} F[j] * Wt[j,:]     (Wt = fcn.weight^T, [J,D])
 * rows: optional int32[T] selecting mask rows (NULL = identity). D must be 128.
 * ld_out: row stride of out (lets the result land inside cat(h_gnn,h_cnn,h_global)). */
int tm_fuse_forward(int64_t T, int64_t J, int64_t D, const int32_t* mask_indptr,
                    const int32_t* mask_cols, const int32_t* rows, const float* F,
                    const float* Wt, const float* bias, float* out, int64_t ld_out, void* stream);
/* The same forward from the RUN-LENGTH form of the selected mask rows: endpoint t owns runs
 * run_ptr[t]..run_ptr[t+1], run r covers columns [run_lo[r], run_hi[r]).  Builds the fp64 prefix table
 * P[j] = sum_{i<j} F[i]*Wt[i,:] in ws (tm_fuse_runs_ws(J) bytes) and sums P[hi]-P[lo] per run:
 * 2 table rows per run instead of one weight row per column. */
size_t tm_fuse_runs_ws(int64_t J);
int tm_fuse_forward_runs(int64_t T, int64_t J, int64_t D, const int32_t* run_ptr, const int32_t* run_lo,
                         const int32_t* run_hi, const float* F, const float* Wt, const float* bias,
                         float* out, int64_t ld_out, void* ws, size_t ws_bytes, void* stream);
/* Column-major pull (deterministic, no atomics): csc_ptr[J+1], csc_t[nnz] list the positions t
 * (0..T-1) whose mask contains column j.
 *   dWt[j,:] = F[j] * sum_t g[t,:];   dF[j] = sum_t sum_c g[t,c] * Wt[j,c]. */
int tm_fuse_backward(int64_t T, int64_t J, int64_t D, const int32_t* csc_ptr, const int32_t* csc_t,
                     const float* g, int64_t ld_g, const float* F, const float* Wt, float* dWt,
                     float* dF, void* stream);

/* ------------------------------------------------------------------------------------
 * N2  path-mask rasteriser (replaces the host code that builds `path_masks`:
 *     find_critical_path, verilog_parser_asap7.py:1433-1450, and the bounding-box rasterisation,
 *     :1302-1369).  src/dst: ALL pin-graph edges in insertion order (the order networkx iterates
 *     predecessors in); level: longest-path levels (tm_levelize / cal_topo_level); pin_xy: int32 [n][2]
 *     bins; endpoints: int32 [T].  Row t of the result = ascending columns x*map_size+y covered by the
 *     union of the bounding boxes of consecutive pins on endpoint t's critical path.
 *     tm_mask_count -> counts[t] (-1: a pin on the path has no predecessor one level below, where the
 *     reference would loop forever); the caller builds indptr = exclusive scan(counts) and sizes cols;
 *     tm_mask_fill writes cols.  Both use the same workspace (tm_mask_ws_bytes), untouched in between.
 * ---------------------------------------------------------------------------------- */
size_t tm_mask_ws_bytes(int64_t n, int64_t T, int64_t map_size);
int tm_mask_count(int64_t n, int64_t E, const int32_t* src, const int32_t* dst, const int32_t* level,
                  int64_t T, const int32_t* endpoints, const int32_t* pin_xy, int64_t map_size,
                  int32_t* counts, void* ws, size_t ws_bytes, void* stream);
int tm_mask_fill(int64_t n, int64_t T, int64_t map_size, const int32_t* indptr, int32_t* cols, void* ws,
                 size_t ws_bytes, void* stream);

/* Row selection of a path-mask CSR for one endpoint batch -- th.index_select(path_masks, 0, th.tensor(paths)),
 * src/train.py:500 (and :213, test.py:201) -- entirely on the device: no host round trip, every size a bound the
 * host already has, so it can be replayed inside a CUDA graph with `rows` as an input.
 *   rows[T]: mask row (path id) of every endpoint of the batch, repeats allowed (oversampled paths, train.py:377-380);
 *   run_ptr[T+1], run_lo / run_hi [cap]: run-length form of the selected rows (tm_fuse_forward_runs);
 *   csc_ptr[J+1], csc_t[cap]: column-major transpose, t ascending inside a column (tm_fuse_backward), built as a
 *            stable counting sort: deterministic without sorting;
 *   cap >= sum of the selected rows' lengths (e.g. T * longest row), J * 4 bytes <= 200 KB of shared memory. */
size_t tm_mask_select_ws(int64_t T, int64_t J);
int tm_mask_select(int64_t T, int64_t J, const int32_t* indptr, const int32_t* cols, const int32_t* rows,
                   int32_t* run_ptr, int32_t* run_lo, int32_t* run_hi, int32_t* csc_ptr, int32_t* csc_t,
                   void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * G6  head helpers (src/model.py:280-292, src/train.py:513-522)
 * ---------------------------------------------------------------------------------- */
/* dst[i, col0:col0+w] = src[rows ? rows[i] : i, 0:w]   (builds cat(h_gnn,h_cnn,h_global)) */
int tm_gather_cols(int64_t T, int64_t w, const float* src, int64_t lds, const int32_t* rows,
                   float* dst, int64_t ldd, int64_t col0, void* stream);
/* dst[rows[i], 0:w] += src[i, col0:col0+w]  (atomic: rows may repeat) */
int tm_scatter_add_cols(int64_t T, int64_t w, const float* src, int64_t lds, int64_t col0,
                        const int32_t* rows, float* dst, int64_t ldd, void* stream);
/* out[c] (+)= sum_r X[rows ? rows[r] : r, c]  (bias gradients; fixed reduction order) */
size_t tm_colsum_ws(int64_t R, int64_t C);
int tm_colsum(int64_t R, int64_t C, const float* X, int64_t ld, const int32_t* rows, float* out,
              int accumulate, void* ws, size_t ws_bytes, void* stream);
/* out[c_rows[m] * ldo] = act(X[a_rows[m], 0:K] . w + bias[0]): nn.Linear(K, 1) (mlp_fuse's last layer, model.py:292),
 * one warp per row. */
int tm_rowdot(int64_t M, int64_t K, const float* X, int64_t ldx, const int32_t* a_rows, const float* w,
              const float* bias, float* out, int64_t ldo, const int32_t* c_rows, int relu, void* stream);
/* tm_colsum (accumulate = 0) that also returns max |X[rows]| in absmax[0] (C % 4 == 0, 16-byte aligned rows). */
int tm_colsum_absmax(int64_t R, int64_t C, const float* X, int64_t ld, const int32_t* rows, float* out,
                     float* absmax, void* ws, size_t ws_bytes, void* stream);
/* loss[0] = mean((pred-y)^2); grad[i] = 2*(pred[i]-y[i])/T * grad_scale  (nn.MSELoss) */
int tm_mse(int64_t T, const float* pred, const float* y, float* loss, float* grad,
           float grad_scale, void* stream);
/* N4: the per-batch statistics the reference reads back one .item() at a time (train.py:513-549):
 * out8 = [mse, R2 (torchmetrics R2Score), correct, tp, fn, tn, fp, T] with predicted critical <=>
 * required - pred < 0 (judge_critical, train.py:391-395) against label (int64, 0 = non-critical).
 * required / label may be NULL (regression statistics only). */
int tm_step_metrics(int64_t T, const float* pred, const float* arrival, const float* required,
                    const int64_t* label, float* out8, void* stream);

/* ------------------------------------------------------------------------------------
 * G4  U-Net / LayoutNet image branch, fp32 NHWC (replaces the cuDNN/ATen calls behind
 *     src/Unet.py:16-21,53,75-77 and src/model.py:227-243)
 *     Activations are NHWC with an explicit pixel stride `ld` (elements) so that a tensor may
 *     live inside a wider concat buffer (Unet.py:67 torch.cat becomes free).
 * ---------------------------------------------------------------------------------- */
int tm_nchw_to_nhwc(int64_t B, int64_t C, int64_t H, int64_t W, const float* in, float* out,
                    int64_t ld, void* stream);
int tm_nhwc_to_nchw(int64_t B, int64_t C, int64_t H, int64_t W, const float* in, int64_t ld,
                    float* out, void* stream);
/* Conv2d weight (Cout,Cin,k,k) -> fprop layout wf[k*k][Cin][Cout] and dgrad layout
 * wb[k*k][Cout][Cin] with flipped taps (either may be NULL). */
int tm_conv_pack_weight(int64_t Cout, int64_t Cin, int64_t k, const float* w, float* wf, float* wb,
                        void* stream);
/* inverse of the fprop layout: dw (Cout,Cin,k,k) = unpack(dwf[k*k][Cin][Cout]) */
int tm_conv_unpack_wgrad(int64_t Cout, int64_t Cin, int64_t k, const float* dwf, float* dw,
                         void* stream);
/* y[b,y,x,co] = bias[co] + sum_{tap,ci} x[b,y+dy,x+dx,ci] * wf[tap][ci][co]; stride 1, pad k/2.
 * flags: TM_EPI_RELU optional.  The same call with wb computes the data gradient. */
int tm_conv2d_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                   const float* x, int64_t ldx, const float* wf, const float* bias, float* y,
                   int64_t ldy, int flags, void* stream);
size_t tm_conv2d_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k);
/* dwf[tap][ci][co] = sum_{b,y,x} x[b,y+dy,x+dx,ci] * dy[b,y,x,co];  dbias[co] = sum dy (optional) */
int tm_conv2d_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                         const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dwf,
                         float* dbias, void* ws, size_t ws_bytes, void* stream);

/* ConvTranspose2d(k=2,s=2) weight (Cin,Cout,2,2) -> wt[Cin][4*Cout] ((dy,dx,co) fastest) and
 * its transpose wtT[4*Cout][Cin]. */
int tm_convt_pack_weight(int64_t Cin, int64_t Cout, const float* w, float* wt, float* wtT,
                         void* stream);
int tm_convt_unpack_wgrad(int64_t Cin, int64_t Cout, const float* dwt, float* dw, void* stream);
/* y[b,2y+dy+oy,2x+dx+ox,co] = bias[co] + sum_ci x[b,y,x,ci]*wt[ci][(dy,dx,co)]  (Unet.py:53,57-63)
 * y is (B,Hy,Wy,*) with pixel stride ldy; (oy,ox) is the F.pad offset. */
int tm_convt2x2_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, const float* x,
                     int64_t ldx, const float* wt, const float* bias, float* y, int64_t ldy,
                     int64_t Hy, int64_t Wy, int64_t oy, int64_t ox, void* stream);
/* dx[b,y,x,ci] = sum_{dy,dx,co} dyo[b,2y+dy+oy,2x+dx+ox,co] * wtT[(dy,dx,co)][ci] */
int tm_convt2x2_dgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                           const float* dyo, int64_t lddy, int64_t Hy, int64_t Wy, int64_t oy,
                           int64_t ox, const float* wtT, float* dx, int64_t lddx, void* stream);
size_t tm_convt2x2_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
/* dwt[ci][(dy,dx,co)] = sum x*dyo ; dbias[co] = sum over the 2H x 2W window of dyo */
int tm_convt2x2_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                           const float* x, int64_t ldx, const float* dyo, int64_t lddy, int64_t Hy,
                           int64_t Wy, int64_t oy, int64_t ox, float* dwt, float* dbias, void* ws,
                           size_t ws_bytes, void* stream);

size_t tm_bn_ws(int64_t npix, int64_t C);
/* Train-mode BatchNorm2d + ReLU (Unet.py:17-18,20-21): batch statistics over B*H*W,
 * y = relu(gamma*(x-mean)*invstd+beta); running stats updated with `momentum`, unbiased var.
 * save_mean/save_invstd [C] kept for backward.  y_bf16 (optional): compact bf16 copy [npix][C] of y written by
 * the same pass (the TMA operand of the next convolution in bf16 mode); dx_bf16 likewise for the backward.
 * stats_part (optional): nparts fp64 partials [nparts][C][2] (sum, sum of squares) already produced by the
 * convolution's epilogue (tm_conv3x3_bf16 stats): the statistics pass over x is skipped. */
int tm_bn_relu_forward(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, float momentum,
                       float eps, float* y, int64_t ldy, float* save_mean, float* save_invstd,
                       void* y_bf16, const void* stats_part, int64_t nparts, void* ws, size_t ws_bytes, void* stream);
/* The same pair with nn.BatchNorm2d in EVAL mode (module.eval()): y = relu(gamma * (x - running_mean) *
 * rsqrt(running_var + eps) + beta); save_mean / save_invstd receive the statistics used (inference only). */
int tm_bn_relu_eval(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* gamma, const float* beta,
                    const float* running_mean, const float* running_var, float eps, float* y, int64_t ldy,
                    float* save_mean, float* save_invstd, void* y_bf16, void* stream);
/* Backward of the pair: dy is the gradient w.r.t. the ReLU output y.
 * dx = BN'( dy * (y>0) ), dgamma, dbeta [C].  dx may be NULL when only dx_bf16 is wanted; y may be NULL:
 * the ReLU mask is then rebuilt bit-for-bit from x, the saved statistics, gamma and beta (one tensor less to
 * read, and the forward need not store y: tm_bn_relu_forward accepts y = NULL when y_bf16 is given). */
int tm_bn_relu_backward(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* y,
                        int64_t ldy, const float* dy, int64_t lddy, const float* gamma, const float* beta,
                        const float* save_mean, const float* save_invstd, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, void* dx_bf16, void* ws, size_t ws_bytes, void* stream);

/* OutConv (Unet.py:74-75): nn.Conv2d(C, 1, kernel_size=1) as streaming kernels, exact fp32.
 * forward y[p] = x[p,:] . w + bias; dgrad dx[p,c] = dy[p] * w[c]; wgrad dw[c] = sum_p x[p,c] dy[p], dbias = sum_p dy[p]. */
int tm_conv1x1_c1_forward(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* w, const float* bias,
                          float* y, int64_t ldy, void* stream);
int tm_conv1x1_c1_dgrad(int64_t npix, int64_t C, const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx,
                        void* stream);
size_t tm_conv1x1_c1_wgrad_ws(int64_t C);
int tm_conv1x1_c1_wgrad(int64_t npix, int64_t C, const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dw,
                        float* dbias, void* ws, size_t ws_bytes, void* stream);

/* 2x2 stride-2 pooling (Unet.py:89-91, model.py:222-224). mode 0 = max (idx: uint8 argmax
 * saved for backward, first maximum in row-major window order), 1 = avg (idx unused).
 * flags: TM_EPI_RELU applies ReLU after pooling (OutConv, Unet.py:76-77). */
int tm_pool2x2_forward(int64_t B, int64_t H, int64_t W, int64_t C, int mode, const float* x,
                       int64_t ldx, float* y, int64_t ldy, uint8_t* idx, int flags, void* stream);
/* dx (B,H,W,C) fully written (zeros where no gradient flows; odd trailing row/col get 0).
 * With TM_EPI_RELU, dy is first masked by (y>0).  add (optional, pixel stride ldadd): a second gradient of the
 * pooled tensor's source (the U-Net skip connection, Unet.py:67) summed into dx by the same pass. */
int tm_pool2x2_backward(int64_t B, int64_t H, int64_t W, int64_t C, int mode, const float* dy,
                        int64_t lddy, const float* y, int64_t ldy, const uint8_t* idx, float* dx,
                        int64_t lddx, int flags, const float* add, int64_t ldadd, void* stream);
/* dx = dy * (y > 0) * (slope where y <= 0)   elementwise helper for bare ReLU / LeakyReLU layers
 * (LayoutNet, model.py:219-220).  y is the activation OUTPUT. n elements, contiguous. */
int tm_leaky_relu_backward(int64_t n, const float* y, const float* dy, float slope, float* dx,
                           void* stream);
int tm_leaky_relu_forward(int64_t n, const float* x, float slope, float* y, void* stream);
/* dst[i*ldd + c] += src[i*lds + c]  for c < C  (adds a gradient living in a strided buffer) */
int tm_add_strided(int64_t npix, int64_t C, const float* src, int64_t lds, float* dst, int64_t ldd,
                   void* stream);

/* ------------------------------------------------------------------------------------
 * Tensor-core variants (tcgen05.mma kind::f16, accumulator in TMEM) of the dense contractions.
 * Same meaning as tm_gemm_nn / tm_gemm_tn / tm_conv2d_nhwc / tm_conv2d_wgrad_nhwc above.
 *   precision 0: operands rounded to bf16 (fp32 accumulate)            -- rtol 2e-2 class
 *   precision 1: operands split into two bf16 terms, three MMAs        -- ~16-bit products
 *   precision 2: operands split into three bf16 terms (24 bits), six MMAs -- fp32-class
 *   precision 3: 3xTF32 -- kind::tf32 on the raw fp32 words (hi) plus the residual x - trunc_tf32(x) (lo),
 *                hi*hi + hi*lo + lo*hi, ~21-bit products -- fp32-class, the default rtol 1e-3 path
 *   precision 4: single-pass TF32
 *   err: optional device int32, set to 1 if a tensor-core barrier timed out (never expected).
 * tm_tc_gemm_nn: b_is_nk != 0 means B is given as [N,K] row-major (an nn.Linear weight as
 * stored), otherwise [K,N] row-major like tm_gemm_nn.
 * ---------------------------------------------------------------------------------- */
int tm_tc_gemm_nn(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const int32_t* a_rows,
                  const float* B, int64_t ldb, int b_is_nk, float* C, int64_t ldc,
                  const int32_t* c_rows, const float* bias, const float* mask, int64_t ldmask,
                  int flags, int precision, int* err, void* stream);
/* Backward of the FIRST layer of an MLP whose input has 1 or 2 columns (fc_net_self.layers.0 = Linear(2,256),
 * model.py:50), fused into the data-gradient GEMM of the second layer: with dh = (G[g_rows] @ W2) * (H > 0)
 * (never written) it returns db1[n] = sum_m dh[m,n] and dW1[n,c] = sum_m dh[m,n] * X[x_rows[m], c].
 * W2t: the second layer's weight transposed to [N,K] row-major (N = hidden width <= 256, K = its outputs);
 * H: [M,N] hidden activations; 3xTF32; deterministic (fixed-order partial sums in ws). */
size_t tm_tc_mlp1_bwd_ws(int64_t N);
int tm_tc_mlp1_bwd_fused(int64_t M, int64_t N, int64_t K, const float* G, int64_t ldg, const int32_t* g_rows,
                         const float* W2t, const float* H, int64_t ldh, const float* X, int64_t ldx,
                         const int32_t* x_rows, int64_t kx, const float* W1, const float* b1, float* dW1,
                         float* db1, void* ws, size_t ws_bytes, int* err, void* stream);
/* The same MLP (first layer Linear(kx <= 2, K)) with the hidden layer GENERATED inside the operand loaders
 * instead of stored (2 FMAs per element; H = NULL above then regenerates the ReLU mask from W1, b1):
 *   forward:  out[out_rows[m], 0:N] = relu(X[x_rows[m]] @ W1^T + b1) @ W2^T + b2     (W2: [N,K] as stored)
 *   wgrad2:   dW2[Mo, hid] = G[g_rows]^T @ relu(X[x_rows] @ W1^T + b1)               (ws: tm_tc_gemm_tn_ws) */
/* The forward of that MLP for the reference's sizes (hidden 256, 128 outputs: fc_net_self, model.py:46,151) as ONE
 * fused kernel: generator warps write the hidden tile straight into fp16 two-term-split operand planes, W2 is split
 * once and stays resident in shared memory, tcgen05.mma with the accumulators in TMEM (csrc/tm_selfmlp.cu).
 * ws: tm_selfmlp_ws_bytes(). */
size_t tm_selfmlp_ws_bytes(void);
/* Its second-layer weight gradient dW2[128][256] = G[g_rows]^T relu(W1 X[x_rows] + b1) as ONE fused kernel of the same
 * kind (the hidden layer generated as the B operand, G transposed into the A operand on the way in, contraction over
 * the rows on tcgen05 with the accumulators resident in TMEM for the whole kernel).  A contraction over rows cannot
 * use per-row scales: `gmax` = device scalar holding max|G| over the rows (tm_colsum_absmax computes it in the pass
 * that takes the bias gradient).  ws: tm_selfmlp_wgrad2_ws_bytes(). */
/* ... and its first-layer gradients db1[256], dW1[256][kx] = reductions of dh = (G[g_rows] W2) * (W1 x + b1 > 0) over
 * the rows; dh is never written (per-lane column accumulators in registers).  W2: [128][256] as stored.
 * ws: tm_selfmlp_bwd1_ws_bytes(). */
size_t tm_selfmlp_bwd1_ws_bytes(void);
int tm_selfmlp_gen_bwd1(int64_t M, const float* G, int64_t ldg, const int32_t* g_rows, const float* X, int64_t ldx,
                        const int32_t* x_rows, int64_t kx, const float* W1, const float* b1, const float* W2,
                        float* dW1, float* db1, void* ws, size_t ws_bytes, void* stream);
/* The same product for an MLP whose hidden activations H were STORED (fc_cell_self, 36 inputs): the hidden-layer
 * gradient DH[r] = (G[g_rows[m]] @ W2) * (H[r] > 0), r = h_rows ? h_rows[m] : m, written out (it feeds the first-layer
 * weight gradient).  Replaces the masked data-gradient GEMM of the general MLP backward. */
/* First layer of such an MLP: HID[m, 0:256] = relu(X[x_rows[m], 0:kin] @ W1^T + b1), kin % 4 == 0, kin <= 48;
 * rowmax (optional, [M]): max of row m's outputs.  Second layer: out[out_rows[m], 0:128] = HID[h_rows[m]] @ W2^T + b2
 * with the per-row operand scale taken from rowmax (ws: tm_selfmlp_ws_bytes()). */
size_t tm_selfmlp_lin1_ws_bytes(void);
int tm_selfmlp_lin1_relu(int64_t M, const float* X, int64_t ldx, const int32_t* x_rows, int64_t kin, const float* W1,
                         const float* b1, float* HID, int64_t ldh, float* rowmax, void* ws, size_t ws_bytes,
                         void* stream);
int tm_selfmlp_rows_forward(int64_t M, const float* H, int64_t ldh, const int32_t* h_rows, const float* rowmax,
                            const float* W2, const float* b2, float* out, int64_t ldo, const int32_t* out_rows,
                            void* ws, size_t ws_bytes, void* stream);
size_t tm_selfmlp_rows_dh_ws_bytes(void);
int tm_selfmlp_rows_dh(int64_t M, const float* G, int64_t ldg, const int32_t* g_rows, const float* W2, const float* H,
                       int64_t ldh, const int32_t* h_rows, float* DH, int64_t lddh, void* ws, size_t ws_bytes,
                       void* stream);
size_t tm_selfmlp_wgrad2_ws_bytes(void);
int tm_selfmlp_gen_wgrad2(int64_t M, const float* G, int64_t ldg, const int32_t* g_rows, const float* X, int64_t ldx,
                          const int32_t* x_rows, int64_t kx, const float* W1, const float* b1, const float* gmax,
                          float* dW2, void* ws, size_t ws_bytes, void* stream);
int tm_selfmlp_gen_forward(int64_t M, const float* X, int64_t ldx, const int32_t* x_rows, int64_t kx,
                           const float* W1, const float* b1, const float* W2, const float* b2, float* out,
                           int64_t ldo, const int32_t* out_rows, void* ws, size_t ws_bytes, void* stream);
int tm_tc_mlp2_smallk_forward(int64_t M, int64_t K, int64_t N, const float* X, int64_t ldx, const int32_t* x_rows,
                              int64_t kx, const float* W1, const float* b1, const float* W2, const float* b2,
                              float* out, int64_t ldo, const int32_t* out_rows, int precision, int* err,
                              void* stream);
int tm_tc_mlp2_smallk_wgrad2(int64_t Mo, int64_t hid, int64_t R, const float* G, int64_t ldg, const int32_t* g_rows,
                             const float* X, int64_t ldx, const int32_t* x_rows, int64_t kx, const float* W1,
                             const float* b1, float* dW2, int precision, void* ws, size_t ws_bytes, int* err,
                             void* stream);
/* At most `cap` persistent CTAs for the tcgen05 GEMM / convolution launches of the calling thread (0 = one per SM);
 * returns the previous value.  Lets a caller keep SMs free for a latency-critical kernel chain on another stream. */
int tm_tc_set_grid_cap(int cap);
size_t tm_tc_gemm_tn_ws(int64_t M, int64_t N, int64_t R);
int tm_tc_gemm_tn(int64_t M, int64_t N, int64_t R, const float* A, int64_t lda, const int32_t* a_rows,
                  const float* B, int64_t ldb, const int32_t* b_rows, float* C, int64_t ldc,
                  int accumulate, int precision, void* ws, size_t ws_bytes, int* err, void* stream);
int tm_tc_conv2d_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                      const float* x, int64_t ldx, const float* wf, const float* bias, float* y,
                      int64_t ldy, int flags, int precision, int* err, void* stream);
/* The same convolution with the K = k*k*Cin range split over CTAs when the layer has too few 128-pixel tiles to
 * fill the GPU (the 64 x 64 and 32 x 32 maps of one 256 x 256 design); partial tiles in `ws`, folded in a fixed
 * order.  Falls through to tm_tc_conv2d_nhwc when no split pays (ws may then be NULL / 0 bytes). */
size_t tm_tc_conv2d_splitk_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k, int precision);
int tm_tc_conv2d_nhwc_splitk(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                             const float* x, int64_t ldx, const float* wf, const float* bias, float* y,
                             int64_t ldy, int flags, int precision, void* ws, size_t ws_bytes, int* err,
                             void* stream);
size_t tm_tc_conv2d_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k);
int tm_tc_conv2d_wgrad_nhwc(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t k,
                            const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dwf,
                            int precision, void* ws, size_t ws_bytes, int* err, void* stream);

/* ------------------------------------------------------------------------------------
 * A6 (bf16 mode)  TMA-fed tcgen05 3x3 convolutions on bf16 NHWC activations
 * Replaces nn.Conv2d(k=3, padding=1, bias=False) of DoubleConv (src/Unet.py:15-22) and its
 * autograd (data + weight gradients) when the image branch runs with bf16 tensor-core operands
 * (BASELINE config 4).  Activations are compact bf16 [B][H][W][C], C a multiple of 16; H and W
 * powers of two.  One cp.async.bulk.tensor box per (128-pixel tile, tap): the box shifted by the
 * tap offset, out-of-bounds zero fill = the convolution's padding.  fp32 accumulate in TMEM,
 * fp32 outputs (they feed the batch-norm statistics).
 * ---------------------------------------------------------------------------------- */
int tm_conv3x3_bf16_supported(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
/* fp32 rows (stride ldx, C channels) -> compact bf16 rows of Cp >= C channels (zero padded) */
int tm_to_bf16_rows(int64_t npix, int64_t C, const float* x, int64_t ldx, void* out, int64_t Cp, void* stream);
/* Narrow layers run as 64-channel layers over rows of P horizontally adjacent pixels (every TMA row a full
 * 128-byte line; the packed weights are block sparse).  P for a convolution reading Cin and producing Cout
 * channels per pixel on images W wide: */
int tm_conv3x3_bf16_pack(int64_t W, int64_t Cin, int64_t Cout);
/* nn.Conv2d weight w [Cout][Cin][3][3] fp32 -> bf16 operand q [9][P*Nc][P*KcP] of tm_conv3x3_bf16:
 * dgrad = 0: forward (Nc = Cout, K = Cin padded to KcP); dgrad = 1: data gradient (Nc = Cin, K = Cout padded
 * to KcP, taps reversed).  P = tm_conv3x3_bf16_pack(W, K padded, Nc). */
int tm_conv3x3_pack_bf16(int64_t Cout, int64_t Cin, const float* w, void* q, int64_t P, int64_t KcP, int dgrad,
                         void* stream);
/* y[pix, 0:N] (fp32, row stride ldy) = sum_{tap,c} xb[pix + tap, c] * W[tap][n][c]; bias optional,
 * flags: TM_EPI_RELU.  The data gradient is the same call on bf16(dy) with the dgrad operand. */
/* stats (optional; tm_conv3x3_bf16_stats_bytes(N, P) bytes; bias must be NULL, P * N <= 128): the batch-norm
 * statistics of the output, taken from the fp32 accumulators by the epilogue -- fp64 partial sums / sums of squares,
 * layout [slots = SMs * 4][P][N][2], consumed by tm_bn_relu_forward(stats_part, nparts = slots * P). */
size_t tm_conv3x3_bf16_stats_bytes(int64_t N, int64_t P);
int tm_conv3x3_bf16(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t N, int64_t P, const void* xb, const void* wq,
                    const float* bias, float* y, int64_t ldy, int flags, void* stats, int* err, void* stream);
/* nn.ConvTranspose2d(k=2, s=2) (Unet.py:53) on the same TMA path.  Weight w [Cin][Cout][2][2] fp32 ->
 * wf bf16 [4*Cout][Cin] (forward operand) and wd bf16 [4][Cin][Cout] (data-gradient operand; either may be NULL).
 * forward: y[b,2y+dy,2x+dx,co] = bias[co] + sum_ci xb[b,y,x,ci] w[ci][co][dy][dx]: ONE tap, the accumulator row of
 *   an input pixel is its 2x2 output window, scattered by the epilogue into (B,2H,2W,*) rows of stride ldy;
 * dgrad: dx[b,y,x,ci] = sum dyb[b,2y+dy,2x+dx,co] w[ci][co][dy][dx]: four taps, each reading every other pixel of
 *   the compact bf16 gradient [B][2H][2W][Cout] through a tensor map with element strides (1,2,2,1). */
int tm_convt2x2_bf16_supported(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
int tm_convt2x2_pack_bf16(int64_t Cin, int64_t Cout, const float* w, void* wf, void* wd, void* stream);
int tm_convt2x2_bf16(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, const void* xb, const void* wf,
                     const float* bias, float* y, int64_t ldy, int* err, void* stream);
int tm_convt2x2_bf16_dgrad(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, const void* dyb, const void* wd,
                           float* dx, int64_t lddx, int* err, void* stream);
/* dw[ci][co][dy][dx] (torch layout) = sum xb[b,y,x,ci] * dyb[b,2y+dy,2x+dx,co]; the bias gradient is tm_colsum of dy */
size_t tm_convt2x2_bf16_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
int tm_convt2x2_bf16_wgrad(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout, const void* xb, const void* dyb,
                           float* dw, void* ws, size_t ws_bytes, int* err, void* stream);
/* dw[co][ci][ky][kx] (torch layout, ci < Cin_real) = sum_pix dyb[pix, co] * xb[pix + (ky-1,kx-1), ci] */
size_t tm_conv3x3_bf16_wgrad_ws(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
int tm_conv3x3_bf16_wgrad(int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t Cin_real, int64_t Cout,
                          const void* xb, const void* dyb, float* dw, void* ws, size_t ws_bytes, int* err,
                          void* stream);

/* ------------------------------------------------------------------------------------
 * N4  fused Adam (torch.optim.Adam defaults, src/train.py:431-435,555)
 * ---------------------------------------------------------------------------------- */
int tm_adam_step(int64_t n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                 void* stream);
/* The same update for EVERY parameter tensor in ONE launch.  table: device array of
 * {float* p; const float* g; float* m; float* v; int64 n} (40 bytes each); chunk c of tm_adam_chunk() elements
 * starts at element chunk_off[c] of tensor chunk_tensor[c]. */
int tm_adam_chunk(void);
int tm_adam_multi(int64_t n_chunks, const void* table, const int32_t* chunk_tensor, const int32_t* chunk_off, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                  void* stream);
/* loss[0] = NaN if the device flag *err (set by a tensor-core kernel whose mbarrier wait timed out) is non-zero:
 * a failed tile can then not pass silently into the optimizer; the host still clears / reports the flag. */
int tm_poison_on_error(float* loss, const int32_t* err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TM_B200_H */

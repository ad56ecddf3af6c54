/*
 * acm_b200.h  --  C ABI of libacm_b200.so, the B200 (sm_100a) kernels behind the ACM
 * graph-convolution layer of SitaoLuan/ACM-GNN.
 *
 * The reference has no FFI: its "plugin API" is the Python class
 * `GraphConvolution` (ACM-Pytorch/models/layers.py:14-242, ACM-Geometric/layers.py:13-120)
 * whose forward is a chain of torch ops.  Each entry point below replaces one group of
 * those torch calls; the reference line(s) it stands in for are cited per function.  The
 * binding a maintainer adds on the reference side is a ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into memory owned by the caller (torch caching
 *     allocator); the library never allocates, frees or retains device memory;
 *   - process-global state: a launch counter, a thread-local error string and the TUNING KNOBS set
 *     through the acm_set_* entry points (gather mode, narrow-row hint, mix_bwd occupancy / ring,
 *     GEMM store path).  The knobs select between implementations with identical results; they are
 *     plain globals read at launch time -- set them once at start-up, not concurrently with launches;
 *   - `stream` is a cudaStream_t; every call is asynchronous on it, no internal
 *     device synchronisation;
 *   - return value 0 = success, otherwise a cudaError_t or ACM_ERR_* code and
 *     acm_last_error_string() describes it (thread local); nothing throws or aborts;
 *   - storage dtype `T` of feature tables: ACM_F32 or ACM_BF16; accumulation is fp32;
 *   - `fp` is the feature width padded to one of {8,16,32,64,128,256}; `f` <= fp the true
 *     width (out_features).  Padding columns are zero.
 */
#ifndef ACM_B200_H_
#define ACM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACM_F32 0
#define ACM_BF16 1

#define ACM_ERR_BAD_ARG 10001
#define ACM_ERR_UNSUPPORTED 10002

/* GEMM implementation selector */
#define ACM_GEMM_SIMT 0   /* CUDA-core fp32-FMA tiles: exact-fp32 parity mode, any shape */
#define ACM_GEMM_TCGEN05 1 /* tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), TMA fed */

/* Channel-attention parameter pack (fp32, device), identical layout for values and for
 * gradients.  Channel order: 0 low, 1 high, 2 identity(mlp), 3 structure.
 *   [0          , 4*fp)      a_vec[k][fp]   att_vec_low/high/mlp, att_struc_low  (layers.py:45-62)
 *   [4*fp       , 4*fp+16)   att_vec[j][k]  row-major 4x4 slots, KxK used        (layers.py:64-67)
 *   [4*fp+16    , 8*fp+16)   ln_gamma[k][fp]  layer_norm_*.weight                (layers.py:52-59)
 *   [8*fp+16    , 12*fp+16)  ln_beta[k][fp]   layer_norm_*.bias
 * VALUE pack only (derived tail, written by acm_pack_params when LayerNorm is live):
 *   [12*fp+16   , 16*fp+16)  gamma*a [k][fp]
 *   [16*fp+16   , 16*fp+24)  sum_f beta*a [4], sum_f gamma*a [4]   (+8 spare)
 */
#define ACM_PACK_FLOATS(fp) (12 * (fp) + 16)        /* gradient pack */
#define ACM_PACK_FLOATS_VALUE(fp) (16 * (fp) + 32)  /* value pack incl. derived tail */

int acm_version(void);
const char* acm_last_error_string(void);
/* number of kernel launches issued by this library in this process (for bench.py's
 * gpu_launches claim) */
int64_t acm_launch_count(void);

/* ---- operator construction --------------------------------------------------------
 * Replaces the adjacency preprocessing the reference does with dense torch / scipy:
 * ACM-Pytorch/utils.py:421-438,626-628 (normalize_tensor(eye + A.to_dense())) and
 * ACM-Geometric/utils.py:5-19, train.py:76-80. */

/* rowptr[int64, n+1] from the row ids of a row-major-sorted (coalesced) COO. */
int acm_csr_rowptr(const int64_t* row_sorted, int64_t nnz, int64_t n, int64_t* rowptr, void* stream);

/* Degree normalisation of M = A + I given CSR multiplicities mult[nnz] (fp32):
 * rowsum_i = sum_j m_ij ; rinv_i = 1/rowsum_i (inf -> 0) ; w_ij = fl32(rinv_i * m_ij).
 * Bit-exact with torch.pow(rowsum,-1) / torch.mm(diag(r_inv), mx) on fp32. */
int acm_degree_normalise(const int64_t* rowptr, const float* mult, int64_t n,
                         float* rowsum, float* rinv, float* w, void* stream);

/* Values of the transposed operator on a SYMMETRIC sparsity pattern: for every stored
 * (i,j) writes w_t[pos(j,i)] = w[pos(i,j)].  *not_symmetric (device int, zeroed by the
 * caller) is set to 1 if some (j,i) is missing, in which case the caller must build an
 * explicit transpose. */
int acm_csr_transpose_values(const int64_t* rowptr, const int32_t* col, const float* w, int64_t n,
                             float* w_t, int* not_symmetric, void* stream);

/* fp32 [rows, cols] (row stride ld_src) -> T [rows, ld_dst], zero-filling columns
 * cols..ld_dst-1.  The bf16 staging copy of the layer input. */
int acm_cast_pad(const float* src, int64_t rows, int64_t cols, int64_t ld_src,
                 void* dst, int dst_dtype, int64_t ld_dst, void* stream);

/* ---- parameter staging ---------------------------------------------------------------------
 * One launch builds everything the kernels consume from the reference's parameter tensors
 * (layers.py:41-67): wcat [fin,3fp] T = [W_low|W_high|W_mlp] zero padded, wcat_t [3fp,ldt] T its
 * transpose (may be NULL), pack (layout above).  a_vecs / ln_gamma / ln_beta are HOST arrays of
 * k_channels device pointers (ln_* may be NULL when LayerNorm is not live). */
int acm_pack_params(int dtype, int fin, int f, int fp, int k_channels, int ln_live, int ldt,
                    const float* w_low, const float* w_high, const float* w_mlp,
                    const float* const* a_vecs, const float* att_vec,
                    const float* const* ln_gamma, const float* const* ln_beta,
                    void* wcat, void* wcat_t, float* pack, void* stream);
/* The inverse for gradients: dwcat [fin,3fp] / dpack -> dW_* [fin,f], da_k [f], d att_vec [K,K],
 * d gamma_k / d beta_k [f] (HOST arrays of device pointers; entries may be NULL to skip). */
int acm_unpack_grads(int fin, int f, int fp, int k_channels, int ln_live,
                     const float* dwcat, const float* dpack,
                     float* dw_low, float* dw_high, float* dw_mlp,
                     float* const* da, float* datt_vec, float* const* dgamma, float* const* dbeta,
                     void* stream);

/* ---- dense feature transforms -------------------------------------------------------
 * X [n, fin] (row stride ldx), Wcat = [W_low | W_high | W_mlp] zero-padded to
 * [fin(ldw rows used: fin), 3*fp]; WcatT its transpose [3*fp, ldx]. */

/* layers.py:163-165,179-194: HL,HH,HI = mm(X, weight_{low,high,mlp}) in ONE pass over X.
 * Writes h_lh [n, 2*fp] = [HL | HH] (the gather table) and h_i [n, fp].  relu_lh != 0
 * applies relu to the [HL|HH] part (variant=True, layers.py:178-184). */
int acm_gemm_xw_fwd(int impl, int dtype, const void* x, int64_t ldx, const void* wcat, const void* wcat_t,
                    void* h_lh, void* h_i, int64_t n, int64_t fin, int64_t fp, int relu_lh, void* stream);

/* Multi-GPU (1-D row partition): the same forward GEMM with the all-gather FUSED into its
 * epilogue.  peer_tables is a HOST array of n_peers (<= 8) device pointers, the base of every
 * rank's [N_pad, 2*fp] gather table mapped into this process (symmetric memory over NVLink);
 * each finished [HL|HH] row is stored into all of them at row index row_off + local row, so the
 * exchange overlaps the GEMM and no separate collective runs (tcgen05 / bf16 path only).
 * multicast_table: NULL, or the NVSwitch multicast (NVLS) address of the same table -- then one
 * multimem.st per 16 bytes replaces the n_peers unicast stores and the switch replicates the row
 * into every rank's copy (egress 1x instead of (n_peers-1)x the produced bytes).
 * acm_mix_bwd takes the same (peer_tables, n_peers, peer_row_off, multicast_table) for the
 * backward table. */
int acm_gemm_xw_fwd_push(const void* x, int64_t ldx, const void* wcat_t, void* const* peer_tables, int n_peers,
                         int64_t row_off, void* multicast_table, void* h_i, int64_t n, int64_t fin, int64_t fp,
                         int relu_lh, void* stream);

/* autograd of the three torch.mm: dWcat[fin, 3*fp] (fp32, zeroed by caller; accumulated
 * atomically over split-K slices) = X^T . dH,  dH [n, 3*fp] = [dHL | dHH | dHI]. */
int acm_gemm_bwd_dw(int impl, int dtype, const void* x, int64_t ldx, const void* dh,
                    float* dwcat, int64_t n, int64_t fin, int64_t fp, void* stream);

/* dX [n, fin] (fp32, row stride lddx) = dH . Wcat^T. */
int acm_gemm_bwd_dx(int impl, int dtype, const void* dh, const void* wcat, const void* wcat_t, int64_t ldwt,
                    float* dx, int64_t lddx, int64_t n, int64_t fin, int64_t fp, void* stream);

/* Generic building blocks of the aggregate-first order (A(XW) = (AX)W):
 * C[m,n] (T, row stride ldc) = A[m,k] . B, B given as b_kn [k,n] (CUDA-core path) and/or K-major
 * b_nk [n,k] (tcgen05 path); optional relu. */
int acm_gemm_ab(int impl, int dtype, const void* a, int64_t lda, const void* b_kn, int64_t ldb_kn,
                const void* b_nk, int64_t ldb_nk, void* c, int64_t ldc,
                int64_t m, int64_t n, int64_t k, int relu, void* stream);
/* y[m,n] (bf16, row stride ldy) = relu?(x[m,k] . w_nk[n,k]^T + bias[n]): the nn.Linear (+ F.relu) of the reference's
 * MLP helper (ACM-Pytorch/models/layers.py:245-285; the acmgcn++ branch xX = relu(mlpX(x)), models.py:116-122) on the
 * tcgen05 path.  x, w_nk bf16 (w_nk is nn.Linear's own [out, in] layout), bias fp32 or NULL; n and k multiples of 8,
 * row strides multiples of 8 elements.  Its weight gradient is acm_gemm_atb(dY, x). */
int acm_linear_fwd(const void* x, int64_t ldx, const void* w_nk, int64_t ldw, const float* bias,
                   void* y, int64_t ldy, int64_t m, int64_t n, int64_t k, int relu, void* stream);
/* C[m,n] (fp32, row stride ldc, zeroed by caller, atomically accumulated) += A[k_rows,m]^T . B[k_rows,n] */
int acm_gemm_atb(int impl, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                 float* c, int64_t ldc, int64_t k_rows, int64_t m, int64_t n, void* stream);
/* Z = A_low . X and D = X - Z for the own rows (table = X of all nodes, T [*, fp]); one gather
 * of the INPUT row per stored edge instead of the 2*out_features wide [HL|HH] row. */
int acm_spmm_agg_first(int dtype, int fp, int64_t n_rows, int64_t row0,
                       const int64_t* rowptr, const int32_t* col, const float* val,
                       const void* table, void* z_out, void* d_out,
                       const int32_t* long_rows, int n_long, const float* long_acc, void* stream);

/* Aggregate-first layer, out_features padded to 256, bf16 storage: the three GEMMs above AND the
 * attention / mix epilogue of acm_spmm_mix_fwd's pre-aggregated mode in ONE tcgen05 launch
 *   [S_L | S_H | HI] = [Z W_L | D W_H | X W_I]   (fp32 accumulators stay in TMEM)
 *   Y = out_scale * sum_k att_k relu(.)_k ,  att = softmax(sigmoid(relu(.)_k . a_k) att_vec / 3)
 * replacing torch.mm x3 + relu + attention3 + the mix of layers.py:163-165,188-204,94-119 without the
 * [S_L|S_H|HI] round trip through HBM.  z, d, x: bf16 [n_rows, ldx] (k = padded input width, multiple
 * of 8, <= 256); wcat_t: bf16 [3*fp, ldw] K-major (acm_pack_params); pack: value pack.  Outputs: y
 * (fp32 or bf16, row stride ldy), att [n_rows,3], h_i bf16 [n_rows, fp] (pre-relu HI; required, the
 * epilogue reads it back); for the backward only (may be NULL for inference): s_lh bf16 [n_rows, 2*fp] =
 * pre-relu [S_L|S_H], sig [n_rows,3].
 * 3 channels, no LayerNorm, variant 0 only; returns ACM_ERR_UNSUPPORTED for fp != 256. */
int acm_fused_agg_fwd(const void* z, const void* d, const void* x, int64_t ldx,
                      const void* wcat_t, int64_t ldw, const float* pack,
                      int64_t n_rows, int k, int f, int fp, float out_scale,
                      void* y, int y_dtype, int64_t ldy, void* s_lh, void* h_i,
                      float* att, float* sig, void* stream);

/* Degree skew.  Rows with more than 256 stored edges ("long rows": local ids in long_rows,
 * ascending) are aggregated by this segment-parallel pass -- one lane group per segment
 * [seg_e0, seg_e1) of <= 256 edges, seg_long = index of the segment's row in long_rows --
 * into acc_out [n_long, halves*fp] (fp32, zeroed by the caller, atomically accumulated).  The
 * row kernels above take (long_rows, n_long, long_acc) and read these sums instead of walking
 * the edges.  halves = 2 for the [L|H] tables (row width 2*fp), 1 for single fp-wide rows. */
int acm_spmm_long_rows(int dtype, int fp, int halves, int64_t n_seg, const int32_t* seg_long,
                       const int64_t* seg_e0, const int64_t* seg_e1, const int32_t* col, const float* val,
                       const void* table, float* acc_out, void* stream);

/* ---- fused aggregation + channel attention + mix (THE hot kernel) -------------------
 * layers.py:176-204 + attention3/attention4 (94-152) in one launch:
 *   acc  = sum_j val_ij * table[col_ij]            (one gather of the [HL|HH] row per edge)
 *   S_L  = rowscale_i*acc_L ; S_H = table_H[i] - rowscale_i*acc_H     (A_high = I - A_low)
 *   variant 0: O_L,O_H = relu(S_L),relu(S_H) ; variant 1 (table already relu'd): O = S
 *   O_I = relu(h_i) ; optional 4th channel O_S given (already relu'd)
 *   z_k = [LayerNorm_k](O_k).a_k ; s = sigmoid(z) ; att = softmax(s.att_vec/K)
 *   Y   = out_scale * sum_k att_k O_k
 * Rows: this call computes rows [0,n_rows) whose global ids are row0+r; `table` is the
 * full (all-gathered) table indexed by global column ids.  val/rowscale may be NULL (=1).
 * o_save/sig may be NULL (inference).
 * y is fp32 (the reference boundary) or, y_dtype = ACM_BF16, bf16 inter-layer activations
 * (the next layer's input cast folded into this epilogue; acm_mix_bwd accepts g in bf16 likewise).
 * rowptr == col == NULL selects the pre-aggregated mode of the aggregate-first order: the
 * own rows of `table` already hold [S_L | S_H] and only the epilogue runs.
 * row_order (may be NULL = natural order): a permutation of [0, n_rows) giving the order in which
 * the rows are PROCESSED.  With narrow rows several rows share a warp (32 / (fp/8) of them) and the
 * warp walks max(degree) edges; an order that puts rows of equal degree next to each other (sorted
 * by degree inside windows of a few thousand consecutive rows, so that the streamed accesses stay
 * local) removes that divergence.  Results do not depend on it.  Ignored for fp >= 256. */
int acm_spmm_mix_fwd(int dtype, int fp, int f, int64_t n_rows, int64_t row0,
                     const int64_t* rowptr, const int32_t* col, const float* val, const float* rowscale,
                     const void* table, const void* h_i, const void* o_s,
                     const float* pack, int k_channels, int ln_live, int variant, float out_scale,
                     void* y, int y_dtype, int64_t ldy, void* o_save, float* att, float* sig,
                     const int32_t* long_rows, int n_long, const float* long_acc,
                     const int32_t* row_order, void* stream);

/* Epilogue of the tcgen05 `tn` GEMMs (forward X.Wcat, dX, acm_gemm_ab): 1 (default) = every thread writes 32-byte
 * pieces of its own output row with 256-bit stores straight from the TMEM registers whenever the output layout
 * allows (32-byte aligned rows, column counts in multiples of 16 bf16 / 8 fp32) and K >= 128; bit 0 clear = always go
 * through the shared-memory transposition tile.  Bit 1 set = keep 8 epilogue warps for store-bound short-K products
 * (default: 16 warps there).  Bit 2 set = acm_fused_agg_fwd releases its first TMEM region before the second epilogue
 * pass (measured slower, opt-in; mixes bf16-rounded S_H).  A/B switches: bits 0-1 give bit-identical results. */
int acm_set_gemm_direct_store(int on);

/* Register/occupancy trade-off of mix_bwd_kernel (plain 3-channel mode): 2 (default) or 3 resident
 * CTAs per SM. */
int acm_set_mix_bwd_occupancy(int min_blocks_per_sm);
/* cp.async shared-memory ring for the streamed inputs of mix_bwd_kernel (bf16 tables): 1 on (default), 0 off. */
int acm_set_mix_bwd_ring(int on);

/* Gather implementation of acm_spmm_mix_fwd: 1 (default) = cp.async ring in shared memory (each
 * lane keeps 8 neighbour rows in flight without register staging), 0 = LDG register staging,
 * 2 = cp.async.bulk ring (one 1-KB bulk copy per neighbour row, width 256 only; measured slower
 * than mode 1, kept as an opt-in), 3 (default) = TMA row gather (cp.async.bulk.tensor ... tile::gather4:
 * FOUR 1-KB neighbour rows per request of the TMA engine into a 2-stage shared-memory ring per warp;
 * width 256 in bf16 -- measured 34.6 vs 37.4 ms for mode 1 on the headline graph; other shapes fall back to
 * mode 1), 4 = 3 plus the same staging in acm_spmm_agg_first (512-byte rows: measured SLOWER, 24.4 vs
 * 20.6 ms, opt-in).  All modes give bit-identical results. */
int acm_set_gather_mode(int mode);

/* Row-local backward of the attention/mix/relu part (autograd of layers.py:94-152,
 * 185-204).  g = dL/dY [n_rows, f].  Writes t_lh [n_rows, 2*fp] = [dS_L | dS_H] (the table
 * the transposed aggregation gathers), dh_all[:, 2fp:3fp] = dHI, optional dos_pre (grad
 * of the structure channel before its relu) and atomically accumulates the parameter
 * gradients into dpack (same layout as pack; zeroed by caller).
 * o_lh [n_rows, 2*fp]: variant 0 -> either the saved [O_L|O_H] or the pre-relu [S_L|S_H] (the
 * kernel applies the relu on load, which is the identity on already-relu'd values: the
 * aggregate-first order passes its [S_L|S_H] table and saves no second copy); variant 1 -> [O_L|O_H].
 * table_mode 0: t_lh rows are [dS_L | dS_H] (2*fp wide).  table_mode 1 (variant 1 without LayerNorm,
 * fp >= 64): the rank-structured table of acm_spmm_t_bwd_rank1 -- with the relu before the aggregation
 * dO_k = c att_k G + dz_k a_k^T, so the table is  T g[table_rows][fp]  followed by
 * float4 {c att_L, c att_H, dz_L, dz_H}[table_rows]  (one allocation of table_rows*(fp*sizeof(T)+16) bytes;
 * row r of this call is written at index peer_row_off + r): half the bytes to gather and to exchange. */
int acm_mix_bwd(int dtype, int fp, int f, int64_t n_rows,
                const void* g, int g_dtype, int64_t ldg, const void* o_lh, const void* h_i, const void* o_s,
                const float* att, const float* sig, const float* pack,
                int k_channels, int ln_live, int variant, float out_scale,
                void* t_lh, int table_mode, int64_t table_rows, void* dh_all, void* dos_pre, float* dpack,
                void* const* peer_tables, int n_peers, int64_t peer_row_off, void* multicast_table, void* stream);

/* Transposed aggregation (autograd of torch.spmm(adj_low,.) / torch.spmm(adj_high,.)):
 *   dHL = A_low^T dS_L ; dHH = dS_H - A_low^T dS_H     -> dh_all[:, 0:2fp]
 * (rowptr_t,col_t,val_t) is the CSR of A_low^T (same arrays as A_low when the pattern is
 * symmetric, with acm_csr_transpose_values).  variant 1: p_table = the relu'd forward
 * table; the result is masked by p > 0 (relu before aggregation).  row_order: as in
 * acm_spmm_mix_fwd, for the rows of A_low^T. */
int acm_spmm_t_bwd(int dtype, int fp, int64_t n_rows, int64_t row0,
                   const int64_t* rowptr_t, const int32_t* col_t, const float* val_t,
                   const void* t_table, const void* p_table, void* dh_all,
                   const int32_t* long_rows, int n_long, const float* long_acc,
                   const int32_t* row_order, void* stream);

/* The same transposed aggregation from the rank-structured table written by acm_mix_bwd in table_mode 1
 * (variant 1, no LayerNorm):  (A^T dO_k)[i,:] = sum_j w_ji (c att_k[j]) G[j,:] + (sum_j w_ji dz_k[j]) a_k^T
 * -- ONE gather of the fp-wide G row plus four scalars per stored edge serves both channels.  pack = the
 * layer's value pack (a_L, a_H); p_table = the relu'd forward table (mask).  No long-row side pass: the
 * caller uses acm_spmm_t_bwd when the transposed operator has rows with more than 256 edges. */
int acm_spmm_t_bwd_rank1(int dtype, int fp, int64_t n_rows, int64_t row0,
                         const int64_t* rowptr_t, const int32_t* col_t, const float* val_t,
                         const void* g_table, int64_t table_rows, const float* pack, const void* p_table,
                         void* dh_all, void* stream);

/* Plain single-table aggregation out = [relu](A . table), table T [*, fp]; used for the
 * structure channel relu(mm(adj_low_unnormalized, struc_low)) (layers.py:207-209) and its
 * transpose.  out dtype out_dtype, row stride ld_out, f_out valid columns. */
int acm_spmm_plain(int dtype, int out_dtype, int fp, int64_t n_rows,
                   const int64_t* rowptr, const int32_t* col, const float* val,
                   const void* table, void* out, int64_t ld_out, int f_out, int relu, void* stream);

/* ---- loss glue of the train step ----------------------------------------
 * Fused F.log_softmax + NLLLoss on the training rows (ACM-Pytorch/utils.py:567-568,
 * ACM-Geometric/train.py:133-134), value and gradient in one pass:
 *   *loss_sum += scale * sum_{mask_i} (logsumexp(x_i) - x_i[label_i])     (zeroed by caller)
 *   dlogits_i  = scale * (softmax(x_i) - onehot(label_i)) on masked rows, 0 elsewhere (may be NULL)
 * mask: uint8 per row (NULL = every row); scale = 1/|train| for the reference's mean reduction. */
int acm_nll_log_softmax(const float* logits, int64_t ld, int64_t n_rows, int n_classes,
                        const int64_t* labels, const uint8_t* mask, float scale,
                        float* loss_sum, float* dlogits, int64_t ld_d, void* stream);

/* Inter-layer glue: y = dropout(relu(x), p) [+ add] in one pass, replacing the
 * reference's F.relu / F.dropout / "+ xX" launches between the two layers
 * (ACM-Pytorch/models/models.py:160-164, ACM-Geometric/models.py:70-74).  x, add (may be NULL), y:
 * contiguous arrays of `total` elements of `dtype`, 16-byte aligned.  mask (may be NULL when no
 * gradient is needed): ceil(total / 8) bytes, bit j of byte t = "element 8 t + j passes" (kept by
 * the dropout AND, with relu, x > 0).  p in [0, 1); p > 0 draws Philox4x32-10 bits from the DEVICE
 * state rng_state = {seed, offset} (element e kept iff philox(seed, [e / 4, offset])[e % 4] >=
 * floor(p 2^32)) and advances the offset by one on the stream, so graph replays draw fresh masks.
 * out_dtype = dtype, or ACM_F32 with dtype ACM_BF16 when `add` (and then y) is fp32 -- torch's type
 * promotion of "bf16 activations + fp32 xX".  Not torch's random stream: the host mirror only uses
 * p > 0 when asked to. */
int acm_glue_fwd(int dtype, int out_dtype, const void* x, const void* add, void* y, uint8_t* mask,
                 int64_t total, int relu, float p, uint64_t* rng_state, void* stream);
/* dx = mask ? g * scale : 0   (scale = 1 / (1 - p)); the gradient of `add` is g itself.  g has
 * out_dtype, dx has dtype (g is rounded to dtype first, as autograd's cast does). */
int acm_glue_bwd(int dtype, int out_dtype, const void* g, const uint8_t* mask, void* dx, int64_t total,
                 float scale, void* stream);

/* cudaLimitMaxL2FetchGranularity (32/64/128 bytes) of the current device: narrow rows
 * (out_features <= 16 -> 64-byte table rows) over-fetch at the default granularity. */
int acm_set_l2_fetch_granularity(int bytes);

/* Narrow-row gathers (padded width <= 32: one 64..128-byte table row per stored edge, layer 1 of
 * every reference model): 1 = issue the neighbour-row loads with the PTX L2::64B prefetch-size
 * hint (ld.global.nc.L1::no_allocate.L2::64B) instead of plain read-only loads (default: 1).
 * Results are bit-identical.  Background: these rows move ~1.6x their algorithmic bytes through
 * DRAM (profiles/README.md); the hint recovers 1-5 % of the two layer-1 gather kernels. */
int acm_set_narrow_row_hint(int on);

#ifdef __cplusplus
}
#endif
#endif /* ACM_B200_H_ */

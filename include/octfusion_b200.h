/*
 * octfusion_b200 -- C ABI of the B200 (sm_100a) kernels behind the OctFusion denoising
 * U-Net hot path.
 *
 * The reference (octree-nn/octfusion) has NO native layer on this path: every operator below
 * is a sequence of ATen library calls issued from Python.  This header is the
 * boundary a maintainer of the reference would bind (ctypes stub: INTEGRATION.md); each entry
 * point names the reference function (file:line under the reference tree) whose arithmetic it
 * replaces.
 *
 * Conventions
 *   - plain device pointers and sizes; no torch / ATen types; caller owns every buffer
 *     (outputs and workspaces included); nothing is allocated, freed or synchronised inside
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant
 *     per stream, CUDA-graph capturable
 *   - return value: 0 = launched; negative = rejected before any launch (OF_E_*); the
 *     text of the last error of the calling thread is available from of_last_error()
 *   - dtype: OF_F32 (0) = float activations, OF_BF16 (1) = __nv_bfloat16 activations;
 *     accumulation is always fp32; norm statistics: fp32 partial sums per (32-row chunk, 4 channels), combined in fp64
 *   - all row strides (ld*) are in ELEMENTS
 */
#ifndef OCTFUSION_B200_H_
#define OCTFUSION_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OF_F32 0
#define OF_BF16 1

#define OF_OK 0
#define OF_E_ARG (-1)       /* inconsistent / unsupported argument combination            */
#define OF_E_UNSUPPORTED (-2) /* shape not supported by this entry point (use the other one) */
#define OF_E_CUDA (-3)      /* a CUDA runtime call failed (text in of_last_error)          */

const char* of_last_error(void);
int of_version(void);              /* ABI version, bumped on any signature change          */
int of_num_sms(void);              /* multiprocessor count of the current device           */
unsigned long long of_launch_count(void); /* kernels launched by this library so far (process-wide) */
/* ABI guards: sizeof(of_gemm_args) / sizeof(of_octree_levels) as compiled into the library -- a binding whose struct
 * mirror has a different size must refuse to run (a stale .so would otherwise read garbage pointers) */
int of_abi_sizeof_gemm_args(void);
int of_abi_sizeof_octree_levels(void);
/* diagnostics: per-role clock64 stamps of CTA `block` of the following of_gather_gemm_tc launches are written to
 * buf [8][cap_per_region] (uint64); buf = NULL switches tracing off (tools/trace_tc.py decodes the stamps) */
int of_tc_trace_set(void* buf, int32_t cap_per_region, int32_t block);
/* kernel-variant switches of of_gather_gemm_tc (experiments / tests; a value outside the set keeps the current one):
 * mt in {1, 2}: 128-row tiles per CTA for N <= 128; uni in {0, 1, 2}: 1 = weight tile inside the gather ring's stage, 2 = deeper separate weight ring;
 * cg in {1, 2}: 2 = CTA pairs (tcgen05 cta_group::2) for the 256-wide tiles; layout in {0, 1}: order of the warp
 * roles (scheduling priority follows the warp id).  Defaults: environment OCTFUSION_TC_MT / _UNI / _CG / _LAYOUT, else
 * 2 / 1 / 1 / 0.  Small M (fewer 128-row tiles than SMs) overrides the tile shape: the widest shape that still fills the
 * SMs (gemm_tc.cu, of_gather_gemm_tc). */
int of_tc_config(int32_t mt, int32_t uni, int32_t cg, int32_t layout);
/* how of_gather_gemm_tc fills the gathered operand tiles: 1 = TMA tile::gather4 (one instruction per 4 neighbour rows,
 * tensor maps built per launch; needs 128-byte aligned rows), 0 = 16-byte cp.async copies by 8 producer warps.  Any other
 * value keeps the current mode.  Default: environment OCTFUSION_TC_TMAG, else 0. */
int of_tc_gather_mode(int32_t tma);

/* ------------------------------------------------------------------------------------------
 * Tap-gather GEMM:   out[m, :] = sum_tap  mean_{j in nbr(m, tap)} [ A[j, :] | onehot(type_j) ] . W[tap]
 *                                 (+ bias) (+ row_add[row_add_idx[m]]) (+ resid[m])
 *
 * One operator covers
 *   GraphConv.forward                 models/networks/modules.py:194-220  (7 taps, mean over the
 *                                     1/4/16 finer neighbours = scatter_mean, utils/scatter.py:42-66,
 *                                     one-hot node type of the NEIGHBOUR appended :199-202)
 *   Conv1x1 / nn.Linear / Conv1d k=1  modules.py:332-339, 523-525 (taps = 1, identity)
 *   Downsample / Upsample GEMMs       modules.py:392-395, 440-443 (identity, in_rows / out_rows maps)
 *   dense Conv3d 3^3, stride 2, and nearest-upsample+conv  modules.py:63-95, 493-502
 *                                     (27 taps over a Morton-ordered voxel table)
 *   the "+ emb_out[batch_id]" loop    modules.py:757-758   (row_add)
 *   the residual / skip add           modules.py:513, 763  (resid)
 *
 * Neighbour table `tap_tab` [M, taps] int32 (row-major):
 *      v >= 0 : exactly one source row v
 *      v == -1: no neighbour (slot contributes zero; count clamps to 1, scatter.py:60)
 *      v <= -2: several sources: o = -(v+2); tap_extra[o] = count n, tap_extra[o+1..o+n] = rows
 *   tap_tab == NULL: identity (taps must be 1); source row = in_rows ? in_rows[m] : m
 * A is the channel concatenation of up to two sources (a0 | a1) -- the torch.cat of the skip
 * stack (graph_unet_hr.py:266) is never materialised.
 * ------------------------------------------------------------------------------------------ */
typedef struct of_gemm_args {
  const void* a0; int64_t lda0; int32_t c0;
  const void* a1; int64_t lda1; int32_t c1;      /* a1 may be NULL (c1 = 0)                   */
  const int32_t* tap_tab; const int32_t* tap_extra;
  const int32_t* in_rows;                        /* identity mode only, may be NULL           */
  int32_t taps;
  const uint8_t* node_type; int32_t ntype;       /* ntype = 0: no one-hot columns             */
  int32_t a_silu;                                /* fp32 path only: apply SiLU to A on load   */
  /* B: fp32 path  -> canonical fp32 [K, N] row-major, K = taps*(c0+c1+ntype), k = tap*(c0+c1+ntype)+c
   *    (exactly GraphConv.weights; other layouts go through of_repack_weight once)
   *    tcgen05 path -> bf16 tile image produced by of_pack_weight_tc                         */
  const void* w;
  const float* bias;                             /* [N] or NULL                               */
  const float* row_add; int64_t ld_row_add; const int32_t* row_add_idx;
  const void* resid; int64_t ld_resid;           /* dtype = activation dtype                  */
  const int32_t* out_rows;                       /* optional scatter of output rows           */
  void* out; int64_t ldo;
  int32_t out_f32;                               /* 1: write fp32 regardless of dtype         */
  int32_t M, N;
  int32_t dtype;                                 /* activation dtype of a0/a1/resid/out       */
  /* tcgen05 path only: slots with several neighbours read their pre-averaged row.  There tap_tab uses
   * the ORDINAL encoding of of_graph_multi_index: v <= -2 -> row -(v+2) of a_multi (built per input tensor
   * by of_gather_mean_rows); multi_types[ord] = per-type neighbour counts, 8 bits per type -- only read when
   * nt_block is NULL, and only exact while a slot has < 256 neighbours of one type (two adaptive levels: <= 16;
   * deeper octrees such as the VAE's depth 8 must pass nt_block).                                       */
  const void* a_multi; int64_t ld_multi;
  const uint64_t* multi_types;
  /* row counts of a0 / a1 (informational; may be 0)                                                          */
  int32_t rows_a0, rows_a1;
  /* tcgen05 path with ntype > 0 (required there): the node-type K block as a precomputed bf16 [M, 64] tensor
   * (of_graph_type_block, record-encoded table)                                                              */
  const void* nt_block;
  /* tcgen05 path: 1 = walk the row tiles from the last to the first.  Alternating the direction from one kernel to
   * the next lets each kernel start on the rows its producer wrote last, which are still in the 126 MB L2.  */
  int32_t reverse;
  /* tcgen05 path, optional: group-norm partial statistics of the OUTPUT, computed in the epilogue from the fp32
   * values before they are rounded to bf16 (the statistics pass of the following DualOctreeGroupNorm,
   * modules.py:291-326, then never reads the tensor).  Rows are cut into 32-row chunks; a chunk is split into
   * SEGMENTS at every change of sample id; stat_chunk_seg [ceil(M/32)+1] is the exclusive prefix sum of segments
   * per chunk (ops.StatPlan in the Python layer builds it once per graph depth).  stat_out [n_segments, N/G, 2]
   * fp32 receives (sum, sum of squares) per G-channel granule of every segment, G = 4 when N %% 128 == 0 else 2 (a
   * 64-channel norm has 2 channels per group) -- one plain store per value, no atomics, fixed summation order:
   * bit-reproducible.  Sample of row m: stat_sample[m], or m / stat_rows_per_sample
   * when stat_sample is NULL.  Requires N % 32 == 0 and out_rows == NULL.  NULL = off.                        */
  float* stat_out;
  const int32_t* stat_chunk_seg;
  const int32_t* stat_seg_slot;                  /* [n_segments] row of stat_out that segment s writes: segments are
                                                  * stored SAMPLE-MAJOR so that of_gn_finalize streams contiguous rows  */
  const int32_t* stat_sample;
  int32_t stat_rows_per_sample;
} of_gemm_args;

/* CUDA-core FFMA path: any shape, fp32-exact accumulation order-insensitive to 1e-6. */
int of_gather_gemm_simt(const of_gemm_args* args, void* stream);

/* tcgen05 / TMEM path (bf16 operands, fp32 accumulate). Requires dtype = OF_BF16,
 * (c0 % 64 == 0), (c1 % 64 == 0), N % 16 == 0 and w packed by of_pack_weight_tc. */
int of_gather_gemm_tc(const of_gemm_args* args, void* stream);
/* Split-K variant for launches whose row tiles cannot fill the GPU (the dense 4^3 / 8^3 levels: M = 2048 at B = 32, with
 * K up to 13824): of_tc_splitk_plan returns the number S of K ranges this library would use for `args` (1 = do not
 * split).  With S > 1 the caller provides a workspace of S * M * N floats and calls of_gather_gemm_tc_splitk: pass 1
 * runs the same tcgen05 kernel over (output tile, K range) pairs writing fp32 partial sums, pass 2 adds the ranges in
 * order (bit-reproducible) with bias / row_add / resid, stores `out` and computes the stat_out statistics.  Same
 * arguments and results as of_gather_gemm_tc up to fp32 summation order. */
int of_tc_splitk_plan(const of_gemm_args* args);
int of_gather_gemm_tc_splitk(const of_gemm_args* args, int32_t splits, float* workspace, void* stream);
/* size in bytes of the packed image for K_feat = taps*(c0+c1) feature rows + ntype one-hot rows */
int64_t of_pack_weight_tc_bytes(int32_t taps, int32_t c, int32_t ntype, int32_t N);
/* w_canonical: fp32 [taps*(c+ntype), N]; out: packed bf16 image */
int of_pack_weight_tc(const float* w_canonical, int32_t taps, int32_t c, int32_t ntype, int32_t N,
                      void* out, void* stream);

/* dst[(tap*c + ci)*N + n] = src[tap*s_tap + ci*s_c + n*s_n]   (element strides).
 * nn.Linear [N,K]: taps=1,s_c=1,s_n=K.  Conv3d [N,C,27]: s_tap=1,s_c=27,s_n=27*C.
 * Downsample [C,C,8] (modules.py:393): taps=1, c=8C, s_c=1, s_n=8C.                        */
int of_repack_weight(const float* src, int64_t s_tap, int64_t s_c, int64_t s_n,
                     int32_t taps, int32_t c, int32_t N, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Group normalisation over ragged per-sample node sets
 *   DualOctreeGroupNorm.forward   models/networks/modules.py:291-326  (count_eps = 1e-5: eps is
 *                                 added to the element count :302 as well as to the variance :310)
 *   GroupNorm32 (dense)           modules.py:26-28                    (count_eps = 0)
 * followed (fused) by SiLU (modules.py:743, 760, graph_unet_hr.py:272) and the channel concat.
 * x is the virtual concatenation (x0 | x1).  sample_id [rows] int32 (NULL => row / rows_per_sample).
 * Statistics are deterministic (no atomics): rows are cut into 32-row chunks, chunks into per-sample SEGMENTS
 * (see of_gemm_args.stat_out), and every segment owns one slot of a partial buffer.
 *   stats:    part [n_segments, C/gran, 2] fp32 = (sum x, sum x^2) per granule of `gran` (2 or 4) channels of each segment, one thread
 *             per (chunk, channel vector), rows added in order.  The tcgen05 GEMM writes the same buffer from its
 *             epilogue (stat_out), in which case this pass is skipped.
 *             Slot of segment s in `part`: seg_slot[s] -- the segments of a sample occupy CONSECUTIVE slots (sample-major,
 *             row order inside the sample).
 *   finalize: for sample b the slots sample_seg_off[b] .. sample_seg_off[b+1] are summed in that order in fp64
 *             (contiguous, coalesced reads) -> mean / variance per group -> scale/shift [B, C] fp32 (gamma, beta folded in).
 *             The normalised tensor is the concat (x0 | x1): part0 / part1 are the partial buffers of the two
 *             tensors with their granule widths (c1 = 0: one tensor).  C/groups and c0 must be multiples of the granules.
 *   apply:    y[r, c] = act(x[r, c] * scale[b, c] + shift[b, c])   act: 0 none, 1 SiLU, 2 GELU (erf)
 * ------------------------------------------------------------------------------------------ */
int of_gn_stats(const void* x0, int64_t ld0, int32_t c0, const void* x1, int64_t ld1, int32_t c1,
                const int32_t* chunk_seg, const int32_t* seg_slot, const int32_t* sample_id, int32_t rows_per_sample,
                int64_t rows, int32_t dtype, int32_t gran, float* part, void* stream);
#define OF_GN_FINALIZE_SPLIT 8   /* CTAs per sample; scratch: batch * OF_GN_FINALIZE_SPLIT * C doubles; ticket: batch
                                  * int32, zero on first use (the kernel leaves them zero); both may be NULL (one CTA) */
int of_gn_finalize(const float* part0, int32_t c0, int32_t gran0, const float* part1, int32_t c1, int32_t gran1,
                   const int32_t* sample_seg_off, int32_t n_segments,
                   const int32_t* rows_of_sample, int32_t rows_per_sample,
                   const float* gamma, const float* beta, int32_t batch, int32_t groups, float eps,
                   float count_eps, float* scale, float* shift, double* scratch, int32_t* ticket, void* stream);
int of_gn_apply(const void* x0, int64_t ld0, int32_t c0, const void* x1, int64_t ld1, int32_t c1,
                const int32_t* sample_id, int32_t rows_per_sample, int64_t rows,
                const float* scale, const float* shift, int32_t act, int32_t dtype,
                void* y, int64_t ldy, int32_t reverse, void* stream);

/* ------------------------------------------------------------------------------------------
 * QKVAttention.forward   models/networks/modules.py:538-547
 * qkv [B*T, 3*C] channels-last, head-major legacy split: head h owns columns
 * [h*3*ch, (h+1)*3*ch) = q | k | v, ch = C / heads; q and k are both scaled by ch^-1/4; softmax
 * over keys in fp32; out [B*T, C] with column h*ch + c.
 * ------------------------------------------------------------------------------------------ */
int of_attention(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int32_t batch,
                 int32_t tokens, int32_t heads, int32_t ch, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small per-step pieces
 * ------------------------------------------------------------------------------------------ */
/* out[b,:] = act(x[b,:]) . W^T + bias for a few rows (B <= a few dozen): the nn.Linear layers of the
 * timestep-embedding path (graph_unet_hr.py:107-111, modules.py:709-715 emb_layers, :479-482 time_mlp).
 * w_nk is the nn.Linear weight in its native [N, K] layout; a_silu applies SiLU to x on load. */
int of_linear_small(const float* x, int64_t ldx, const float* w_nk, const float* bias, int32_t B, int32_t K,
                    int32_t N, int32_t a_silu, float* out, int64_t ldo, void* stream);
/* timestep_embedding  diffusion_networks/ldm_diffusion_util.py:171-191  ->  out [B, dim] fp32 */
int of_timestep_embedding(const float* t, int32_t batch, int32_t dim, float max_period, float* out,
                          void* stream);
/* LearnedSinusoidalPosEmb.forward  modules.py:558-563  -> out [B, 2*half+1] fp32 */
int of_learned_sinusoidal(const float* t, const float* w, int32_t batch, int32_t half, float* out,
                          void* stream);
/* out[b,:] += table[label[b],:]   (nn.Embedding add, graph_unet_hr.py:232-234) */
int of_embedding_add(const float* table, const int32_t* label, int32_t batch, int32_t dim, float* out,
                     void* stream);
/* eps-DDIM update of sample_loop, models/octfusion_model_union.py:345-350.
 * log_snr / log_snr_next are device scalars (so the step is graph-capturable).
 * x (fp32, [n]) is updated in place; x_act (activation dtype copy fed to the first conv) is
 * refreshed when not NULL. */
int of_ddim_eps_update(float* x, const float* eps, const float* log_snr, const float* log_snr_next,
                       int64_t n, void* x_act, int32_t act_dtype, void* stream);
/* "x0" branch of sample_loop (stage 1), models/octfusion_model_union.py:324-344: optional sign() of the
 * prediction (truncation, :324-325), ancestral mean + sqrt(variance) * noise.  noise == NULL: no noise term. */
int of_ddpm_x0_update(float* x, float* pred, const float* noise, const float* log_snr, const float* log_snr_next,
                      int64_t n, int32_t do_sign, void* stream);
/* dtype conversion / strided row copy: dst[r, 0:c] = src[src_rows ? src_rows[r] : r, 0:c]
 * written to row (dst_rows ? dst_rows[r] : r) */
int of_copy_rows(const void* src, int64_t lds, int32_t src_dtype, const int32_t* src_rows,
                 void* dst, int64_t ldd, int32_t dst_dtype, const int32_t* dst_rows,
                 int64_t rows, int32_t c, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dual-octree graph build   models/networks/dualoctree_networks/dual_octree.py:19-63,119-239,
 * 241-271,332-341,381-409  (DualOctree.__init__ + post_processing_for_docnn), for ONE graph depth.
 *
 * Inputs: the octree levels full_depth..depth: keys[d] (int64 Morton | batch<<48, sorted),
 * children[d] (int32, -1 = leaf else rank among non-empty), node counts nnum[d].
 * The graph at depth D has rows  [leaves of full_depth .. leaves of D-1, all nodes of D]
 * (remap_node_idx, dual_octree.py:265-271).
 *
 * of_leaf_rank / of_compact_idx   per level: rank of every leaf among the leaves of its depth, and the
 *                 index lists of leaf / non-empty nodes (row maps of GraphDownsample / GraphUpsample)
 * of_graph_count  pass 1: need[row*7+dir] = words of tap_extra the slot needs (0 for <= 1 neighbour)
 * of_exclusive_scan_i32 over `need`
 * of_graph_fill   pass 2: final tap_tab [rows, 7] (dir 6 = self loop, dual_octree.py:241-249) and
 *                 tap_extra; also node_type [rows] uint8 (:381-389) and batch_id [rows] int32 (:65-79)
 * ------------------------------------------------------------------------------------------ */
typedef struct of_octree_levels {
  const int64_t* keys[16];      /* indexed by depth; only full_depth..depth are read          */
  const int32_t* children[16];
  const int32_t* leaf_rank[16]; /* exclusive scan of (children < 0), from of_leaf_rank         */
  int32_t nnum[16];
  int32_t full_depth, depth, batch;
} of_octree_levels;

/* out[i] = number of j < i with in-flag set (children[j] < 0); *total_out (device) = #leaves.
 * scratch: >= of_scan_scratch_bytes(n) bytes. */
int64_t of_scan_scratch_bytes(int64_t n);
int of_leaf_rank(const int32_t* children, int32_t n, int32_t* rank_out, int32_t* total_out,
                 void* scratch, void* stream);
int of_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out,
                          void* scratch, void* stream);
/* leaf_idx[leaf_rank[i]] = i for leaves, nonempty_idx[children[i]] = i for the others */
int of_compact_idx(const int32_t* children, const int32_t* leaf_rank, int32_t n,
                   int32_t* leaf_idx, int32_t* nonempty_idx, void* stream);
/* number of rows of the depth-D graph (host arithmetic on nnum; no device work) */
int64_t of_graph_rows(const of_octree_levels* oct, int32_t D);
/* pass 1: need[row*7 + dir] = words of tap_extra the slot needs (0 when it has <= 1 neighbour) */
int of_graph_count(const of_octree_levels* oct, int32_t D, int32_t* need, void* stream);
/* pass 2: need_off = exclusive scan of need */
int of_graph_fill(const of_octree_levels* oct, int32_t D, const int32_t* need_off,
                  int32_t* tap_tab, int32_t* tap_extra, uint8_t* node_type, int32_t* batch_id,
                  void* stream);
/* Multi-neighbour slots (coarse leaf next to a subdivided cell: 4..16 finer neighbours, averaged by
 * scatter_mean, utils/scatter.py:42-66; up to 4^k for k adaptive levels).  of_graph_multi_flags marks them (flags[i] = tap_tab[i] <= -2);
 * after an exclusive scan of the flags, of_graph_multi_index writes the ordinal-encoded table used by the
 * tcgen05 path (v <= -2 -> -(ordinal+2)), multi_off[ord] = offset of the slot's record in tap_extra, and
 * multi_types[ord] = packed per-type neighbour counts (8 bits per node type).
 * of_gather_mean_rows: out[ord, :] = mean over the slot's neighbours of (a0|a1)[row, :]  (per input tensor). */
int of_graph_multi_flags(const int32_t* tap_tab, int64_t slots, int32_t* flags, void* stream);
int of_graph_multi_index(const int32_t* tap_tab, const int32_t* tap_extra, const uint8_t* node_type,
                         int64_t slots, const int32_t* flag_scan, int32_t* tap_tab_ord, int32_t* multi_off,
                         uint64_t* multi_types, void* stream);
/* Node-type K block of the tcgen05 GEMM, a per-graph constant: out [rows, 64] bf16, column tap*ntype + type =
 * (#neighbours of that type in slot (row, tap)) / (#neighbours) = the scatter_mean of the one-hot columns that
 * GraphConv.forward appends to the features (models/networks/modules.py:199-202, 208-210); zero elsewhere.
 * tap_tab / tap_extra are the RECORD-encoded tables of of_graph_fill.  Requires taps*ntype <= 64, ntype <= 8. */
int of_graph_type_block(const int32_t* tap_tab, const int32_t* tap_extra, const uint8_t* node_type, int64_t rows,
                        int32_t taps, int32_t ntype, void* out_bf16, void* stream);
int of_gather_mean_rows(const void* a0, int64_t lda0, int32_t c0, const void* a1, int64_t lda1, int32_t c1,
                        const int32_t* tap_extra, const int32_t* multi_off, int32_t count, int32_t dtype,
                        void* out, int64_t ldo, void* stream);
/* hist[v] += 1 for v = values[i] (caller zeroes hist) -- rows per sample for the norm count */
int of_histogram_i32(const int32_t* values, int64_t n, int32_t bins, int32_t* hist, void* stream);
/* reference-format edge list (edge_idx [2,E], edge_dir [E] int64, sorted by row*7+dir:
 * dual_octree.py:332-341) from a tap table.  per_slot[row*taps+tap] = #edges, then edges are
 * written at slot_off (exclusive scan of per_slot). */
int of_graph_edge_count(const int32_t* tap_tab, const int32_t* tap_extra, int64_t slots,
                        int32_t* per_slot, void* stream);
int of_graph_edges(const int32_t* tap_tab, const int32_t* tap_extra, int64_t slots, int32_t taps,
                   const int32_t* slot_off, int64_t* edge_row, int64_t* edge_col, int64_t* edge_dir,
                   void* stream);

/* 3^3 neighbour table of the third-party operator ocnn.nn.OctreeConv (BASELINE.json configs[0]; the reference never
 * calls it -- SURVEY.md section 0 -- its semantics are restated from ocnn-pytorch 2.2.x, SURVEY.md Appendix B: parity
 * is UNPINNED at the ocnn boundary): neigh [nnum[depth], 27] int32, entry (dx+1)*9 + (dy+1)*3 + (dz+1) = index within
 * `depth` of the node at (x+dx, y+dy, z+dz), -1 outside the volume or where no node exists.  The table is a tap table
 * of the tap-gather GEMM (taps = 27, weights [27, Cin, Cout] flattened to [27*Cin, Cout]).  Only keys / children / nnum
 * / full_depth / depth / batch of `oct` are read. */
int of_octree_neigh27(const of_octree_levels* oct, int32_t depth, int32_t* neigh, void* stream);

/* Dense voxel neighbour tables in Morton order for the LR middle U-Net (graph_unet_lr.py):
 * mode 0: 3^3 conv, same resolution `res_log2`        (Conv3d padding=1, modules.py:493-502)
 * mode 1: 3^3 stride-2 conv, out res = in res / 2      (ConvDownsample, modules.py:80-95)
 * mode 2: nearest x2 upsample then 3^3 conv            (ConvUpsample, modules.py:63-77)
 * tap = (dx+1)*9 + (dy+1)*3 + (dz+1); rows = batch * 8^out_res_log2; table values are rows of
 * the INPUT tensor (batch * 8^in_res_log2 rows), -1 outside the grid. */
int of_dense_tap_table(int32_t mode, int32_t out_res_log2, int32_t batch, int32_t* tap_tab,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * NeuralMPU: the implicit function defined by the GraphVAE decoder's regression values
 * (reference models/networks/dualoctree_networks/mpu.py:55-140 `octree_linear_pts` + `get_linear_pred`,
 *  utils/spmm.py `spmm` / `modulated_spmm`).
 *   pos   [npts, 4] fp32: x, y, z in [-1, 1] and the batch index
 *   reg   [sum_{d=full_depth..depth} nnum[d], 4] fp32: per octree node (gradient xyz, value) -- `reg_voxs[depth]` of
 *         GraphVAE.octree_decoder (graph_vae.py:214-221), padded over all nodes
 *   fval  [npts] fp32 = sum_w (F . [offset, 1]) / (sum_w + 1e-8) over the existing cells around the point at depths
 *         full_depth..depth (leaves only below `depth`), w = prod(1 - |offset|) * d^2 / 50
 *   touched [npts] uint8 = 1 when a depth-`depth` cell surrounds the point (the `flgs` mask, mpu.py:139)
 * Only `children`, `nnum`, `full_depth`, `depth`, `batch` of `oct` are read.
 * ------------------------------------------------------------------------------------------ */
int of_mpu_eval(const of_octree_levels* oct, int32_t depth, const float* pos, int64_t npts, const float* reg,
                float* fval, uint8_t* touched, void* stream);
/* The same evaluation on the regular sampling grid of `calc_sdf` (reference utils/util_dualoctree.py:99-118 with
 * get_mgrid :23-42; the 256^3 grid marching cubes consumes): point p of the size^3 grid of shape `batch_idx` is
 * (p / size^2, (p / size) % size, p % size) * ((bbmax - bbmin) / size) + bbmin, generated inside the kernel (no
 * coordinate tensor).  Points [head, head + count) are written to fval[head ..]; fval is the [size^3] array. */
int of_mpu_eval_grid(const of_octree_levels* oct, int32_t depth, int32_t batch_idx, int32_t size, float bbmin,
                     float bbmax, int64_t head, int64_t count, const float* reg, float* fval, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCTFUSION_B200_H_ */

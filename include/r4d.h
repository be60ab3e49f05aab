/*
 * r4d.h — C ABI of the B200-native query-by-pool scoring + top-K engine (libr4d.so).
 *
 * The reference (YuxiaWu/RAG4DyG) has no FFI; its boundary for this path is a set of Python
 * functions.  Each export below names the reference interface it replaces (file:line relative to
 * the reference root).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - every pointer marked [dev] is a device pointer owned by the caller (a torch allocation);
 *    [host] pointers are host memory.  The library never allocates device memory: scratch comes
 *    from a caller-provided workspace sized by the matching *_workspace_bytes() query.
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t); it never synchronises.
 *  - return value: 0 = OK, negative = error (R4D_E_*); r4d_last_error() gives a thread-local text.
 *  - indices are int32 (pool sizes < 2^31).  R4D_IDX_NONE marks "no candidate" (pool shorter than k).
 *  - canonical order everywhere: (score descending, pool index ascending) == np.argsort(-s, kind='stable').
 *  - sm_100a only.  There is no CPU fallback: with no usable device every call returns R4D_E_CUDA.
 */
#ifndef R4D_H_
#define R4D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* r4d_stream_t; /* cudaStream_t */

#define R4D_OK 0
#define R4D_E_ARG (-1)     /* bad argument (null pointer, k out of range, misaligned pitch ...) */
#define R4D_E_CUDA (-2)    /* CUDA runtime / driver error, or no sm_100 device */
#define R4D_E_WORKSPACE (-3) /* workspace too small */
#define R4D_IDX_NONE 0x7fffffff
#define R4D_TOPK_MAX 32    /* fused top-K width limit (one list entry per lane) */
#define R4D_MAX_PEERS 16   /* GPUs of one NVLink domain addressed by the fused exchange */

/* dense scorer epilogue modes (r4d_dense_*) */
#define R4D_DENSE_HALF_COS 0       /* (cos+1)/2                  train/train_retriever.py:437-438 */
#define R4D_DENSE_COS_DECAY 1      /* cos * exp(-lambda*|dt|)    train/train_retriever.py:47-55   */
#define R4D_DENSE_HALF_COS_DECAY 2 /* ((cos+1)/2) * exp(-lambda*|dt|)                              */
/* dense contraction precision */
#define R4D_PREC_BF16 0   /* one tcgen05 kind::f16 pass on bf16-rounded operands                  */
#define R4D_PREC_BF16X3 1 /* hi/lo bf16 split of both operands, three products (hi.hi + hi.lo + lo.hi), fp32
                             accumulate: <= 1e-5 of the reference's fp32 scores                                */

int r4d_version(void);
const char* r4d_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is visible, else 0 (never touches the device otherwise). */
int r4d_device_ok(void);

/* Tuning / measurement knobs (process-wide; defaults in brackets).  Returns the previous value, or R4D_E_ARG.
 *   "jaccard_skip_zero" [1]  skip 8-word spans that are all-zero across a warp (exact; 0 = execute every word-op)
 *   "jaccard_sparse_q"  [1]  fused top-K: sparse query tiles are served by the query-index kernel, which streams the
 *                            pool once per 8 192-query batch (exact; 0 = bitset-streaming kernel for every tile)
 *   "jaccard_debug"     [0]  query-index kernel: stage bypass for measurements (results are WRONG unless 0)
 *   "jaccard_warps"     [16] consumer warps per CTA (8 or 16)
 *   "dense_pair_kernel" [1]  use the CTA-pair (cta_group::2) kernel for bf16 top-K when it applies
 *   "dense_pair_qres"   [-1] query tile resident in smem: -1 auto (when >= 4 pool stages fit), 0 never
 *   "jaccard_stripes"   [0]  0 = automatic; > 0 forces the number of pool stripes of the fused top-K (experiments)
 *   "dense_stripes"     [0]  same for the dense CTA-pair kernel
 *   "kernel_timing"     [0]  record CUDA events around the dominant kernels (see r4d_profile_read)
 *   "stripe_interleave" [0]  dense pair kernel: 1 = stripe s owns pool tiles s, s+S, s+2S, ... (measured: same
 *                            time, 1.7x the DRAM reads), 0 = contiguous stripes
 *   "postings_log_t"    [0]  postings path: 0 = automatic, 9 / 10 = force 512- / 1 024-slot per-warp hash tables
 *   "postings_kernel"   [0]  postings path, label-like sets: first stage 0 = head kernel (k <= 16), 1 = hash-table kernel,
 *                            2 = register-resident kernel (comparison points; results are identical)
 *   "postings_relay"    [1]  packed lists bound for pinned host memory leave in whole 64-query blocks (0 = per chunk)
 *   "postings_best"     [1]  postings path: use the per-id best lists (0 = every query is joined in full; k > 16 or 0
 *                            start the chain at the register-resident kernel)
 *   "postings_chunk"    [0]  postings path: queries a warp takes per grab of the work counter (0 = automatic, <= 8)
 *   "dense_x3_combined" [1]  dense pair kernel, split precision: hi + lo planes of a k-block share one pipeline stage
 *   "dense_walker_window" [4] dense pair kernel: tiles a walker may lead the slowest walker of its stripe (0 = off) */
int r4d_set_option(const char* key, int value);

/* Measurement aid for the roofline figures (bench.py): after r4d_set_option("kernel_timing", 1) the library brackets
 * every launch of its dominant kernels with CUDA events on the launch stream; this call waits for the recorded
 * events, returns their summed duration and count, and clears them.  kernel: "jaccard_qindex" | "jaccard_postings" | "dense_pair". */
int r4d_profile_read(const char* kernel, double* total_ms, int64_t* launches);
/* Kernels this library has enqueued since it was loaded (every <<<>>> of its own; memsets and copies not counted). */
int64_t r4d_kernel_launches(void);

/* ---------------------------------------------------------------- set encoder (subsystem 1)
 * Replaces the per-pair `set(seq_i)`, `set(seq_j)` construction of co_occurrence_ratio
 * (retrieval_data_annotation.py:12-13): each row's token set becomes a fixed-width bitset. */

/* words per row for n_bits (ceil(n_bits/32)) and the row pitch (words, multiple of 32 => 128 B rows). */
int32_t r4d_bitset_words(int32_t n_bits);
int32_t r4d_bitset_pitch_words(int32_t n_bits);

/* CSR -> bitsets.  bit_pos[row_off[r] .. row_off[r+1]) are the bit positions (duplicates allowed,
 * each < n_bits) of row r.  Writes bits[n_rows][pitch_words] (fully, including zero padding) and
 * card[n_rows] = |set| (popcount).  All pointers [dev]. */
int r4d_bitset_encode(const int32_t* bit_pos, const int64_t* row_off, int64_t n_rows, int32_t n_bits,
                      int32_t pitch_words, uint32_t* bits, uint32_t* card, r4d_stream_t stream);

/* ---------------------------------------------------------------- Jaccard scorer (subsystem 2)
 * Replaces occurrence_matrix / co_occurrence_ratio (retrieval_data_annotation.py:36-41, :5-15). */

/* Full matrix: inter[q][p] = |Q_q & P_p| (uint32, row stride ld_inter elements) and, if score != NULL,
 * score[q][p] = inter/union as float64 (0.0 when either set is empty; row stride ld_score).
 * zero_diag != 0 forces the entry with query_base+q == pool_base+p to 0
 * (np.fill_diagonal, retrieval_data_annotation.py:172-173).  All pointers [dev]. */
int r4d_jaccard_full(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                     const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t zero_diag,
                     int64_t query_base, int64_t pool_base, uint32_t* inter, int64_t ld_inter, double* score,
                     int64_t ld_score, r4d_stream_t stream);

/* Fused scorer + top-K (subsystem 4): never materialises [nq, np].
 * Replaces occurrence_matrix + np.argsort(-row)[:k] (retrieval_data_annotation.py:97-103).
 * Outputs [nq][k]: exact integer counts (score = inter/union) and GLOBAL pool index pool_base+p,
 * ordered (score desc, index asc).  1 <= k <= R4D_TOPK_MAX.  Rows short of k are padded with
 * (0, 1, R4D_IDX_NONE).  Bitset rows must be zero padded up to pitch_words (r4d_bitset_encode does that); the
 * workspace must be 16-byte aligned.  Calls with more than 8 192 query rows are served as consecutive 8 192-row
 * launch sequences on the same workspace (one pass over the pool each). */
size_t r4d_jaccard_topk_workspace_bytes(int64_t nq, int64_t np, int32_t k);
int r4d_jaccard_topk(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                     const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t k,
                     int32_t zero_diag, int64_t query_base, int64_t pool_base, uint32_t* top_inter,
                     uint32_t* top_union, int32_t* top_idx, void* workspace, size_t workspace_bytes,
                     r4d_stream_t stream);

/* Merge n_lists candidate lists per query (layout [n_lists][nq][k_in], e.g. the all-gather of the
 * per-shard outputs of r4d_jaccard_topk) into [nq][k_out]; exact rational compare, index tiebreak, so
 * the result does not depend on how the pool was sharded. */
int r4d_jaccard_topk_merge(const uint32_t* inter, const uint32_t* uni, const int32_t* idx, int32_t n_lists,
                           int64_t nq, int32_t k_in, int32_t k_out, uint32_t* out_inter, uint32_t* out_union,
                           int32_t* out_idx, r4d_stream_t stream);

/* Fused exchange (multi-GPU): like r4d_jaccard_topk on this rank's pool shard, but the final per-rank
 * lists are not written locally: the merge kernel stores them straight into slot `rank` of every peer's gather
 * buffer over NVLink.  peer_base [host] holds `world` device pointers (this rank's own buffer included) to buffers of
 * layout uint32 [3][world][nq][k] (planes: inter, union, idx) obtained from a symmetric-memory rendezvous.  After a
 * cross-GPU barrier each rank runs r4d_jaccard_topk_merge over its own buffer's planes (n_lists = world). */
int r4d_jaccard_topk_scatter(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                             const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t k,
                             int32_t zero_diag, int64_t query_base, int64_t pool_base, void* const* peer_base,
                             int32_t world, int32_t rank, void* workspace, size_t workspace_bytes, r4d_stream_t stream);

/* ---------------------------------------------------------------- Jaccard scorer over pool-side postings
 * Same replacement of occurrence_matrix + np.argsort(-row)[:k] (retrieval_data_annotation.py:36-41, :97-103) for SPARSE
 * sets: instead of streaming the pool's bitset rows, an inverted index of the pool (node id -> rows holding it) names
 * exactly the pairs that share an id; every other pair scores 0 and only matters as a lowest-index filler.  Results are
 * bit-identical to r4d_jaccard_topk.  n_bits <= 65 535.
 *
 * r4d_postings_build: pool bitsets (r4d_bitset_encode) -> index blob [dev], 256-byte aligned, of
 * r4d_postings_index_bytes(np, n_bits, nnz) bytes (0 = unsupported shape), nnz = total set bits of the pool = sum of
 * pcard.  The first 8 bytes of the blob are {uint32 magic, uint32 status}; status != 0 after the build means nnz was too
 * small (the index is unusable).  Built once per pool (shard); ~2 passes over the bitsets. */
size_t r4d_postings_index_bytes(int64_t np, int32_t n_bits, int64_t nnz);
size_t r4d_postings_build_workspace_bytes(int64_t np, int32_t n_bits);
int r4d_postings_build(const uint32_t* pbits, const uint32_t* pcard, int64_t np, int32_t n_bits, int32_t pitch_words,
                       int64_t nnz, void* index, size_t index_bytes, void* workspace, size_t workspace_bytes,
                       r4d_stream_t stream);

/* Fused scorer + top-K over the postings.  Queries arrive as CSR id lists q_ids[q_off[q] .. q_off[q+1]) [dev] (ids outside
 * [0, n_bits) are ignored, duplicates inside a row collapse: Python set() semantics, retrieval_data_annotation.py:12-13);
 * no query bitsets are needed.  q_off may be a row range of a larger CSR (its values index q_ids absolutely);
 * q_nnz = q_off[nq] - q_off[0], the number of ids of the nq rows (known to the host), only sizes the hash tables.
 * index / pcard / np / n_bits / nnz as given to r4d_postings_build.  Outputs as
 * r4d_jaccard_topk: [nq][k] exact counts + GLOBAL pool index pool_base+p, order (score desc, index asc), rows short of k
 * padded with (0, 1, R4D_IDX_NONE).  Any set size and skew is served exactly (queries too dense for the per-warp hash
 * tables are completed by a per-window counting kernel). */
size_t r4d_jaccard_topk_postings_workspace_bytes(int64_t nq);
int r4d_jaccard_topk_postings(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                              const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k, int32_t zero_diag,
                              int64_t query_base, int64_t pool_base, uint32_t* top_inter, uint32_t* top_union,
                              int32_t* top_idx, void* workspace, size_t workspace_bytes, r4d_stream_t stream);
/* Packed results — for [nq][k] lists that cross PCIe (HostTopK stores them straight into pinned host memory): 8 bytes per
 * entry instead of 12.  top_pair[q][j] = inter << 16 | |pool set| (both <= 65 535 because n_bits is), top_idx as above,
 * q_card[q] = |query set| (distinct valid ids of the row).  The consumer recovers union = q_card + |pool set| - inter, so
 * the score inter / union (retrieval_data_annotation.py:12-14) is the same rational; a padding entry (idx R4D_IDX_NONE)
 * packs as 0 and stands for (0, 1).  Same kernels, same order, same exactness as r4d_jaccard_topk_postings.
 * When top_pair / top_idx / q_card are pinned HOST buffers (device-addressable, 128-byte aligned) and the workspace is
 * r4d_jaccard_topk_postings_workspace_bytes(nq) + r4d_jaccard_topk_postings_relay_bytes(nq, k) bytes or more, the lists
 * are staged in that extra room and leave for the host in whole 64-query blocks of full 128-byte PCIe writes (writes
 * from an SM are bound by their number, not their bytes); with the smaller workspace the kernels store chunk by chunk. */
size_t r4d_jaccard_topk_postings_relay_bytes(int64_t nq, int32_t k);
int r4d_jaccard_topk_postings_packed(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                                     const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k,
                                     int32_t zero_diag, int64_t query_base, int64_t pool_base, uint32_t* top_pair,
                                     int32_t* top_idx, uint32_t* q_card, void* workspace, size_t workspace_bytes,
                                     r4d_stream_t stream);
/* Fused exchange variant (see r4d_jaccard_topk_scatter): the final lists go to slot `rank` of every peer's gather buffer. */
int r4d_jaccard_topk_postings_scatter(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz,
                                      const void* index,
                                      const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k,
                                      int32_t zero_diag, int64_t query_base, int64_t pool_base, void* const* peer_base,
                                      int32_t world, int32_t rank, void* workspace, size_t workspace_bytes,
                                      r4d_stream_t stream);

/* ---------------------------------------------------------------- ranking of score matrices
 * Full descending STABLE ranking of every row: order[q][:] = np.argsort(-scores[q], kind='stable').
 * Replaces np.argsort(-M, axis=1) in save_index_score (retrieval_data_annotation.py:89,
 * train/train_retriever.py:358).  NaN sorts last.  workspace from the *_workspace_bytes query. */
size_t r4d_rank_rows_workspace_bytes(int64_t nq, int64_t n, int32_t elem_bytes);
int r4d_rank_rows_f64(const double* scores, int64_t nq, int64_t n, int64_t ld, int32_t* order, void* workspace,
                      size_t workspace_bytes, r4d_stream_t stream);
int r4d_rank_rows_f32(const float* scores, int64_t nq, int64_t n, int64_t ld, int32_t* order, void* workspace,
                      size_t workspace_bytes, r4d_stream_t stream);

/* Top-k of every row of an explicit float64 matrix, canonical order
 * (save_score_file_train on a caller-provided matrix, retrieval_data_annotation.py:97-103). */
int r4d_topk_rows_f64(const double* scores, int64_t nq, int64_t n, int64_t ld, int32_t k, double* top_score,
                      int32_t* top_idx, r4d_stream_t stream);

/* ---------------------------------------------------------------- triplet mining
 * Device part of save_train_annotation (retrieval_data_annotation.py:54-71): for every row i of the
 * [n, n] float64 matrices `out` (label-set Jaccard) and `in` (history-set Jaccard):
 *   n_pos[i]          = #{j : out[i,j] > thr}
 *   neg[i][0..n_neg)  = the first neg_num indices j, in descending-in[i,j] (stable) order, with
 *                       out[i,j] <= thr and out[i,j] > 0; if fewer than neg_num, continued in the same
 *                       order with out[i,j] == 0.
 * The host then lists positives (np.where) and replays np.random.choice (:79).  neg_num <= 32. */
int r4d_triplet_mine_f64(const double* out, const double* in, int64_t n, int64_t ld, double thr,
                         int32_t neg_num, int32_t* n_pos, int32_t* neg /*[n][neg_num]*/, int32_t* n_neg,
                         r4d_stream_t stream);

/* The same mining straight from the OUT (label) and IN (history) BITSETS of the train pool — no [n, n] matrix in HBM
 * (replaces occurrence_matrix x 2 + the loop of retrieval_data_annotation.py:53-71 in one pass).
 * zero_diag != 0 forces both scores of the pairs (i, i) to 0, as :172-173 does before the mining.
 *   positives: every (i, j) with out[i,j] > thr is APPENDED, in no particular order, as
 *              pos_key[t] = i << 32 | j and pos_counts[t] = |A_i n A_j| << 32 | |A_i u A_j| (OUT sets), t < pos_cap;
 *              *pos_total [dev] receives the number of positives found (if > pos_cap: call again with larger buffers);
 *              n_pos[i] their number per row.  Sorting the keys gives np.where's row-major order (:54).
 *   negatives: neg / n_neg as r4d_triplet_mine_f64, plus the OUT counts of every negative (neg_inter, neg_union
 *              [n][neg_num]; 0 / 1 for unused slots), so that out[i, neg] = neg_inter / neg_union needs no matrix.
 * Scores compare as exact rationals; out[i,j] > thr is evaluated on double(inter) / double(union) like the reference.
 * Bitset rows as produced by r4d_bitset_encode (16-byte aligned, pitch a multiple of 4 words, zero padded). */
int r4d_triplet_mine(const uint32_t* obits, const uint32_t* ocard, int32_t owords, int32_t opitch, const uint32_t* ibits,
                     const uint32_t* icard, int32_t iwords, int32_t ipitch, int64_t n, double thr, int32_t neg_num,
                     int32_t zero_diag, int32_t* n_pos, int64_t* pos_key, int64_t* pos_counts, int64_t pos_cap,
                     int64_t* pos_total, int32_t* neg, uint32_t* neg_inter, uint32_t* neg_union, int32_t* n_neg,
                     r4d_stream_t stream);

/* HOST function (no GPU work): replays `n_calls` consecutive np.random.choice(a) calls (:79) of numpy's legacy global
 * RandomState on its MT19937 state — key[624] and *pos as returned by np.random.get_state() — for arrays of
 * sizes[t] >= 1 elements: choice[t] = the index numpy would pick; key / *pos are advanced exactly as numpy advances
 * them (a one-element array consumes no randomness; otherwise masked rejection sampling on 32-bit outputs). */
int r4d_mt19937_choice_replay(uint32_t* key, int32_t* pos, const int32_t* sizes, int64_t n_calls, int32_t* choice);

/* Counter-based replacement of the sequential np.random.choice (:79) — NOT bit-compatible with numpy's RNG, offered
 * as the deterministic device-side mode of SURVEY.md 8f-2.  For positive pair t (row pos_row[t], its rank within the
 * row = t - row_start[row]): choice[t] = neg[row][splitmix64(seed, row, rank) % n_neg[row]] (or -1 if n_neg == 0).
 * pos_row [n_pairs] and row_start [n_rows+1] are int64 [dev]. */
int r4d_triplet_sample(const int64_t* pos_row, const int64_t* row_start, int64_t n_pairs, const int32_t* neg,
                       const int32_t* n_neg, int32_t neg_num, uint64_t seed, int32_t* choice, r4d_stream_t stream);

/* ---------------------------------------------------------------- dense scorer (subsystem 3)
 * Replaces the scoring block of test() (train/train_retriever.py:433-438) and, optionally, the
 * exp(-lambda*|dt|) factor of CLtime_loss (train/train_retriever.py:50-55).
 *
 * r4d_dense_prepare: rows of x[n][d] (fp32) are L2-normalised (x / ||x||, no epsilon: a zero row gives
 * NaN exactly like :433,:436) and written as bf16 planes hi[n][d_pad] and (prec == BF16X3) lo[n][d_pad],
 * d_pad = r4d_dense_dpad(d) (multiple of 64, zero padded). */
int32_t r4d_dense_dpad(int32_t d);
int r4d_dense_prepare(const float* x, int64_t n, int32_t d, int64_t ld, int32_t prec, void* hi, void* lo,
                      r4d_stream_t stream);

/* Fused tcgen05 contraction + epilogue + top-K over prepared operands.
 * q_time/p_time may be NULL for mode R4D_DENSE_HALF_COS.  Outputs [nq][k] float32 scores and global
 * pool indices, ordered (score desc, index asc). */
size_t r4d_dense_topk_workspace_bytes(int64_t nq, int64_t np, int32_t k);
int r4d_dense_topk(const void* q_hi, const void* q_lo, int64_t nq, const void* p_hi, const void* p_lo, int64_t np,
                   int32_t d_pad, int32_t prec, const float* q_time, const float* p_time, float lambda,
                   int32_t mode, int32_t k, int64_t pool_base, float* top_score, int32_t* top_idx,
                   void* workspace, size_t workspace_bytes, r4d_stream_t stream);

/* Fused exchange for the dense scorer: peer buffers of layout [2][world][nq][k] (planes: float32 score, int32 idx). */
int r4d_dense_topk_scatter(const void* q_hi, const void* q_lo, int64_t nq, const void* p_hi, const void* p_lo, int64_t np,
                           int32_t d_pad, int32_t prec, const float* q_time, const float* p_time, float lambda,
                           int32_t mode, int32_t k, int64_t pool_base, void* const* peer_base, int32_t world,
                           int32_t rank, void* workspace, size_t workspace_bytes, r4d_stream_t stream);

/* Same contraction + epilogue, full score rows scores[nq][ld] (the `.gen` score files need them,
 * train/train_retriever.py:357-368). */
int r4d_dense_full(const void* q_hi, const void* q_lo, int64_t nq, const void* p_hi, const void* p_lo, int64_t np,
                   int32_t d_pad, int32_t prec, const float* q_time, const float* p_time, float lambda,
                   int32_t mode, float* scores, int64_t ld, r4d_stream_t stream);

/* Pool-embedding producer: hidden[batch][len][d] fp32 -> mean over the padded length
 * (train/train_retriever.py:420, :432) -> L2-normalised bf16 planes exactly as r4d_dense_prepare would give for that
 * mean; mean_out (nullable) receives the fp32 means [batch][d].  Deterministic (no float atomics). */
size_t r4d_meanpool_workspace_bytes(int64_t batch, int32_t d);
int r4d_meanpool_prepare(const float* hidden, int64_t batch, int32_t len, int32_t d, int32_t prec, float* mean_out,
                         void* hi, void* lo, void* workspace, size_t workspace_bytes, r4d_stream_t stream);

/* Merge [n_lists][nq][k_in] dense candidate lists into [nq][k_out] (score desc, index asc). */
int r4d_dense_topk_merge(const float* score, const int32_t* idx, int32_t n_lists, int64_t nq, int32_t k_in,
                         int32_t k_out, float* out_score, int32_t* out_idx, r4d_stream_t stream);

/* ---------------------------------------------------------------- device-side text assembly
 * The same files (retrieval_data_annotation.py:88-103, train/train_retriever.py:357-368) assembled in HBM from device
 * arrays: line q == ' '.join(field(vals[q][j]) for j in range(n)) + '\n', where field(v) = the decimal text of v
 * (lut_off == NULL) or the string lut_blob[lut_off[v] : lut_off[v + 1]] of a caller-built table (one entry per distinct
 * float, formatted by the reference's own formatter).  Two calls: _sizes fills row_off [nq + 1] (int64, dev; row_off[nq]
 * = file size) and *status (dev, nullable: 1 if a code was out of range); the second writes row_off[nq] bytes to `text`
 * (dev).  vals [nq][ld] int32, lut_blob / lut_off device pointers. */
int r4d_format_rows_device_sizes(const int32_t* vals, int64_t nq, int64_t n, int64_t ld, const int64_t* lut_off,
                                 int32_t n_codes, int64_t* row_off, int32_t* status, r4d_stream_t stream);
int r4d_format_rows_device(const int32_t* vals, int64_t nq, int64_t n, int64_t ld, const char* lut_blob, const int64_t* lut_off,
                           int32_t n_codes, const int64_t* row_off, char* text, r4d_stream_t stream);

/* ---------------------------------------------------------------- host-side text formatters
 * [host] pointers.  Produce the exact bytes of ' '.join(str(x) for x in row) + '\n' per row
 * (retrieval_data_annotation.py:92-93,102-103; train/train_retriever.py:362-363).  Return bytes written (>= 0)
 * or a negative error code.  r4d_format_lut_rows looks every element's text up in a caller-built table
 * (lut_off has n_codes + 1 entries into lut_blob), so float formatting stays the reference's own. */
size_t r4d_format_int_rows_bound(int64_t nq, int64_t n);
int64_t r4d_format_int_rows(const int32_t* rows, int64_t nq, int64_t n, int64_t ld, char* out, size_t cap);
int64_t r4d_format_lut_rows(const int32_t* codes, int64_t nq, int64_t n, int64_t ld, const char* lut_blob,
                            const int64_t* lut_off, int32_t n_codes, char* out, size_t cap);

/* ---------------------------------------------------------------- host-side text readers (consumer half)
 * [host] pointers.  The generator's dataloader reads the index / score files back with
 *   [list(map(int, l.split())) for l in f.read().splitlines() if len(l) > 0 and not l.isspace()]   (and float)
 * (dataloader/generator.py:32-48).  These calls parse the whole file text at once, multi-threaded: rows = lines holding at
 * least one field, CSR offsets row_off[n_rows + 1] into values[].  strtoll / strtod semantics (correctly rounded:
 * identical to Python's int() / float() for everything the writers above emit; Python-only spellings such as
 * underscores are rejected with R4D_E_ARG).  Return the number of rows (>= 0) or a negative error code;
 * R4D_E_WORKSPACE when cap_rows / cap_fields (from r4d_parse_rows_count) are too small. */
int64_t r4d_parse_rows_count(const char* text, size_t len, int64_t* n_fields);
int64_t r4d_parse_int_rows(const char* text, size_t len, int64_t* row_off, int64_t* values, int64_t cap_rows,
                           int64_t cap_fields);
int64_t r4d_parse_float_rows(const char* text, size_t len, int64_t* row_off, double* values, int64_t cap_rows,
                             int64_t cap_fields);

#ifdef __cplusplus
}
#endif
#endif /* R4D_H_ */

// jaccard_common.cuh — shared between the dense-bitset Jaccard kernel (jaccard.cu) and the query-index kernel
// (jaccard_sparse.cu).
#pragma once
#include "r4d_common.cuh"

namespace r4d {

constexpr int SQ_TQ = 128;            // query rows per tile (same tiling as jaccard.cu)
constexpr int SQ_TP = 128;            // pool rows per tile
constexpr int SQ_T1 = 1024;           // non-zero words per query tile served by the query-index path (8 per row)
constexpr int SQ_TBITS = 2560;        // set bits per query tile served by the query-index path (20 per row)
constexpr int SQ_E_CAP = 20480;       // index entries (one per set bit) resident in shared memory at once: one "group"
                                      // of consecutive query tiles
constexpr int SQ_BATCH_ROWS = 8;      // pool rows a warp scans before it looks its non-zero words up
constexpr int SQ_RING_BYTES = 7680;   // per-warp ring of bulk-copy slots (whole pool rows, contiguous in HBM)
constexpr int SQ_NBK = 20480;         // buckets of the bit index in shared memory (bit id >> shift; shift 0 up to 20 480 bits)
constexpr int SQ_MAX_SLOTS = 4;
constexpr int SQ_QB = 8192;           // query rows per batch (one launch sequence)
constexpr int SQ_GC = 1024;           // slots of a query's global candidate array (first arrivals; the rest goes to the
                                      // per-stripe lists)
constexpr int SQ_MAX_TILES = SQ_QB / SQ_TQ;
constexpr int SQ_MAX_WORDS = 1900;    // one row (+16 zero bytes) fits a warp's ring; bit ids fit 16 bits
constexpr int SQ_ROWOFF_LD = 132;     // 129 row offsets per tile, padded
constexpr int SQ_WARPS = 16;
constexpr int SQ_THREADS = SQ_WARPS * 32;

// Per-stripe partial candidate lists, query-major [nq][n_stripes][k]: one 16-byte entry {inter, union, pool index, 0} per
// candidate, so a candidate costs one memory sector to store and one to merge.
constexpr int SQ_PART_BYTES = 16;

// By-row index of the non-zero words of one query batch + per-(stripe, query) candidate counts (device memory inside
// the caller's workspace).
struct QIndex {
    uint32_t* tile_cnt;    // [n_qtiles]  non-zero words of the tile (0 when the tile is flagged dense)
    uint32_t* tile_bits;   // [n_qtiles]  set bits of the tile (0 when the tile is flagged dense)
    uint32_t* tile_dense;  // [n_qtiles]  1: more than SQ_T1 words / SQ_TBITS bits -> the dense kernel handles the tile
    uint16_t* rowoff;      // [n_qtiles][SQ_ROWOFF_LD]  exclusive offsets of every row's entries inside the tile
    uint16_t* ent_word;    // [n_qtiles][SQ_T1]  word id
    uint32_t* ent_val;     // [n_qtiles][SQ_T1]  word value
    uint8_t* ent_row;      // [n_qtiles][SQ_T1]  row of the entry inside its tile
    uint32_t* gcount;      // [nq]  candidates offered to the query's global array (may exceed SQ_GC); cleared with `cnt`
    uint4* glist;          // [nq][SQ_GC] {inter, |pool set|, idx, 0}: the first SQ_GC candidates of a query, contiguous
    uint32_t* any_dense;   // [1] != 0 when some tile of the batch is flagged dense
    uint8_t* cnt;          // [nq][n_stripes]    candidates stored in the (stripe, query) partial list (unsorted)
};

bool sparseq_supported(int32_t words, int32_t k);
size_t sparseq_index_bytes(int64_t nq_total);
size_t sparseq_batch_bytes(int64_t nq_batch, int32_t n_stripes);
QIndex sparseq_carve_index(void* base, int64_t nq_total);
QIndex sparseq_batch_view(const QIndex& all, int64_t q0, void* batch_base, int64_t nq_batch, int32_t n_stripes);
// Build the by-row index of every tile of the call with one launch; clear one batch's candidate counts.
int sparseq_build_all(const uint32_t* qbits, int64_t nq_total, int32_t words, int32_t pitch_words, const QIndex& all,
                      cudaStream_t st);
int sparseq_clear_batch(const QIndex& qi, int64_t nq, int32_t n_stripes, cudaStream_t st);
// Stream the pool once per group of query tiles; every non-zero pool word looks up the queries holding that word.
int sparseq_topk_launch(const uint32_t* pbits, const uint32_t* qcard, const uint32_t* pcard, int64_t nq, int64_t np,
                        int32_t words, int32_t pitch_words, int32_t k, int32_t zero_diag, int64_t query_base,
                        int64_t pool_base, int32_t n_qtiles, int32_t n_ptiles, int32_t n_stripes, int32_t ptiles_per_stripe,
                        uint4* part, const QIndex& qi, cudaStream_t st);

}  // namespace r4d

// jaccard_common.cuh — shared between the dense-bitset Jaccard kernel (jaccard.cu) and the sparse-query kernel
// (jaccard_sparse.cu).
#pragma once
#include "r4d_common.cuh"

namespace r4d {

constexpr int SQ_TQ = 128;            // query rows per tile (same tiling as jaccard.cu)
constexpr int SQ_TP = 128;            // pool rows per tile
constexpr int SQ_CHUNK_WORDS = 32;    // words of every row per TMA stage
constexpr int SQ_STAGE_BYTES = SQ_TP * SQ_CHUNK_WORDS * 4;  // 16 KB: pool chunk only
constexpr int SQ_E_MAX = 1024;        // non-zero 8-word spans per query tile served by the sparse path
constexpr int SQ_MAX_CHUNKS = 64;     // sparse path serves W <= 2048 words
constexpr int SQ_OFF_LD = SQ_MAX_CHUNKS + 4;  // chunk offsets per tile (uint16), even and 8-byte rows
constexpr int SQ_WARPS = 16;
constexpr int SQ_THREADS = SQ_WARPS * 32;
constexpr int SQ_TOUCH_CAP = 256;     // remembered non-zero cells per warp per pool tile before a full row scan

// Span lists of every query tile (device memory inside the caller's workspace)
struct SparseQ {
    uint32_t* tile_dense;  // [n_qtiles]  1: more than SQ_E_MAX spans -> the dense kernel handles the tile
    uint16_t* off;         // [n_qtiles][SQ_OFF_LD]  exclusive offsets of the entries of every chunk; [n_chunks] = total
    uint32_t* hdr;         // [n_qtiles][SQ_E_MAX]   row | (span-in-chunk << 8)
    uint32_t* words;       // [n_qtiles][SQ_E_MAX][8] the 8 words of the span
};

size_t sparseq_workspace_bytes(int64_t nq);
bool sparseq_supported(int32_t words, int32_t k);
SparseQ sparseq_carve(void* base, int64_t nq);
int sparseq_build(const uint32_t* qbits, int64_t nq, int32_t words, int32_t pitch_words, const SparseQ& sq, cudaStream_t st);
int sparseq_topk_launch(const uint32_t* pbits, const uint32_t* qcard, const uint32_t* pcard, int64_t nq, int64_t np,
                        int32_t words, int32_t pitch_words, int32_t k, int32_t zero_diag, int64_t query_base,
                        int64_t pool_base, int32_t n_qtiles, int32_t n_ptiles, int32_t n_stripes, int32_t ptiles_per_stripe,
                        uint32_t* part_inter, uint32_t* part_union, int32_t* part_idx, const SparseQ& sq, cudaStream_t st);

}  // namespace r4d

// text_format.cu — host-side text formatters for the reference's file formats.  Pure CPU code in the
// same shared object: after GPU scoring the Python-level ' '.join(str(x) ...) of retrieval_data_annotation.py:92-93 /
// train/train_retriever.py:362-363 dominates wall time (dialog: 175 MB of text).
//   r4d_format_int_rows : rows of int32 -> "a b c\n..." (decimal, single spaces)
//   r4d_format_lut_rows : rows of codes -> strings looked up in a caller-built table (each distinct float is formatted
//                         once in Python with the reference's own formatter, so the bytes are identical by construction)
#include <cstring>

#include "r4d_common.cuh"

extern "C" {

size_t r4d_format_int_rows_bound(int64_t nq, int64_t n) { return (size_t)nq * ((size_t)n * 12 + 1) + 1; }

// returns bytes written, or a negative error code
int64_t r4d_format_int_rows(const int32_t* rows, int64_t nq, int64_t n, int64_t ld, char* out, size_t cap) {
    if (nq < 0 || n < 0 || ld < n || (!rows && nq * n > 0) || !out) return R4D_E_ARG;
    if (cap < r4d_format_int_rows_bound(nq, n)) {
        r4d::set_error("format_int_rows: buffer %zu B < bound %zu B", cap, r4d_format_int_rows_bound(nq, n));
        return R4D_E_WORKSPACE;
    }
    char* p = out;
    char tmp[16];
    for (int64_t q = 0; q < nq; ++q) {
        const int32_t* r = rows + q * ld;
        for (int64_t j = 0; j < n; ++j) {
            int64_t v = r[j];
            if (j) *p++ = ' ';
            if (v < 0) {
                *p++ = '-';
                v = -v;
            }
            int len = 0;
            do {
                tmp[len++] = (char)('0' + v % 10);
                v /= 10;
            } while (v);
            while (len) *p++ = tmp[--len];
        }
        *p++ = '\n';
    }
    return (int64_t)(p - out);
}

// lut_blob: concatenated strings; lut_off[c]..lut_off[c+1] delimits the text of code c (n_codes + 1 offsets)
int64_t r4d_format_lut_rows(const int32_t* codes, int64_t nq, int64_t n, int64_t ld, const char* lut_blob,
                            const int64_t* lut_off, int32_t n_codes, char* out, size_t cap) {
    if (nq < 0 || n < 0 || ld < n || (!codes && nq * n > 0) || !lut_blob || !lut_off || !out || n_codes <= 0)
        return R4D_E_ARG;
    int64_t max_len = 0;
    for (int32_t c = 0; c < n_codes; ++c)
        if (lut_off[c + 1] - lut_off[c] > max_len) max_len = lut_off[c + 1] - lut_off[c];
    if (cap < (size_t)nq * ((size_t)n * (size_t)(max_len + 1) + 1) + 1) {
        r4d::set_error("format_lut_rows: buffer too small");
        return R4D_E_WORKSPACE;
    }
    char* p = out;
    for (int64_t q = 0; q < nq; ++q) {
        const int32_t* r = codes + q * ld;
        for (int64_t j = 0; j < n; ++j) {
            const int32_t c = r[j];
            if (c < 0 || c >= n_codes) {
                r4d::set_error("format_lut_rows: code %d out of range", c);
                return R4D_E_ARG;
            }
            if (j) *p++ = ' ';
            const int64_t len = lut_off[c + 1] - lut_off[c];
            memcpy(p, lut_blob + lut_off[c], (size_t)len);
            p += len;
        }
        *p++ = '\n';
    }
    return (int64_t)(p - out);
}

}  // extern "C"

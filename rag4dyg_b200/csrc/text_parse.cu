// text_parse.cu — host-side readers for the reference's index / score files (consumer half).
// The generator's dataloader parses them back with
//     lines = [l for l in f.read().splitlines() if len(l) > 0 and not l.isspace()]
//     index = [list(map(int, l.split())) for l in lines];  score = [list(map(float, l.split())) for l in lines]
// (dataloader/generator.py:32-48): one Python object per number — 12 M floats for dialog's val scores.  Pure CPU code in
// the same shared object; threads split the file by lines.
//   r4d_parse_rows_count : rows (lines that are neither empty nor all whitespace) and fields (whitespace separated)
//   r4d_parse_int_rows   : CSR row offsets + int64 values   (strtoll, base 10)
//   r4d_parse_float_rows : CSR row offsets + float64 values (strtod: correctly rounded, so == Python's float(text))
// The accepted number syntax is C's ("C" locale), which covers everything the writers of this path emit (decimal
// integers; str(np.float64) / "%.4f" floats incl. exponents, nan, inf); Python-only forms (underscores, non-ASCII
// digits) are rejected with R4D_E_ARG.  Line structure follows Python's text mode + str.splitlines() on ASCII input:
// \n, \r, \r\n, \v, \f, \x1c, \x1d, \x1e end a line; space, \t and \x1f separate fields (str.split()).  Non-ASCII
// bytes (where Python knows further separators: \x85, U+2028, U+2029, Unicode spaces) are rejected with R4D_E_ARG —
// the writers of this path never produce them.
#include <locale.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "r4d_common.cuh"

namespace r4d {

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\x1f'; }
static inline bool is_line_end(char c) {
    return c == '\n' || c == '\r' || c == '\v' || c == '\f' || c == '\x1c' || c == '\x1d' || c == '\x1e';
}

struct Line {
    const char* beg;
    const char* end;  // exclusive, the '\n' is not part of the line
};

// lines that hold at least one field, with their field counts; false when the text holds a non-ASCII byte
static bool split_lines(const char* text, size_t len, std::vector<Line>& lines, std::vector<int64_t>& fields) {
    const char* p = text;
    const char* const stop = text + len;
    while (p < stop) {
        int64_t n = 0;
        bool in_field = false;
        const char* c = p;
        for (; c < stop && !is_line_end(*c); ++c) {
            if (static_cast<unsigned char>(*c) >= 0x80) return false;
            const bool sp = is_space(*c);
            if (!sp && !in_field) ++n;
            in_field = !sp;
        }
        if (n > 0) {
            lines.push_back(Line{p, c});
            fields.push_back(n);
        }
        p = c < stop ? c + 1 : stop;
    }
    return true;
}

static locale_t c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

template <class T>
static bool parse_field(const char* b, const char* e, T* out);

template <>
bool parse_field<int64_t>(const char* b, const char* e, int64_t* out) {
    {   // plain decimals of up to 18 digits need no libc call
        const char* c = b;
        bool neg = false;
        if (c < e && (*c == '-' || *c == '+')) neg = *c++ == '-';
        if (c < e && e - c <= 18) {
            int64_t v = 0;
            for (; c < e && *c >= '0' && *c <= '9'; ++c) v = v * 10 + (*c - '0');
            if (c == e) {
                *out = neg ? -v : v;
                return true;
            }
        }
    }
    char tmp[48];
    const size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof(tmp)) return false;
    memcpy(tmp, b, n);
    tmp[n] = 0;
    char* endp = nullptr;
    errno = 0;
    const long long v = strtoll(tmp, &endp, 10);
    if (endp != tmp + n || errno == ERANGE) return false;
    *out = (int64_t)v;
    return true;
}

// Clinger's fast path: a decimal with at most 2^53 as its digit string and a power of ten up to 10^22 is the correctly
// rounded result of ONE IEEE multiplication / division of two exactly representable doubles.  Covers every
// str(np.float64) / "%.4f" text of this path ("0.0", "0.06666666666666667", "1e-05", ...); anything else -> strtod.
static bool fast_double(const char* b, const char* e, double* out) {
    static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* c = b;
    bool neg = false;
    if (c < e && (*c == '-' || *c == '+')) neg = *c++ == '-';
    uint64_t m = 0;
    int digits = 0, frac = 0;
    bool seen_point = false, any = false;
    for (; c < e; ++c) {
        if (*c >= '0' && *c <= '9') {
            any = true;
            if (m > (uint64_t)900719925474099ull) return false;   // the next digit could pass 2^53
            m = m * 10 + (uint64_t)(*c - '0');
            if (seen_point) ++frac;
            ++digits;
        } else if (*c == '.' && !seen_point) {
            seen_point = true;
        } else {
            break;
        }
    }
    if (!any || m > (uint64_t)9007199254740992ull) return false;
    int ex = 0;
    if (c < e && (*c == 'e' || *c == 'E')) {
        ++c;
        bool eneg = false;
        if (c < e && (*c == '-' || *c == '+')) eneg = *c++ == '-';
        if (c >= e) return false;
        int v = 0;
        for (; c < e; ++c) {
            if (*c < '0' || *c > '9' || v > 1000) return false;
            v = v * 10 + (*c - '0');
        }
        ex = eneg ? -v : v;
    }
    if (c != e) return false;
    ex -= frac;
    if (ex < -22 || ex > 22) return false;
    double v = (double)m;
    v = ex < 0 ? v / P10[-ex] : v * P10[ex];
    *out = neg ? -v : v;
    return true;
}

template <>
bool parse_field<double>(const char* b, const char* e, double* out) {
    if (fast_double(b, e, out)) return true;
    char tmp[96];
    const size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof(tmp)) return false;
    memcpy(tmp, b, n);
    tmp[n] = 0;
    if (memchr(tmp, 'x', n) || memchr(tmp, 'X', n) || memchr(tmp, 'p', n) || memchr(tmp, 'P', n) || memchr(tmp, '(', n))
        return false;  // C-only forms (hex floats, nan(...)) that Python's float() rejects
    char* endp = nullptr;
    const double v = strtod_l(tmp, &endp, c_locale());   // Python's float() ignores LC_NUMERIC; so does this
    if (endp != tmp + n) return false;
    *out = v;
    return true;
}

template <class T>
static int64_t parse_rows(const char* text, size_t len, int64_t* row_off, T* values, int64_t cap_rows, int64_t cap_fields) {
    if ((!text && len) || !row_off || (!values && cap_fields > 0)) return R4D_E_ARG;
    std::vector<Line> lines;
    std::vector<int64_t> fields;
    if (!split_lines(text, len, lines, fields)) {
        set_error("parse_rows: non-ASCII byte in the text (only ASCII index / score files are accepted)");
        return R4D_E_ARG;
    }
    const int64_t n_rows = (int64_t)lines.size();
    int64_t total = 0;
    for (int64_t n : fields) total += n;
    if (n_rows > cap_rows || total > cap_fields) {
        set_error("parse_rows: %lld rows / %lld fields exceed the buffers (%lld / %lld)", (long long)n_rows, (long long)total,
                  (long long)cap_rows, (long long)cap_fields);
        return R4D_E_WORKSPACE;
    }
    row_off[0] = 0;
    for (int64_t r = 0; r < n_rows; ++r) row_off[r + 1] = row_off[r] + fields[r];
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)(hw ? (hw > 16 ? 16 : hw) : 1);
    if ((int64_t)n_thr > n_rows / 256 + 1) n_thr = (int)(n_rows / 256 + 1);
    std::vector<int64_t> bad(n_thr, -1);
    auto work = [&](int t) {
        const int64_t r0 = n_rows * t / n_thr, r1 = n_rows * (t + 1) / n_thr;
        for (int64_t r = r0; r < r1; ++r) {
            T* dst = values + row_off[r];
            const char* c = lines[r].beg;
            const char* const e = lines[r].end;
            while (c < e) {
                while (c < e && is_space(*c)) ++c;
                if (c >= e) break;
                const char* f = c;
                while (c < e && !is_space(*c)) ++c;
                if (!parse_field<T>(f, c, dst++)) {
                    bad[t] = r;
                    return;
                }
            }
        }
    };
    if (n_thr <= 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < n_thr; ++t)
        if (bad[t] >= 0) {
            set_error("parse_rows: malformed number in row %lld", (long long)bad[t]);
            return R4D_E_ARG;
        }
    return n_rows;
}

}  // namespace r4d

extern "C" {

int64_t r4d_parse_rows_count(const char* text, size_t len, int64_t* n_fields) {
    if ((!text && len) || !n_fields) return R4D_E_ARG;
    std::vector<r4d::Line> lines;
    std::vector<int64_t> fields;
    if (!r4d::split_lines(text, len, lines, fields)) {
        r4d::set_error("parse_rows_count: non-ASCII byte in the text (only ASCII index / score files are accepted)");
        return R4D_E_ARG;
    }
    int64_t total = 0;
    for (int64_t n : fields) total += n;
    *n_fields = total;
    return (int64_t)lines.size();
}

int64_t r4d_parse_int_rows(const char* text, size_t len, int64_t* row_off, int64_t* values, int64_t cap_rows,
                           int64_t cap_fields) {
    return r4d::parse_rows<int64_t>(text, len, row_off, values, cap_rows, cap_fields);
}

int64_t r4d_parse_float_rows(const char* text, size_t len, int64_t* row_off, double* values, int64_t cap_rows,
                             int64_t cap_fields) {
    return r4d::parse_rows<double>(text, len, row_off, values, cap_rows, cap_fields);
}

}  // extern "C"

// nb200_hostfmt.cu -- host-side text formatting of result tracks (no device code): the bedgraph run-length
// writer of pyatac/tracks.py:37-74 with the reference's number format (Python-2 str(float) = "%.12g", plus ".0"
// on integral values).  Formatting ~10^4 runs per track per chunk in Python costs more than scoring the chunk on
// the GPU, so the writer lives next to the kernels.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <charconv>

#include "../../include/nucleo_b200.h"

// "%.12g" through std::to_chars (same correctly rounded digits and exponent form as printf in the C locale, several times
// faster), then Python 2's ".0" on values that print as integers.
static inline int fmt12(char *dst, double x)
{
    if (x != x) {
        memcpy(dst, "nan", 4);
        return 3;
    }
    if (isinf(x)) {
        if (x > 0) {
            memcpy(dst, "inf", 4);
            return 3;
        }
        memcpy(dst, "-inf", 5);
        return 4;
    }
    const std::to_chars_result r = std::to_chars(dst, dst + 40, x, std::chars_format::general, 12);
    int n = (int)(r.ptr - dst);
    bool plain = true;
    for (int i = 0; i < n; i++)
        if (dst[i] == '.' || dst[i] == 'e') plain = false;
    if (plain) {
        dst[n++] = '.';
        dst[n++] = '0';
    }
    dst[n] = 0;
    return n;
}

static inline int fmt_i64(char *dst, long long v)
{
    const std::to_chars_result r = std::to_chars(dst, dst + 24, v);
    return (int)(r.ptr - dst);
}

extern "C" {

// Returns the number of bytes the text needs; writes it when it fits into `cap` (no terminator).  Semantics of
// Track.write_track: consecutive equal values merge into one row; NaN rows are skipped; a run that is directly
// followed by NaN is never flushed (reference quirk, tracks.py:59-60); zeros are written unless write_zero == 0.
int64_t nb200_format_track(const char *chrom, int64_t start, const double *vals, int64_t n, int32_t write_zero, char *out,
                           int64_t cap)
{
    const size_t lc = strlen(chrom);
    int64_t used = 0;
    char row[160];
    int64_t i = 0;
    while (i < n) {
        const double v = vals[i];
        int64_t j = i + 1;
        if (v != v) {
            while (j < n && vals[j] != vals[j]) j++;
        } else {
            while (j < n && vals[j] == v) j++;
            const bool followed_by_nan = (j < n) && (vals[j] != vals[j]);
            if (!followed_by_nan && (write_zero || v != 0.0)) {
                int m = 0;
                row[m++] = '\t';
                m += fmt_i64(row + m, (long long)(start + i));
                row[m++] = '\t';
                m += fmt_i64(row + m, (long long)(start + j));
                row[m++] = '\t';
                m += fmt12(row + m, v);
                row[m++] = '\n';
                if (used + (int64_t)lc + m <= cap && out) {
                    memcpy(out + used, chrom, lc);
                    memcpy(out + used + lc, row, (size_t)m);
                }
                used += (int64_t)lc + m;
            }
        }
        i = j;
    }
    return used;
}

}  // extern "C"

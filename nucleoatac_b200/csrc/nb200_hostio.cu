// nb200_hostio.cu -- host-side (no device code) BGZF compression + tabix indexing of the text outputs:
// what the reference does with pysam.tabix_compress + pysam.tabix_index(preset="bed") at the end of run_occ /
// run_nuc (nucleoatac/run_occ.py:130-136, run_nuc.py:194-201).  One pass over the plain BED / bedgraph file:
// 0xff00-byte blocks are deflated by a pool of threads, the .tbi (UCSC binning + 16 kb linear index, same layout as
// htslib writes) is built from the first three columns of every row.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <charconv>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nucleo_b200.h"

namespace {

const size_t BLOCK = 0xff00;

struct Chunk64 {
    uint64_t beg, end;  // uncompressed offsets until finalisation
};

struct RefIndex {
    std::map<uint32_t, std::vector<Chunk64>> bins;
    std::vector<int64_t> linear;  // -1 = unset
};

inline uint32_t reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// one BGZF block: gzip member with the BC extra field
size_t bgzf_block(const unsigned char *src, size_t n, unsigned char *dst, int level)
{
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<unsigned char *>(src);
    zs.avail_in = (uInt)n;
    zs.next_out = dst + 18;
    zs.avail_out = 65536 - 18 - 8;
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    const unsigned char head[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
    memcpy(dst, head, 12);
    dst[12] = 66;
    dst[13] = 67;
    dst[14] = 2;
    dst[15] = 0;
    const uint16_t bsize = (uint16_t)(clen + 25);
    dst[16] = (unsigned char)(bsize & 0xff);
    dst[17] = (unsigned char)(bsize >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n);
    const uint32_t isize = (uint32_t)n;
    memcpy(dst + 18 + clen, &crc, 4);
    memcpy(dst + 22 + clen, &isize, 4);
    return clen + 26;
}

const unsigned char BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// deflate n bytes (whole BLOCK-sized blocks, the last one may be short) as BGZF blocks into `fh`, `threads` blocks at a
// time; appends the compressed offset of every block to coff and advances pos
template <typename Meanwhile>
bool bgzf_write_blocks(FILE *fh, const unsigned char *data, size_t n, int threads, int level, std::vector<uint64_t> &coff, uint64_t &pos,
                       Meanwhile meanwhile)
{
    const size_t nblk = (n + BLOCK - 1) / BLOCK;
    if (threads < 1) threads = 1;
    std::vector<std::vector<unsigned char>> out(nblk);
    std::vector<size_t> olen(nblk, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            for (size_t b = (size_t)t; b < nblk; b += (size_t)threads) {
                const size_t len = (b + 1) * BLOCK <= n ? BLOCK : n - b * BLOCK;
                out[b].resize(65536);
                olen[b] = bgzf_block(data + b * BLOCK, len, out[b].data(), level);
            }
        });
    meanwhile();   // the caller's own work on the same (read-only) text while the workers deflate it
    for (auto &th : pool) th.join();
    for (size_t b = 0; b < nblk; b++) {
        coff.push_back(pos);
        if (fwrite(out[b].data(), 1, olen[b], fh) != olen[b]) return false;
        pos += olen[b];
    }
    return true;
}

// deflate `data` as BGZF into `fh` with the EOF marker; returns compressed offsets of every block (+ the end offset)
bool bgzf_write(FILE *fh, const std::vector<unsigned char> &data, int threads, int level, std::vector<uint64_t> &coff)
{
    coff.clear();
    uint64_t pos = 0;
    if (!bgzf_write_blocks(fh, data.data(), data.size(), threads, level, coff, pos, []() {})) return false;
    coff.push_back(pos);
    return fwrite(BGZF_EOF, 1, 28, fh) == 28;
}

// decimal integer in [p, end) (no terminator needed: the text buffer is not NUL-terminated); returns the first byte after
// the digits, or nullptr when there is no digit
const char *parse_i64(const char *p, const char *end, int64_t &v)
{
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) neg = *p++ == '-';
    if (p >= end || *p < '0' || *p > '9') return nullptr;
    int64_t x = 0;
    while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
    v = neg ? -x : x;
    return p;
}

// virtual offset of uncompressed position u; a position at the very end of the data is the start of the block that
// would follow (what bgzf_tell reports after the last line)
inline uint64_t voffset(uint64_t u, const std::vector<uint64_t> &coff, uint64_t total)
{
    if (u >= total) return coff.back() << 16;
    return (coff[u / BLOCK] << 16) | (u % BLOCK);
}

}  // namespace

extern "C" {

// plain sorted BED / bedgraph -> path_gz (BGZF) + path_gz.tbi (preset "bed": 0-based, columns 1-2-3, '#' comments).
// Returns 0 on success, non-zero on an I/O or format error (message into err[0..errcap) when given).
int nb200_bgzip_tabix(const char *path_plain, const char *path_gz, int threads, char *err, int errcap)
{
    return nb200_bgzip_tabix_level(path_plain, path_gz, threads, -1, err, errcap);
}

// The same with a deflate level: -1 = zlib's default (6: the level htslib writes with, byte-identical files), 1..9 as zlib.
// Level 1 deflates bedgraph text about five times faster for files ~7 % larger; the index does not depend on it.
int nb200_bgzip_tabix_level(const char *path_plain, const char *path_gz, int threads, int level, char *err, int errcap)
{
    if (level < -1 || level > 9 || level == 0) {
        if (err && errcap > 0) snprintf(err, (size_t)errcap, "deflate level must be -1 (default) or 1..9");
        return 1;
    }
    if (level < 0) level = 6;
    auto fail = [&](const char *msg) {
        if (err && errcap > 0) snprintf(err, (size_t)errcap, "%s", msg);
        return 1;
    };
    FILE *in = fopen(path_plain, "rb");
    if (!in) return fail("cannot open the input file");
    FILE *out = fopen(path_gz, "wb");
    if (!out) {
        fclose(in);
        return fail("cannot open the output file");
    }
    // The file streams through in waves of whole BGZF blocks (genome-wide tracks are many GB, the reference streams them
    // through tabix_compress too): a wave is deflated by the thread pool and written, and its rows are indexed by their
    // uncompressed offsets; only the per-block compressed offsets and the index stay in memory.
    if (threads < 1) threads = 1;
    const size_t wave_bytes = (size_t)threads * 64 * BLOCK;
    std::vector<unsigned char> wave(wave_bytes);
    std::vector<std::string> names;
    std::map<std::string, int> tid;
    std::vector<RefIndex> refs;
    std::vector<uint64_t> coff;
    uint64_t cpos = 0, N = 0;        // compressed / uncompressed bytes so far
    int cur = -1;
    std::string cur_name, carry;     // carry: the start of a row that continues in the next wave (it begins at offset N - carry.size())
    const char *bad = nullptr;
    auto index_row = [&](const char *row, size_t len, uint64_t off) {   // row [off, off + len) incl. its newline, if any
        if (row[0] == '#' || len <= 1) return;
        const char *end = row + len;
        const char *t1 = (const char *)memchr(row, '\t', len);
        if (!t1) {
            bad = "row without tab-separated columns";
            return;
        }
        const size_t ln = (size_t)(t1 - row);
        if (cur < 0 || cur_name.size() != ln || memcmp(cur_name.data(), row, ln) != 0) {
            cur_name.assign(row, ln);
            auto it = tid.find(cur_name);
            if (it == tid.end()) {
                cur = (int)names.size();
                tid[cur_name] = cur;
                names.push_back(cur_name);
                refs.emplace_back();
            } else
                cur = it->second;
        }
        int64_t beg = 0, stop = 0;
        const char *q = parse_i64(t1 + 1, end, beg);
        if (!q || q >= end || *q != '\t') {
            bad = "bad start column";
            return;
        }
        q = parse_i64(q + 1, end, stop);
        if (!q) {
            bad = "bad end column";
            return;
        }
        if (stop <= beg) stop = beg + 1;
        RefIndex &r = refs[cur];
        auto &chunks = r.bins[reg2bin(beg, stop)];
        if (!chunks.empty() && chunks.back().end == off)
            chunks.back().end = off + len;
        else
            chunks.push_back({off, off + len});
        const int64_t w0 = beg >> 14, w1 = (stop - 1) >> 14;
        if ((int64_t)r.linear.size() <= w1) r.linear.resize((size_t)w1 + 1, -1);
        for (int64_t w = w0; w <= w1; w++)
            if (r.linear[w] < 0) r.linear[w] = (int64_t)off;
    };
    bool ok = true;
    for (;;) {
        size_t got = 0, n;
        while (got < wave_bytes && (n = fread(wave.data() + got, 1, wave_bytes - got, in)) > 0) got += n;
        if (got == 0) break;
        // the rows of this wave are indexed on this thread while the workers deflate it; a row that began in the previous wave
        // is completed first
        auto index_wave = [&]() {
            size_t p = 0;
            if (!carry.empty()) {
                const unsigned char *nl = (const unsigned char *)memchr(wave.data(), '\n', got);
                const size_t take = nl ? (size_t)(nl - wave.data()) + 1 : got;
                const uint64_t off = N - carry.size();
                carry.append((const char *)wave.data(), take);
                p = take;
                if (nl) {
                    index_row(carry.data(), carry.size(), off);
                    carry.clear();
                }
            }
            while (p < got && !bad) {
                const unsigned char *nl = (const unsigned char *)memchr(wave.data() + p, '\n', got - p);
                if (!nl) {
                    carry.assign((const char *)wave.data() + p, got - p);
                    break;
                }
                const size_t e = (size_t)(nl - wave.data()) + 1;
                index_row((const char *)wave.data() + p, e - p, N + p);
                p = e;
            }
        };
        ok = bgzf_write_blocks(out, wave.data(), got, threads, level, coff, cpos, index_wave);
        if (!ok) break;
        N += got;
        if (bad || got < wave_bytes) break;
    }
    fclose(in);
    if (ok && !bad && !carry.empty()) index_row(carry.data(), carry.size(), N - carry.size());   // last row without a newline
    if (ok && !bad) {
        coff.push_back(cpos);
        ok = fwrite(BGZF_EOF, 1, 28, out) == 28;
    }
    fclose(out);
    if (bad) return fail(bad);
    if (!ok) return fail("write error");
    // ---- .tbi
    std::vector<unsigned char> tbi;
    auto put32 = [&](int32_t v) { tbi.insert(tbi.end(), (unsigned char *)&v, (unsigned char *)&v + 4); };
    auto put64 = [&](uint64_t v) { tbi.insert(tbi.end(), (unsigned char *)&v, (unsigned char *)&v + 8); };
    tbi.insert(tbi.end(), {'T', 'B', 'I', 1});
    put32((int32_t)names.size());
    put32(0x10000);
    put32(1);
    put32(2);
    put32(3);
    put32('#');
    put32(0);
    int32_t l_nm = 0;
    for (auto &n : names) l_nm += (int32_t)n.size() + 1;
    put32(l_nm);
    for (auto &n : names) {
        tbi.insert(tbi.end(), n.begin(), n.end());
        tbi.push_back(0);
    }
    for (auto &r : refs) {
        put32((int32_t)r.bins.size());
        for (auto &kv : r.bins) {
            put32((int32_t)kv.first);
            put32((int32_t)kv.second.size());
            for (auto &c : kv.second) {
                put64(voffset(c.beg, coff, N));
                put64(voffset(c.end, coff, N));
            }
        }
        // windows without rows: leading ones point at the first row, later ones repeat the previous window
        int64_t prev = 0;
        for (auto v : r.linear)
            if (v >= 0) {
                prev = v;
                break;
            }
        put32((int32_t)r.linear.size());
        for (auto v : r.linear) {
            if (v >= 0) prev = v;
            put64(voffset((uint64_t)prev, coff, N));
        }
    }
    std::string tpath = std::string(path_gz) + ".tbi";
    FILE *tf = fopen(tpath.c_str(), "wb");
    if (!tf) return fail("cannot open the index file");
    std::vector<uint64_t> tcoff;
    const bool ok2 = bgzf_write(tf, tbi, 1, 6, tcoff);   // the index itself always at the default level
    fclose(tf);
    return ok2 ? 0 : fail("index write error");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Region read of a bgzip'd, tabix-indexed bedgraph into a dense array: BedGraphFile.read of pyatac/bedgraph.py:6-16, which the
// reference does through pysam.Tabixfile.fetch (`--occ_track` of `nucleoatac nuc`, the tracks `nucleoatac nfr` reads).  The
// linear index of the .tbi gives the block to start from; rows are scanned until the chromosome changes or a row starts at
// or past `end`.
namespace {

struct TbiLite {
    std::vector<std::string> names;
    std::vector<std::vector<uint64_t>> linear;
    int sc = 1, bc = 2, ec = 3;
};

bool load_tbi(const std::string &path, TbiLite &t, std::string &err)
{
    gzFile g = gzopen(path.c_str(), "rb");   // BGZF is a multi-member gzip file
    if (!g) {
        err = "cannot open the .tbi";
        return false;
    }
    std::vector<unsigned char> raw;
    unsigned char buf[1 << 16];
    int n;
    while ((n = gzread(g, buf, sizeof(buf))) > 0) raw.insert(raw.end(), buf, buf + n);
    gzclose(g);
    auto i32 = [&](size_t off, int32_t &v) {
        if (off + 4 > raw.size()) return false;
        memcpy(&v, raw.data() + off, 4);
        return true;
    };
    int32_t n_ref = 0, fmt = 0, sc = 0, bc = 0, ec = 0, meta = 0, skip = 0, l_nm = 0;
    if (raw.size() < 36 || memcmp(raw.data(), "TBI\1", 4) != 0 || !i32(4, n_ref) || !i32(8, fmt) || !i32(12, sc) || !i32(16, bc) || !i32(20, ec) ||
        !i32(24, meta) || !i32(28, skip) || !i32(32, l_nm) || n_ref < 0 || l_nm < 0 || 36 + (size_t)l_nm > raw.size() || sc < 1 || bc < 1 || ec < 1) {
        err = "not a tabix index";
        return false;
    }
    t.sc = sc;
    t.bc = bc;
    t.ec = ec;
    size_t off = 36;
    for (size_t p = off; p < off + (size_t)l_nm;) {
        const void *z = memchr(raw.data() + p, 0, off + (size_t)l_nm - p);
        if (!z) break;
        t.names.emplace_back((const char *)raw.data() + p, (const char *)z);
        p = (size_t)((const unsigned char *)z - raw.data()) + 1;
    }
    off += (size_t)l_nm;
    for (int r = 0; r < n_ref; r++) {
        int32_t n_bin = 0;
        if (!i32(off, n_bin) || n_bin < 0) goto bad;
        off += 4;
        for (int b = 0; b < n_bin; b++) {
            int32_t n_chunk = 0;
            if (!i32(off + 4, n_chunk) || n_chunk < 0) goto bad;
            off += 8 + 16 * (size_t)n_chunk;
        }
        int32_t n_intv = 0;
        if (!i32(off, n_intv) || n_intv < 0 || off + 4 + 8 * (size_t)n_intv > raw.size()) goto bad;
        off += 4;
        std::vector<uint64_t> lin((size_t)n_intv);
        if (n_intv) memcpy(lin.data(), raw.data() + off, 8 * (size_t)n_intv);
        off += 8 * (size_t)n_intv;
        t.linear.push_back(std::move(lin));
    }
    return true;
bad:
    err = "truncated tabix index";
    return false;
}

// inflate the BGZF block at compressed offset coff into dst (cleared); returns its size on disk, 0 at the end of the file, -1 on error
long read_bgzf_block(FILE *fh, uint64_t coff, std::vector<unsigned char> &dst)
{
    dst.clear();
    if (fseeko(fh, (off_t)coff, SEEK_SET) != 0) return -1;
    unsigned char head[18];
    const size_t got = fread(head, 1, 18, fh);
    if (got == 0) return 0;
    if (got < 18 || head[0] != 0x1f || head[1] != 0x8b || head[2] != 8 || head[3] != 4) return -1;
    const unsigned xlen = head[10] | (head[11] << 8);
    if (xlen < 6) return -1;
    std::vector<unsigned char> extra(xlen);
    memcpy(extra.data(), head + 12, 6);
    if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, fh) != xlen - 6) return -1;
    long bsize = -1;
    for (size_t o = 0; o + 4 <= extra.size();) {
        const unsigned slen = extra[o + 2] | (extra[o + 3] << 8);
        if (extra[o] == 66 && extra[o + 1] == 67 && o + 6 <= extra.size()) bsize = (long)(extra[o + 4] | (extra[o + 5] << 8)) + 1;
        o += 4 + slen;
    }
    const long clen = bsize - 12 - (long)xlen - 8;
    if (bsize < 0 || clen < 0) return -1;
    std::vector<unsigned char> cdata((size_t)clen);
    if (clen && fread(cdata.data(), 1, (size_t)clen, fh) != (size_t)clen) return -1;
    dst.resize(65536);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    zs.next_in = cdata.data();
    zs.avail_in = (uInt)clen;
    zs.next_out = dst.data();
    zs.avail_out = (uInt)dst.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t produced = zs.total_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    dst.resize(produced);
    return bsize;
}

}  // namespace

extern "C" int nb200_bedgraph_fetch(const char *path_gz, const char *chrom, int64_t start, int64_t end, double empty, double *out, char *err,
                                    int errcap)
{
    auto fail = [&](const char *msg) {
        if (err && errcap > 0) snprintf(err, (size_t)errcap, "%s", msg);
        return 1;
    };
    if (!path_gz || !chrom || !out || end < start) return fail("bad argument");
    const int64_t len = end - start;
    for (int64_t i = 0; i < len; i++) out[i] = empty;   // np.ones(end - start) * empty, pyatac/bedgraph.py:10
    TbiLite tbi;
    std::string e;
    if (!load_tbi(std::string(path_gz) + ".tbi", tbi, e)) return fail(e.c_str());
    size_t t = 0;
    const size_t lc = strlen(chrom);
    while (t < tbi.names.size() && tbi.names[t] != chrom) t++;
    if (t >= tbi.names.size() || t >= tbi.linear.size() || tbi.linear[t].empty() || len == 0) return 0;   // nothing on this chromosome
    const std::vector<uint64_t> &lin = tbi.linear[t];
    const uint64_t voff = lin[std::min<size_t>((size_t)(std::max<int64_t>(start, 0) >> 14), lin.size() - 1)];
    FILE *fh = fopen(path_gz, "rb");
    if (!fh) return fail("cannot open the bedgraph");
    uint64_t coff = voff >> 16;
    size_t skip = (size_t)(voff & 0xffff);
    std::vector<unsigned char> block;
    std::string line;           // a row may straddle blocks
    bool matched = false, done = false;
    const int maxcol = std::max(std::max(tbi.sc, tbi.bc), std::max(tbi.ec, 3));
    while (!done) {
        const long size = read_bgzf_block(fh, coff, block);
        if (size < 0) {
            fclose(fh);
            return fail("corrupt BGZF block");
        }
        if (size == 0) break;
        coff += (uint64_t)size;
        size_t p = std::min(skip, block.size());
        skip = 0;
        while (p < block.size() && !done) {
            const unsigned char *nl = (const unsigned char *)memchr(block.data() + p, '\n', block.size() - p);
            if (!nl) {
                line.append((const char *)block.data() + p, block.size() - p);
                break;
            }
            const char *row;
            size_t rl;
            if (!line.empty()) {
                line.append((const char *)block.data() + p, (size_t)(nl - (block.data() + p)));
                row = line.data();
                rl = line.size();
            } else {
                row = (const char *)block.data() + p;
                rl = (size_t)(nl - (block.data() + p));
            }
            p = (size_t)(nl - block.data()) + 1;
            // columns
            const char *col[8] = {nullptr};
            size_t cl[8] = {0};
            int nc = 0;
            for (size_t a = 0; nc < 8;) {
                const char *tab = (const char *)memchr(row + a, '\t', rl - a);
                col[nc] = row + a;
                cl[nc] = tab ? (size_t)(tab - (row + a)) : rl - a;
                nc++;
                if (!tab) break;
                a = (size_t)(tab - row) + 1;
            }
            const bool usable = nc >= maxcol && maxcol <= 8 && !(rl > 0 && row[0] == '#');   // short rows and comments are skipped
            if (usable) {
                if (cl[tbi.sc - 1] != lc || memcmp(col[tbi.sc - 1], chrom, lc) != 0) {
                    if (matched) done = true;
                } else {
                    int64_t b = 0, en = 0;
                    if (!parse_i64(col[tbi.bc - 1], col[tbi.bc - 1] + cl[tbi.bc - 1], b) || !parse_i64(col[tbi.ec - 1], col[tbi.ec - 1] + cl[tbi.ec - 1], en)) {
                        fclose(fh);
                        return fail("bad coordinate column");
                    }
                    if (b >= end)
                        done = true;
                    else if (en > start) {
                        matched = true;
                        if (nc < 4) {
                            fclose(fh);
                            return fail("row without a value column");
                        }
                        double v = 0.0;
                        const char *vb = col[3], *ve = col[3] + cl[3];
                        while (vb < ve && (*vb == ' ' || *vb == '+')) vb++;
                        const std::from_chars_result fr = std::from_chars(vb, ve, v);
                        if (fr.ec != std::errc()) {
                            fclose(fh);
                            return fail("bad value column");
                        }
                        const int64_t lo = std::max<int64_t>(b - start, 0), hi = std::min<int64_t>(en - start, len);
                        for (int64_t i = lo; i < hi; i++) out[i] = v;
                    }
                }
            }
            line.clear();
        }
    }
    fclose(fh);
    return 0;
}

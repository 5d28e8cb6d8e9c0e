// nb200_hostio.cu -- host-side (no device code) BGZF compression + tabix indexing of the text outputs:
// what the reference does with pysam.tabix_compress + pysam.tabix_index(preset="bed") at the end of run_occ /
// run_nuc (nucleoatac/run_occ.py:130-136, run_nuc.py:194-201).  One pass over the plain BED / bedgraph file:
// 0xff00-byte blocks are deflated by a pool of threads, the .tbi (UCSC binning + 16 kb linear index, same layout as
// htslib writes) is built from the first three columns of every row.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

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

// deflate `data` as BGZF into `fh`; returns compressed offsets of every block (+ the end offset)
bool bgzf_write(FILE *fh, const std::vector<unsigned char> &data, int threads, int level, std::vector<uint64_t> &coff)
{
    const size_t nblk = (data.size() + BLOCK - 1) / BLOCK;
    std::vector<std::vector<unsigned char>> out(nblk);
    std::vector<size_t> olen(nblk, 0);
    if (threads < 1) threads = 1;
    const size_t wave = (size_t)threads * 64;
    coff.clear();
    uint64_t pos = 0;
    for (size_t b0 = 0; b0 < nblk; b0 += wave) {
        const size_t b1 = b0 + wave < nblk ? b0 + wave : nblk;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t]() {
                for (size_t b = b0 + (size_t)t; b < b1; b += (size_t)threads) {
                    const size_t n = (b + 1) * BLOCK <= data.size() ? BLOCK : data.size() - b * BLOCK;
                    out[b].resize(65536);
                    olen[b] = bgzf_block(data.data() + b * BLOCK, n, out[b].data(), level);
                }
            });
        for (auto &th : pool) th.join();
        for (size_t b = b0; b < b1; b++) {
            coff.push_back(pos);
            if (fwrite(out[b].data(), 1, olen[b], fh) != olen[b]) return false;
            pos += olen[b];
            std::vector<unsigned char>().swap(out[b]);
        }
    }
    coff.push_back(pos);
    return fwrite(BGZF_EOF, 1, 28, fh) == 28;
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
    auto fail = [&](const char *msg) {
        if (err && errcap > 0) snprintf(err, (size_t)errcap, "%s", msg);
        return 1;
    };
    FILE *in = fopen(path_plain, "rb");
    if (!in) return fail("cannot open the input file");
    std::vector<unsigned char> data;
    {
        unsigned char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), in)) > 0) data.insert(data.end(), buf, buf + n);
        fclose(in);
    }
    // ---- index from the text
    std::vector<std::string> names;
    std::map<std::string, int> tid;
    std::vector<RefIndex> refs;
    size_t p = 0;
    const size_t N = data.size();
    int cur = -1;
    std::string cur_name;
    while (p < N) {
        const unsigned char *nl = (const unsigned char *)memchr(data.data() + p, '\n', N - p);
        const size_t e = nl ? (size_t)(nl - data.data()) + 1 : N;
        if (data[p] != '#' && e - p > 1) {
            const unsigned char *t1 = (const unsigned char *)memchr(data.data() + p, '\t', e - p);
            if (!t1) return fail("row without tab-separated columns");
            const size_t ln = (size_t)(t1 - (data.data() + p));
            if (cur < 0 || cur_name.size() != ln || memcmp(cur_name.data(), data.data() + p, ln) != 0) {
                cur_name.assign((const char *)data.data() + p, ln);
                auto it = tid.find(cur_name);
                if (it == tid.end()) {
                    cur = (int)names.size();
                    tid[cur_name] = cur;
                    names.push_back(cur_name);
                    refs.emplace_back();
                } else
                    cur = it->second;
            }
            char *endp;
            const int64_t beg = strtoll((const char *)t1 + 1, &endp, 10);
            if (*endp != '\t') return fail("bad start column");
            int64_t end = strtoll(endp + 1, &endp, 10);
            if (end <= beg) end = beg + 1;
            RefIndex &r = refs[cur];
            auto &chunks = r.bins[reg2bin(beg, end)];
            if (!chunks.empty() && chunks.back().end == p)
                chunks.back().end = e;
            else
                chunks.push_back({(uint64_t)p, (uint64_t)e});
            const int64_t w0 = beg >> 14, w1 = (end - 1) >> 14;
            if ((int64_t)r.linear.size() <= w1) r.linear.resize((size_t)w1 + 1, -1);
            for (int64_t w = w0; w <= w1; w++)
                if (r.linear[w] < 0) r.linear[w] = (int64_t)p;
        }
        p = e;
    }
    // ---- compress
    FILE *out = fopen(path_gz, "wb");
    if (!out) return fail("cannot open the output file");
    std::vector<uint64_t> coff;
    const bool ok = bgzf_write(out, data, threads, 6, coff);
    fclose(out);
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
    const bool ok2 = bgzf_write(tf, tbi, 1, 6, tcoff);
    fclose(tf);
    return ok2 ? 0 : fail("index write error");
}

}  // extern "C"

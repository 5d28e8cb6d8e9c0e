// nb200_bamio.cu -- host-side (no device code) BAM region reader: what pysam's AlignmentFile.fetch + the per-read loop of
// pyatac/fragments.pyx:21-25 (and :47-50, :128-131) give the scoring path -- the (pos, tlen) of the proper-pair forward
// reads overlapping a region -- for MANY regions at once: the regions are decoded by a pool of threads (BGZF blocks
// inflated with zlib from pread(2) on one descriptor), because with the scoring on the device the BAM decode is what a
// run on real files waits for.  The caller resolves each region's starting virtual offset from the .bai linear index
// (hostio.py); records are then scanned in file order until the reference changes or pos >= end, exactly like
// hostio.BamFile._fetch_indexed (kept as the cross-check).
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nucleo_b200.h"

namespace {

struct BlockReader {
    int fd;
    uint64_t coff;                       // compressed offset of the next block
    std::vector<unsigned char> cbuf, buf;   // compressed block, inflated bytes not yet consumed
    size_t p = 0;                        // read position in buf
    z_stream zs;
    bool zs_ok = false;
    std::string err;

    BlockReader(int fd_, uint64_t coff_) : fd(fd_), coff(coff_), cbuf(65536 + 64) {}
    ~BlockReader()
    {
        if (zs_ok) inflateEnd(&zs);
    }
    // append the next BGZF block to buf; false at a clean end of file (no byte left at a block boundary, err empty) or on
    // error (err set: a file that stops in the middle of a block is truncated, not finished)
    bool next_block()
    {
        unsigned char head[18];
        ssize_t got = pread(fd, head, 18, (off_t)coff);
        if (got == 0) return false;
        if (got < 18) {
            err = got < 0 ? "read error" : "truncated BGZF block header";
            return false;
        }
        if (!(head[0] == 31 && head[1] == 139 && head[2] == 8 && head[3] == 4)) {
            err = "not a BGZF block";
            return false;
        }
        const unsigned xlen = head[10] | (head[11] << 8);
        // the BC subfield is the first (and only) one in every BGZF writer in use; walk the extra field otherwise
        unsigned bsize = 0;
        if (head[12] == 66 && head[13] == 67 && xlen >= 6)
            bsize = (head[16] | (head[17] << 8)) + 1u;
        else {
            std::vector<unsigned char> extra(xlen);
            if (pread(fd, extra.data(), xlen, (off_t)coff + 12) < (ssize_t)xlen) {
                err = "truncated BGZF block header";
                return false;
            }
            for (unsigned o = 0; o + 4 <= xlen;) {
                const unsigned slen = extra[o + 2] | (extra[o + 3] << 8);
                if (extra[o] == 66 && extra[o + 1] == 67 && o + 6 <= xlen) bsize = (extra[o + 4] | (extra[o + 5] << 8)) + 1u;
                o += 4 + slen;
            }
        }
        if (bsize < 12 + xlen + 8) {
            err = "BGZF block without BC field";
            return false;
        }
        const size_t clen = bsize - 12 - xlen - 8;
        if (pread(fd, cbuf.data(), clen + 8, (off_t)coff + 12 + xlen) < (ssize_t)(clen + 8)) {
            err = "truncated BGZF block";
            return false;
        }
        const uint32_t isize = cbuf[clen + 4] | (cbuf[clen + 5] << 8) | (cbuf[clen + 6] << 16) | ((uint32_t)cbuf[clen + 7] << 24);
        if (isize > 65536) {   // a BGZF block inflates to at most 64 KiB: anything else is a corrupt (or hostile) trailer
            err = "BGZF block with an impossible uncompressed size";
            return false;
        }
        coff += bsize;
        if (isize == 0) return true;     // empty block (e.g. the EOF marker): nothing to append
        if (p > 0) {                      // drop what has been consumed
            buf.erase(buf.begin(), buf.begin() + p);
            p = 0;
        }
        const size_t old = buf.size();
        buf.resize(old + isize);
        if (!zs_ok) {
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) {
                err = "inflateInit2 failed";
                return false;
            }
            zs_ok = true;
        } else
            inflateReset(&zs);
        zs.next_in = cbuf.data();
        zs.avail_in = (uInt)clen;
        zs.next_out = buf.data() + old;
        zs.avail_out = isize;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) {
            err = "inflate failed";
            return false;
        }
        return true;
    }
    // make n bytes available at buf[p..]; false when the file ends first
    bool need(size_t n)
    {
        while (buf.size() - p < n) {
            const uint64_t before = coff;
            if (!next_block()) return false;
            if (coff == before) return false;
        }
        return true;
    }
};

inline int32_t rd_i32(const unsigned char *q) { return (int32_t)(q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24)); }

// hostio.BamFile._records + the filter of _fetch_indexed: reads of `tid` from the virtual offset on, until the reference
// changes or pos >= end; kept when proper pair (0x2), not reverse (0x10) and pos + max(l_seq, 1) + 64 > start (a superset
// of htslib's overlap test by alignment end -- every consumer re-checks its cell bounds like fragments.pyx:37 does).
bool fetch_region(int fd, uint64_t voffset, int32_t tid, int32_t start, int32_t end, std::vector<int32_t> &pos,
                  std::vector<int32_t> &tlen, std::string &err)
{
    BlockReader br(fd, voffset >> 16);
    const size_t uoff = (size_t)(voffset & 0xffff);
    if (!br.need(uoff)) {
        err = br.err;
        return br.err.empty();   // offset past the end of the file: no reads
    }
    br.p = uoff;
    for (;;) {
        if (!br.need(4)) {
            if (br.err.empty() && br.buf.size() - br.p > 0) br.err = "truncated BAM record";   // file ends inside a record length
            break;
        }
        const int32_t bs = rd_i32(br.buf.data() + br.p);
        if (bs < 32) {
            err = "corrupt BAM record";
            return false;
        }
        if (!br.need(4 + (size_t)bs)) {
            if (br.err.empty()) br.err = "truncated BAM record";   // the file ends in the middle of a record
            break;
        }
        const unsigned char *q = br.buf.data() + br.p + 4;
        const int32_t rtid = rd_i32(q), rpos = rd_i32(q + 4);
        const unsigned flag = q[14] | (q[15] << 8);
        const int32_t l_seq = rd_i32(q + 16), rtlen = rd_i32(q + 28);
        br.p += 4 + (size_t)bs;
        if (rtid != tid || rpos >= end) break;
        const int64_t rend = (int64_t)rpos + (l_seq > 1 ? l_seq : 1) + 64;
        if (rend > start && (flag & 0x2) && !(flag & 0x10)) {
            pos.push_back(rpos);
            tlen.push_back(rtlen);
        }
    }
    err = br.err;
    return br.err.empty();
}

void set_err(char *err, int cap, const std::string &m)
{
    if (err && cap > 0) {
        strncpy(err, m.c_str(), (size_t)cap - 1);
        err[cap - 1] = 0;
    }
}

}  // namespace

extern "C" {

int nb200_bam_fetch_many(const char *path, int32_t n_regions, const uint64_t *voffset, const int32_t *tid, const int32_t *start,
                         const int32_t *end, int32_t threads, int64_t *frag_off, int32_t **pos_out, int32_t **tlen_out, char *err,
                         int errcap)
{
    if (!path || n_regions < 0 || !frag_off || !pos_out || !tlen_out || (n_regions > 0 && (!voffset || !tid || !start || !end))) {
        set_err(err, errcap, "nb200_bam_fetch_many: bad argument");
        return NB200_ERR_ARG;
    }
    *pos_out = *tlen_out = nullptr;
    frag_off[0] = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_err(err, errcap, std::string("cannot open ") + path);
        return NB200_ERR_ARG;
    }
    std::vector<std::vector<int32_t>> ps(n_regions), ts(n_regions);
    std::vector<std::string> errs(n_regions);
    std::atomic<int> next(0), failed(0);
    auto work = [&]() {
        for (;;) {
            const int r = next.fetch_add(1);
            if (r >= n_regions) break;
            if (voffset[r] == 0 && tid[r] < 0) continue;   // region on an unknown reference: no reads
            if (!fetch_region(fd, voffset[r], tid[r], start[r], end[r], ps[r], ts[r], errs[r])) failed.store(1);
        }
    };
    int nt = threads < 1 ? 1 : threads;
    if (nt > n_regions) nt = n_regions > 0 ? n_regions : 1;
    if (nt == 1)
        work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    close(fd);
    if (failed.load()) {
        for (auto &e : errs)
            if (!e.empty()) {
                set_err(err, errcap, e + " in " + path);
                break;
            }
        return NB200_ERR_ARG;
    }
    for (int r = 0; r < n_regions; r++) frag_off[r + 1] = frag_off[r] + (int64_t)ps[r].size();
    const int64_t total = frag_off[n_regions];
    int32_t *P = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    int32_t *T = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    if (!P || !T) {
        free(P);
        free(T);
        set_err(err, errcap, "out of memory");
        return NB200_ERR_CAPACITY;
    }
    for (int r = 0; r < n_regions; r++)
        if (!ps[r].empty()) {
            memcpy(P + frag_off[r], ps[r].data(), sizeof(int32_t) * ps[r].size());
            memcpy(T + frag_off[r], ts[r].data(), sizeof(int32_t) * ts[r].size());
        }
    *pos_out = P;
    *tlen_out = T;
    return NB200_OK;
}

void nb200_free(void *p) { free(p); }

}  // extern "C"

"""Host I/O without pysam/htslib: BGZF + BAM (+ .bai region fetch), .fai-indexed FASTA, BGZF
writer, bgzip'd bedgraph reader.  Replaces what the reference gets from pysam on this path:
``AlignmentFile.fetch`` (pyatac/fragments.pyx:21-24), ``FastaFile.fetch`` (pyatac/seq.py:17-18),
``tabix_compress`` (nucleoatac/run_occ.py:130-136) and ``Tabixfile.fetch`` (pyatac/bedgraph.py:9-14).

Only what the occ / nuc paths need is decoded: for every alignment its reference id, position,
flag and template length; reads are filtered to ``is_proper_pair and not is_reverse`` exactly as
fragments.pyx:25 does, and handed to the device as int32 (pos, tlen) arrays.
"""
import gzip
import os
import struct
import zlib

import numpy as np

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# ----------------------------------------------------------------------------------------- BGZF
def _read_block(fh, coffset):
    """Inflate the BGZF block at compressed offset `coffset` -> (data, size of the block on disk)."""
    fh.seek(coffset)
    head = fh.read(18)
    if len(head) < 18:
        return b"", 0
    if head[:4] != b"\x1f\x8b\x08\x04":
        raise ValueError("not a BGZF block at offset %d" % coffset)
    xlen = struct.unpack_from("<H", head, 10)[0]
    extra = head[12:18] + fh.read(xlen - 6)
    bsize, off = None, 0
    while off + 4 <= len(extra):
        si1, si2, slen = extra[off], extra[off + 1], struct.unpack_from("<H", extra, off + 2)[0]
        if si1 == 66 and si2 == 67:
            bsize = struct.unpack_from("<H", extra, off + 4)[0] + 1
        off += 4 + slen
    if bsize is None:
        raise ValueError("BGZF block without BC field")
    cdata = fh.read(bsize - 12 - xlen - 8)
    fh.read(8)
    return zlib.decompress(cdata, -15), bsize


class BgzfWriter:
    """Write BGZF (what pysam.tabix_compress produces): <= 64 KiB deflate blocks + EOF marker."""

    def __init__(self, path):
        self.fh = open(path, "wb")
        self.buf = bytearray()

    def write(self, data):
        if isinstance(data, str):
            data = data.encode()
        self.buf += data
        while len(self.buf) >= 0xff00:
            self._flush(self.buf[:0xff00])
            del self.buf[:0xff00]

    def _flush(self, chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(bytes(chunk)) + c.flush()
        bsize = len(comp) + 25
        self.fh.write(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
        self.fh.write(comp)
        self.fh.write(struct.pack("<II", zlib.crc32(bytes(chunk)) & 0xffffffff, len(chunk)))

    def close(self):
        if self.buf:
            self._flush(self.buf)
            self.buf = bytearray()
        self.fh.write(_BGZF_EOF)
        self.fh.close()


def bgzip_file(src, dst):
    w = BgzfWriter(dst)
    with open(src, "rb") as fh:
        while True:
            block = fh.read(1 << 20)
            if not block:
                break
            w.write(block)
    w.close()


def host_workers(limit=16):
    """How many host threads / worker processes a helper may use: the cores this process may run on, at most `limit`, at
    most NB200_HOST_WORKERS when that is set (the drivers set it from `--cores N`, N > 1: the reference's pool size)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    n = min(n, limit)
    env = os.environ.get("NB200_HOST_WORKERS")
    if env:
        n = min(n, max(1, int(env)))
    return max(1, n)


def bgzip_tabix(path_plain, path_gz, threads=None, level=None):
    """BGZF-compress a sorted BED / bedgraph and write its .tbi in one native pass (nb200_bgzip_tabix_level).  `level`: deflate
    level, default -1 = zlib's default, which is what pysam.tabix_compress writes with (byte-identical files); the environment
    variable NB200_GZ_LEVEL (1..9) sets it for a whole run -- level 1 is about five times faster on bedgraph text for ~7 %
    larger files, same rows, same index."""
    import ctypes as C
    from . import _lib
    if level is None:
        level = int(os.environ.get("NB200_GZ_LEVEL", "-1"))
    err = C.create_string_buffer(256)
    st = _lib.load().nb200_bgzip_tabix_level(path_plain.encode(), path_gz.encode(), int(threads or host_workers()), int(level), err, 256)
    if st != 0:
        raise IOError("bgzip/tabix of %s failed: %s" % (path_plain, err.value.decode()))
    return path_gz


# ----------------------------------------------------------------------------------------- BAM
class BamFile:
    """Coordinate-sorted BAM reader.  `fetch_fragments(chrom, start, end)` returns the int32 arrays
    (pos, tlen) of the reads overlapping [start, end) that are proper-pair and forward."""

    def __init__(self, path):
        self.path = path
        self.fh = open(path, "rb")
        data, size = _read_block(self.fh, 0)
        buf, coff = data, size
        while len(buf) < 12 or len(buf) < 12 + struct.unpack_from("<i", buf, 4)[0] + 4:
            d, s = _read_block(self.fh, coff)
            buf += d
            coff += s
        if buf[:4] != b"BAM\x01":
            raise ValueError("not a BAM file: %s" % path)
        l_text = struct.unpack_from("<i", buf, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from("<i", buf, off)[0]
        off += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            while len(buf) < off + 4:
                d, s = _read_block(self.fh, coff)
                buf += d
                coff += s
            l_name = struct.unpack_from("<i", buf, off)[0]
            while len(buf) < off + 8 + l_name:
                d, s = _read_block(self.fh, coff)
                buf += d
                coff += s
            self.references.append(buf[off + 4:off + 4 + l_name - 1].decode())
            self.lengths.append(struct.unpack_from("<i", buf, off + 4 + l_name)[0])
            off += 8 + l_name
        self._tid = {n: i for i, n in enumerate(self.references)}
        self._index = None
        self._all = None
        bai = path + ".bai" if os.path.exists(path + ".bai") else path[:-4] + ".bai"
        if os.path.exists(bai):
            self._index = self._read_bai(bai)

    def close(self):
        self.fh.close()

    @staticmethod
    def _read_bai(path):
        with open(path, "rb") as fh:
            raw = fh.read()
        if raw[:4] != b"BAI\x01":
            raise ValueError("not a BAI index: %s" % path)
        n_ref = struct.unpack_from("<i", raw, 4)[0]
        off = 8
        linear = []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, off)[0]
            off += 4
            for _b in range(n_bin):
                _bin, n_chunk = struct.unpack_from("<Ii", raw, off)
                off += 8 + 16 * n_chunk
            n_intv = struct.unpack_from("<i", raw, off)[0]
            off += 4
            linear.append(np.frombuffer(raw, dtype="<u8", count=n_intv, offset=off).copy())
            off += 8 * n_intv
        return linear

    def _records(self, voffset):
        """Yield (tid, pos, flag, tlen, end) from virtual offset `voffset` onwards."""
        coff, uoff = voffset >> 16, voffset & 0xffff
        data, size = _read_block(self.fh, coff)
        buf = data[uoff:]
        coff += size
        p = 0
        while True:
            while len(buf) - p < 4:
                d, s = _read_block(self.fh, coff)
                if s == 0:
                    return
                buf = buf[p:] + d
                p = 0
                coff += s
            bs = struct.unpack_from("<i", buf, p)[0]
            while len(buf) - p < 4 + bs:
                d, s = _read_block(self.fh, coff)
                if s == 0:
                    return
                buf = buf[p:] + d
                p = 0
                coff += s
            tid, pos, _l_rn, _mq, _bin, _n_cig, flag, l_seq, _nt, _np, tlen = struct.unpack_from("<iiBBHHHiiii", buf, p + 4)
            # end of the alignment approximated by the read length (a superset of pysam's overlap test is
            # enough: the device re-checks every cell bound like fragments.pyx:37)
            yield tid, pos, flag, tlen, pos + max(l_seq, 1) + 64
            p += 4 + bs

    def _start_voffset(self, tid, start):
        """BGZF virtual offset to scan a region of reference `tid` from: the linear-index entry of the 16 kb window holding
        `start` (the previous non-empty one if that window has no reads); None when the reference has no reads."""
        lin = self._index[tid]
        if len(lin) == 0:
            return None
        w = min(max(start, 0) >> 14, len(lin) - 1)
        voff = int(lin[w])
        if voff == 0:  # window without reads: walk back to the previous non-empty one
            nz = np.nonzero(lin[:w + 1])[0]
            if len(nz) == 0:
                nz2 = np.nonzero(lin)[0]
                if len(nz2) == 0:
                    return None
                voff = int(lin[nz2[0]])
            else:
                voff = int(lin[nz[-1]])
        return voff

    def _fetch_indexed(self, tid, start, end):
        voff = self._start_voffset(tid, start)
        if voff is None:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        ps, ts = [], []
        for rtid, pos, flag, tlen, rend in self._records(voff):
            if rtid != tid or pos >= end:
                break
            if rend > start and (flag & 0x2) and not (flag & 0x10):  # fragments.pyx:25
                ps.append(pos)
                ts.append(tlen)
        return np.asarray(ps, dtype=np.int32), np.asarray(ts, dtype=np.int32)

    def _load_all(self):
        per = {i: ([], []) for i in range(len(self.references))}
        # first alignment follows the header: re-scan from the start and skip the header bytes
        with gzip.open(self.path, "rb") as fh:
            raw = fh.read()
        l_text = struct.unpack_from("<i", raw, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from("<i", raw, off)[0]
        off += 4
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", raw, off)[0]
            off += 8 + l_name
        n = len(raw)
        while off < n:
            bs = struct.unpack_from("<i", raw, off)[0]
            tid, pos, _l, _m, _b, _c, flag, _ls, _nt, _np, tlen = struct.unpack_from("<iiBBHHHiiii", raw, off + 4)
            off += 4 + bs
            if tid >= 0 and (flag & 0x2) and not (flag & 0x10):
                per[tid][0].append(pos)
                per[tid][1].append(tlen)
        self._all = {i: (np.asarray(p, dtype=np.int32), np.asarray(t, dtype=np.int32)) for i, (p, t) in per.items()}

    def fetch_fragments_many(self, regions, threads=None):
        """(frag_off int64[n+1], pos int32[], tlen int32[]) of the proper-pair forward reads of every (chrom, start, end)
        in `regions`, CSR-packed in region order.  With a .bai index the decode is native and multi-threaded
        (nb200_bam_fetch_many); without one it falls back to the per-region reader."""
        n = len(regions)
        if self._index is None or n == 0:
            ps, ts, off = [], [], [0]
            for chrom, start, end in regions:
                p, t = self.fetch_fragments(chrom, start, end)
                ps.append(p)
                ts.append(t)
                off.append(off[-1] + len(p))
            cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int32)
            return np.asarray(off, dtype=np.int64), cat(ps), cat(ts)
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        voff, tids = np.zeros(n, dtype=np.uint64), np.full(n, -1, dtype=np.int32)
        starts, ends = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        for i, (chrom, start, end) in enumerate(regions):
            tid = self._tid.get(chrom)
            v = None if tid is None else self._start_voffset(tid, max(0, start))
            if v is None:
                continue   # unknown reference / reference without reads: voffset 0, tid -1 -> skipped by the library
            voff[i], tids[i], starts[i], ends[i] = v, tid, max(0, start), end
        if threads is None:
            threads = max(1, min(host_workers(), n))
        off = np.zeros(n + 1, dtype=np.int64)
        pp, tp = _lib.c_int32_p(), _lib.c_int32_p()
        err = C.create_string_buffer(512)
        st = lib.nb200_bam_fetch_many(self.path.encode(), n, voff.ctypes.data_as(C.POINTER(C.c_uint64)), _lib.ptr(tids, C.c_int32),
                                      _lib.ptr(starts, C.c_int32), _lib.ptr(ends, C.c_int32), int(threads), _lib.ptr(off, C.c_int64),
                                      C.byref(pp), C.byref(tp), err, 512)
        if st != 0:
            raise IOError(err.value.decode() or "nb200_bam_fetch_many failed")
        total = int(off[-1])
        try:
            pos = np.ctypeslib.as_array(pp, shape=(max(total, 1),))[:total].copy()
            tlen = np.ctypeslib.as_array(tp, shape=(max(total, 1),))[:total].copy()
        finally:
            lib.nb200_free(pp)
            lib.nb200_free(tp)
        return off, pos, tlen

    def fetch_fragments(self, chrom, start, end):
        if chrom not in self._tid:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        tid = self._tid[chrom]
        if self._index is not None:   # native reader; _fetch_indexed is its pure-Python twin (kept as the cross-check)
            _off, pos, tlen = self.fetch_fragments_many([(chrom, max(0, start), end)], threads=1)
            return pos, tlen
        if self._all is None:
            self._load_all()
        pos, tlen = self._all[tid]
        # without an index: a superset by position (the device re-checks every cell bound, fragments.pyx:37)
        sel = (pos >= start - 5000) & (pos < end)
        return pos[sel], tlen[sel]


def index_bam(path, out=None):
    """Write a .bai for a coordinate-sorted BAM (`samtools index` for the needs of this package): the 16 kb linear index
    only -- for every window the virtual offset of the first record that overlaps it -- and no bins, which is what
    BamFile's region fetch uses.  Plain Python: meant for small files and the tests."""
    bam = BamFile.__new__(BamFile)
    BamFile.__init__(bam, path)
    fh = bam.fh
    # position of the first record = end of the header, found by re-walking it block by block
    coff, ubase = 0, 0          # compressed offset of the current block, and of the block the buffer starts in
    data, size = _read_block(fh, 0)
    buf, next_coff = bytearray(data), size
    starts = [(0, 0, len(data))]  # (offset in buf, coff, length) of every block in buf

    def grow():
        nonlocal next_coff
        d, sz = _read_block(fh, next_coff)
        if sz == 0:
            return False
        starts.append((len(buf), next_coff, len(d)))
        buf.extend(d)
        next_coff += sz
        return True

    def voffset(p):
        for b0, c0, ln in reversed(starts):
            if p >= b0 and (p < b0 + ln or ln == 0):
                return (c0 << 16) | (p - b0)
        b0, c0, ln = starts[-1]
        return (next_coff << 16) if p == b0 + ln else None

    def need(p, n):
        while len(buf) - p < n:
            if not grow():
                return False
        return True
    need(0, 12)
    l_text = struct.unpack_from("<i", buf, 4)[0]
    p = 8 + l_text
    need(p, 4)
    n_ref = struct.unpack_from("<i", buf, p)[0]
    p += 4
    for _ in range(n_ref):
        need(p, 4)
        l_name = struct.unpack_from("<i", buf, p)[0]
        need(p, 8 + l_name)
        p += 8 + l_name
    linear = [dict() for _ in range(n_ref)]
    while need(p, 4):
        bs = struct.unpack_from("<i", buf, p)[0]
        if not need(p, 4 + bs):
            break
        tid, pos, _l, _m, _b, _c, _flag, l_seq = struct.unpack_from("<iiBBHHHi", buf, p + 4)
        if tid >= 0:
            v = voffset(p)
            for w in range(max(pos, 0) >> 14, (max(pos, 0) + max(l_seq, 1) - 1 >> 14) + 1):
                if w not in linear[tid]:
                    linear[tid][w] = v
        p += 4 + bs
    bam.close()
    out = out or path + ".bai"
    with open(out, "wb") as o:
        o.write(b"BAI\x01" + struct.pack("<i", n_ref))
        for t in range(n_ref):
            n_intv = (max(linear[t]) + 1) if linear[t] else 0
            o.write(struct.pack("<ii", 0, n_intv))
            o.write(np.asarray([linear[t].get(w, 0) for w in range(n_intv)], dtype="<u8").tobytes())
        o.write(struct.pack("<Q", 0))
    return out


# ----------------------------------------------------------------------------------------- FASTA
class FastaFile:
    """.fai-indexed FASTA fetch (pysam.FastaFile stand-in): `.references`, `.lengths`, `.fetch`."""

    def __init__(self, path):
        self.path = path
        self.index = {}
        if not os.path.exists(path + ".fai"):
            self._build_index()
        with open(path + ".fai") as fh:
            for line in fh:
                name, length, offset, linebases, linewidth = line.rstrip("\n").split("\t")[:5]
                self.index[name] = (int(length), int(offset), int(linebases), int(linewidth))
        self.references = list(self.index.keys())
        self.lengths = [self.index[k][0] for k in self.references]
        self.fh = open(path, "rb")

    def _build_index(self):
        entries, name, length, offset, lb, lw, pos = [], None, 0, 0, 0, 0, 0
        with open(self.path, "rb") as fh:
            for line in fh:
                if line.startswith(b">"):
                    if name is not None:
                        entries.append((name, length, offset, lb, lw))
                    name, length, lb, lw = line[1:].split()[0].decode(), 0, 0, 0
                    offset = pos + len(line)
                else:
                    if lb == 0:
                        lb, lw = len(line.rstrip(b"\r\n")), len(line)
                    length += len(line.rstrip(b"\r\n"))
                pos += len(line)
        if name is not None:
            entries.append((name, length, offset, lb, lw))
        with open(self.path + ".fai", "w") as out:
            for e in entries:
                out.write("\t".join(str(x) for x in e) + "\n")

    def fetch(self, chrom, start, end):
        length, offset, linebases, linewidth = self.index[chrom]
        start, end = max(0, start), min(length, end)
        if end <= start:
            return ""
        b0 = offset + (start // linebases) * linewidth + start % linebases
        b1 = offset + ((end - 1) // linebases) * linewidth + (end - 1) % linebases + 1
        self.fh.seek(b0)
        return self.fh.read(b1 - b0).replace(b"\n", b"").replace(b"\r", b"").decode()

    def close(self):
        self.fh.close()


# ----------------------------------------------------------------------------------------- bedgraph
class BedGraphReader:
    """bgzip'd (or plain) bedgraph; region queries by scanning per-chromosome row tables read once."""

    def __init__(self, path):
        opener = gzip.open if path.endswith(".gz") else open
        self.rows = {}
        with opener(path, "rt") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if len(f) >= 4:
                    self.rows.setdefault(f[0], []).append((int(f[1]), int(f[2]), float(f[3])))
        self.tables = {c: (np.array([r[0] for r in v]), np.array([r[1] for r in v]), np.array([r[2] for r in v]))
                       for c, v in self.rows.items()}

    def read(self, chrom, start, end, empty=np.nan):
        out = np.ones(end - start) * empty  # pyatac/bedgraph.py:10
        if chrom not in self.tables:
            return out
        s, e, v = self.tables[chrom]
        lo, hi = np.searchsorted(e, start, side="right"), np.searchsorted(s, end, side="left")
        for i in range(lo, hi):
            out[max(s[i] - start, 0):min(e[i] - start, end - start)] = v[i]
        return out


# ----------------------------------------------------------------------------------------- tabix (.tbi)
def _reg2bin(beg, end):
    """UCSC binning scheme of tabix/BAI (min_shift 14, depth 5)."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_lines(path):
    """Yield (virtual offset of line start, virtual offset after the line, line bytes) of a BGZF text file."""
    with open(path, "rb") as fh:
        coff = 0
        pending, pending_start = b"", None
        while True:
            data, size = _read_block(fh, coff)
            if size == 0:
                break
            pos = 0
            while True:
                nl = data.find(b"\n", pos)
                if nl < 0:
                    if pos < len(data):
                        if pending_start is None:
                            pending_start = (coff << 16) | pos
                        pending += data[pos:]
                    break
                start = pending_start if pending_start is not None else (coff << 16) | pos
                line = pending + data[pos:nl]
                pending, pending_start = b"", None
                end_u = nl + 1
                end_v = ((coff << 16) | end_u) if end_u < len(data) else ((coff + size) << 16)
                yield start, end_v, line
                pos = nl + 1
            coff += size


def tabix_index(path_gz, seq_col=1, beg_col=2, end_col=3, zero_based=True):
    """Write `path_gz + '.tbi'` for a coordinate-sorted, BGZF-compressed BED / bedgraph -- what
    `pysam.tabix_index(..., preset='bed')` produces for the reference's outputs (run_occ.py:130-136)."""
    names, bins, linear = [], [], []
    tid = {}
    for start_v, end_v, line in _bgzf_lines(path_gz):
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t")
        chrom = f[seq_col - 1].decode()
        beg = int(f[beg_col - 1]) - (0 if zero_based else 1)
        end = int(f[end_col - 1])
        if end <= beg:
            end = beg + 1
        if chrom not in tid:
            tid[chrom] = len(names)
            names.append(chrom)
            bins.append({})
            linear.append([])
        t = tid[chrom]
        b = _reg2bin(beg, end)
        chunks = bins[t].setdefault(b, [])
        if chunks and chunks[-1][1] == start_v:
            chunks[-1][1] = end_v            # records are consecutive in the file: extend the chunk
        else:
            chunks.append([start_v, end_v])
        lin = linear[t]
        w0, w1 = beg >> 14, (end - 1) >> 14
        if len(lin) <= w1:
            lin.extend([-1] * (w1 + 1 - len(lin)))
        for w in range(w0, w1 + 1):
            if lin[w] < 0:
                lin[w] = start_v
    out = bytearray()
    nm = b"".join(n.encode() + b"\x00" for n in names)
    out += b"TBI\x01" + struct.pack("<iiiiiii", len(names), 0x10000 if zero_based else 0, seq_col, beg_col, end_col, ord("#"), 0)
    out += struct.pack("<i", len(nm)) + nm
    for t in range(len(names)):
        out += struct.pack("<i", len(bins[t]))
        for b in sorted(bins[t]):
            out += struct.pack("<Ii", b, len(bins[t][b]))
            for c0, c1 in bins[t][b]:
                out += struct.pack("<QQ", c0, c1)
        lin = linear[t]
        # windows without records: leading ones point at the first record, later ones repeat the previous window
        # (the layout of the .tbi files the reference shipped, written by pysam/htslib)
        first = next((v for v in lin if v >= 0), 0)
        prev = first
        for w in range(len(lin)):
            if lin[w] < 0:
                lin[w] = prev
            else:
                prev = lin[w]
        out += struct.pack("<i", len(lin)) + b"".join(struct.pack("<Q", v) for v in lin)
    w = BgzfWriter(path_gz + ".tbi")
    w.write(bytes(out))
    w.close()
    return path_gz + ".tbi"


class TabixFile:
    """Region queries on a BGZF file through its .tbi (pysam.Tabixfile stand-in for `--occ_track`)."""

    def __init__(self, path_gz):
        self.path = path_gz
        with gzip.open(path_gz + ".tbi", "rb") as fh:
            raw = fh.read()
        if raw[:4] != b"TBI\x01":
            raise ValueError("not a tabix index")
        n_ref, self.fmt, self.sc, self.bc, self.ec, _meta, _skip, l_nm = struct.unpack_from("<iiiiiiii", raw, 4)
        off = 36
        self.names = [x.decode() for x in raw[off:off + l_nm].split(b"\x00")[:-1]]
        off += l_nm
        self.linear, self.bins = [], []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, off)[0]
            off += 4
            bd = {}
            for _b in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", raw, off)
                off += 8
                bd[b] = [struct.unpack_from("<QQ", raw, off + 16 * i) for i in range(n_chunk)]
                off += 16 * n_chunk
            n_intv = struct.unpack_from("<i", raw, off)[0]
            off += 4
            self.linear.append(list(struct.unpack_from("<%dQ" % n_intv, raw, off)))
            off += 8 * n_intv
            self.bins.append(bd)
        self.fh = open(path_gz, "rb")

    @property
    def contigs(self):
        return self.names

    def fetch(self, chrom, start, end):
        """Lines (split on tabs) overlapping [start, end)."""
        if chrom not in self.names:
            return []
        t = self.names.index(chrom)
        lin = self.linear[t]
        if not lin:
            return []
        voff = lin[min(max(start, 0) >> 14, len(lin) - 1)]   # a start left of the chromosome begins at its first window
        coff, uoff = voff >> 16, voff & 0xffff
        out, buf = [], b""
        data, size = _read_block(self.fh, coff)
        buf = data[uoff:]
        coff += size
        while True:
            nl = buf.find(b"\n")
            if nl < 0:
                data, size = _read_block(self.fh, coff)
                if size == 0:
                    break
                buf += data
                coff += size
                continue
            f = buf[:nl].split(b"\t")
            buf = buf[nl + 1:]
            if len(f) < 3 or f[0].startswith(b"#"):
                continue
            if f[self.sc - 1].decode() != chrom:
                if out:
                    break
                continue
            b, e = int(f[self.bc - 1]), int(f[self.ec - 1])
            if b >= end:
                break
            if e > start:
                out.append([x.decode() for x in f])
        return out

    def close(self):
        self.fh.close()

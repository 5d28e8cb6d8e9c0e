"""Host-side driver of libnucleo_b200: one `Engine` per (process, GPU).

It owns the C context, uploads the run constants (PWM, VMat, fragment-size
distribution, occupancy model, smoothing windows, peak jitter) and runs batches of
chunks through the device occ / nuc paths.  The reference-named classes
(`OccChunk`, `NucChunk`, `FragmentMat2D`, ...) in this package are thin layers over
an Engine; `run_occ` / `run_nuc` feed it round-robin shards of the BED chunk list.
"""
import ctypes as C

import numpy as np

from . import _lib as L

FLAG_COV, FLAG_LR, FLAG_Z, FLAG_NONREDUNDANT = 1, 2, 4, 8


def gaussian_window(M, std):
    """scipy.signal.gaussian(M, std) (sym=True): exp(-n^2 / (2 std^2)), n centred."""
    n = np.arange(0, M) - (M - 1.0) / 2.0
    sig2 = 2 * std * std
    return np.exp(-n ** 2 / sig2)


def peak_jitter(n):
    """RandomState(25).uniform(0, 1e-12, n) of pyatac/utils.py:94-97 (a prefix of any longer draw)."""
    return np.random.RandomState(seed=25).uniform(0, 10 ** -12, int(n))


class PackedBatch:
    """Host arrays of a batch of chunks in the layout nb200_batch_upload takes."""

    def __init__(self, starts, ends, frag_off, frag_pos, frag_tlen, seq_off=None, seq_start=None, seq=None):
        self.starts = L.as_i32(starts)
        self.ends = L.as_i32(ends)
        self.frag_off = L.as_i64(frag_off)
        self.frag_pos = L.as_i32(frag_pos)
        self.frag_tlen = L.as_i32(frag_tlen)
        self.seq_off = None if seq_off is None else L.as_i64(seq_off)
        self.seq_start = None if seq_start is None else L.as_i32(seq_start)
        self.seq = None if seq is None else np.ascontiguousarray(seq, dtype=np.uint8)
        self.n = len(self.starts)
        self.lengths = (self.ends - self.starts).astype(np.int64)
        self.out_off = np.concatenate(([0], np.cumsum(self.lengths))).astype(np.int64)
        self.total_len = int(self.out_off[-1])

    @staticmethod
    def from_chunks(chunks):
        """chunks: iterable of dicts/tuples (start, end, pos[], tlen[], seq bytes or None, seq_start)."""
        starts, ends, fo, ps, ts, so, ss, sq = [], [], [0], [], [], [0], [], []
        have_seq = True
        for (s, e, pos, tlen, seq, seq_start) in chunks:
            starts.append(s)
            ends.append(e)
            ps.append(np.asarray(pos, dtype=np.int32))
            ts.append(np.asarray(tlen, dtype=np.int32))
            fo.append(fo[-1] + len(pos))
            if seq is None:
                have_seq = False
            else:
                arr = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
                sq.append(arr)
                so.append(so[-1] + len(arr))
                ss.append(seq_start)
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)
        if have_seq:
            return PackedBatch(starts, ends, fo, cat(ps, np.int32), cat(ts, np.int32), so, ss, cat(sq, np.uint8))
        return PackedBatch(starts, ends, fo, cat(ps, np.int32), cat(ts, np.int32))


class Engine:
    def __init__(self, device=0):
        self.lib = L.load()
        h = C.c_void_p()
        st = self.lib.nb200_ctx_create(int(device), C.byref(h))
        if st != 0:
            raise L.NB200Error(st, self.lib.nb200_last_error(None).decode())
        self.h = h
        self.device = device
        self._jitter_n = 0
        self.occ_params = None
        self.nuc_params = None
        self._keep = {}

    # ------------------------------------------------------------------ plumbing
    def check(self, st):
        if st != 0:
            raise L.NB200Error(st, self.lib.nb200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.nb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, ma, mi, hbm = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self.check(self.lib.nb200_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hbm)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), hbm_bytes=hbm.value)

    def pinned(self, shape, dtype):
        """numpy array on page-locked host memory (cudaHostAlloc) for the end-to-end path."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self.check(self.lib.nb200_host_alloc(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._keep[arr.ctypes.data] = p
        return arr

    # ------------------------------------------------------------------ run constants
    def set_pwm(self, mat, up, down, nucleotides):
        """PWM.open + log, pyatac/bias.py:47-76,90."""
        logm = L.as_f64(np.log(np.asarray(mat, dtype=np.float64)))
        nucs = "".join(nucleotides)
        self.check(self.lib.nb200_set_pwm(self.h, L.ptr(logm, C.c_double), logm.shape[0], int(up), int(down), nucs.encode()))
        self.pwm_up, self.pwm_down = int(up), int(down)

    def set_vmat(self, mat, lower, upper):
        mat = L.as_f64(mat)
        if mat.ndim != 2 or mat.shape[0] != upper - lower:
            raise Exception("mat shape is not consistent with insert limits")  # VMat.py:33-34
        self.check(self.lib.nb200_set_vmat(self.h, L.ptr(mat, C.c_double), mat.shape[0], mat.shape[1], int(lower), int(upper)))
        self.vmat_shape = mat.shape
        self.vmat_lower, self.vmat_upper = int(lower), int(upper)

    def set_fragment_sizes(self, freq):
        freq = L.as_f64(freq)
        self.check(self.lib.nb200_set_fragment_sizes(self.h, L.ptr(freq, C.c_double), len(freq)))

    def set_occ_model(self, nuc_probs, nfr_probs, alphas=None, cutoff=None, ci=0.9):
        """OccupancyCalcParams, nucleoatac/Occupancy.py:89-102 (probabilities already normalised)."""
        nuc_probs, nfr_probs = L.as_f64(nuc_probs), L.as_f64(nfr_probs)
        if alphas is None:
            alphas = np.linspace(0, 1, 101)
        alphas = L.as_f64(alphas)
        if cutoff is None:
            from scipy import stats
            cutoff = stats.chi2.ppf(ci, 1)
        self.check(self.lib.nb200_set_occ_model(self.h, L.ptr(nuc_probs, C.c_double), L.ptr(nfr_probs, C.c_double),
                                                len(nuc_probs), L.ptr(alphas, C.c_double), len(alphas), float(cutoff)))

    def ensure_jitter(self, n):
        if n > self._jitter_n:
            n = max(int(n), 16384)
            j = L.as_f64(peak_jitter(n))
            self.check(self.lib.nb200_set_jitter(self.h, L.ptr(j, C.c_double), n))
            self._jitter_n = n

    def configure_occ(self, upper=251, flank=60, step=5, sep=120, min_occ=0.1, atac=True, use_bias=True):
        """OccupancyParameters, nucleoatac/Occupancy.py:175-193."""
        window = 2 * flank + 1
        win = L.as_f64(gaussian_window(window, flank / 3.0))  # Occupancy.py:220
        p = L.OccParams(int(upper), int(flank), int(step), int(sep), float(min_occ), int(bool(atac)), int(bool(use_bias)),
                        L.ptr(win, C.c_double), len(win))
        self.check(self.lib.nb200_occ_configure(self.h, C.byref(p)))
        self.occ_params = dict(upper=upper, flank=flank, step=step, sep=sep, min_occ=min_occ, atac=atac, use_bias=use_bias)

    def configure_nuc(self, sd=10, nonredundant_sep=120, redundant_sep=25, min_z=3, min_lr=0, min_reads=1, atac=True,
                      use_bias=True, xcor_mode=0):
        """NucParameters, nucleoatac/NucleosomeCalling.py:204-226."""
        win = L.as_f64(gaussian_window(6 * sd + 1, sd))  # :276-283
        p = L.NucParams(int(bool(atac)), int(bool(use_bias)), int(sd), int(nonredundant_sep), int(redundant_sep),
                        float(min_z), float(min_lr), float(min_reads), L.ptr(win, C.c_double), len(win), int(xcor_mode))
        self.check(self.lib.nb200_nuc_configure(self.h, C.byref(p)))
        self.nuc_params = dict(sd=sd, nonredundant_sep=nonredundant_sep, redundant_sep=redundant_sep, min_z=min_z,
                               min_lr=min_lr, min_reads=min_reads, atac=atac, use_bias=use_bias, xcor_mode=xcor_mode)

    # ------------------------------------------------------------------ batches
    def upload(self, pb, handle=None):
        """nb200_batch_upload; returns the device batch handle (recycles `handle`'s buffers)."""
        self.ensure_jitter(int(pb.lengths.max()))
        hb = L.Batch(pb.n, L.ptr(pb.starts, C.c_int32), L.ptr(pb.ends, C.c_int32), L.ptr(pb.frag_off, C.c_int64),
                     L.ptr(pb.frag_pos, C.c_int32), L.ptr(pb.frag_tlen, C.c_int32), L.ptr(pb.seq_off, C.c_int64),
                     L.ptr(pb.seq_start, C.c_int32), L.ptr(pb.seq, C.c_uint8))
        h = handle if handle is not None else C.c_void_p()
        self.check(self.lib.nb200_batch_upload(self.h, C.byref(hb), C.byref(h)))
        return h

    def free_batch(self, h):
        self.check(self.lib.nb200_batch_free(self.h, h))

    def sync(self, h):
        self.check(self.lib.nb200_batch_sync(self.h, h))

    def occ_run(self, h):
        self.check(self.lib.nb200_occ_run(self.h, h))

    def nuc_run(self, h):
        self.check(self.lib.nb200_nuc_run(self.h, h))

    def occ_capacity(self, pb):
        sep = self.occ_params["sep"]
        return np.concatenate(([0], np.cumsum(pb.lengths // sep + 2))).astype(np.int64)

    def nuc_capacity(self, pb):
        sep = self.nuc_params["redundant_sep"]
        return np.concatenate(([0], np.cumsum(pb.lengths // sep + 2))).astype(np.int64)

    def occ_alloc(self, pb, raw=True, alloc=None, track_dtype=np.float64):
        """Host result buffers for nb200_occ_download (pinned when alloc=self.pinned); track_dtype=np.float32 makes the
        per-position tracks float32 (nb200_occ_download32: converted on the device, half the bytes on the host link)."""
        alloc = alloc or (lambda shape, dt: np.empty(shape, dtype=dt))
        n, tl, up = pb.n, pb.total_len, self.occ_params["upper"]
        po = self.occ_capacity(pb)
        td = np.dtype(track_dtype)
        out = dict(smoothed_vals=alloc(tl, td), smoothed_lower=alloc(tl, td),
                   smoothed_upper=alloc(tl, td), cov=alloc(tl, td), nuc_dist=alloc((n, up), np.float64),
                   peak_count=alloc(n, np.int32), peak_off=po, peak_pos=alloc(int(po[-1]), np.int32),
                   peak_occ=alloc(int(po[-1]), np.float64), peak_lower=alloc(int(po[-1]), np.float64),
                   peak_upper=alloc(int(po[-1]), np.float64), peak_reads=alloc(int(po[-1]), np.float64))
        if raw:
            out.update(vals=alloc(tl, td), lower_bound=alloc(tl, td), upper_bound=alloc(tl, td))
        return out

    def nuc_alloc(self, pb, cov=True, alloc=None, track_dtype=np.float64):
        alloc = alloc or (lambda shape, dt: np.empty(shape, dtype=dt))
        n, tl = pb.n, pb.total_len
        co = self.nuc_capacity(pb)
        nc = int(co[-1])
        td = np.dtype(track_dtype)
        out = dict(nuc_signal=alloc(tl, td), background=alloc(tl, td), norm_signal=alloc(tl, td),
                   smoothed=alloc(tl, td), cand_count=alloc(n, np.int32), cand_off=co,
                   cand_pos=alloc(nc, np.int32), cand_flag=alloc(nc, np.int32), cand_z=alloc(nc, np.float64),
                   cand_lr=alloc(nc, np.float64), cand_norm_signal=alloc(nc, np.float64),
                   cand_nuc_signal=alloc(nc, np.float64), cand_nuc_cov=alloc(nc, np.float64),
                   cand_nfr_cov=alloc(nc, np.float64), cand_smoothed=alloc(nc, np.float64))
        if cov:
            out.update(nuc_cov=alloc(tl, td), nfr_cov=alloc(tl, td))
        return out

    @staticmethod
    def _fill(struct, out):
        for name, ctype in struct._fields_:
            a = out.get(name)
            if a is not None:
                setattr(struct, name, a.ctypes.data_as(ctype))
        return struct

    @staticmethod
    def _track_dtype(out, names):
        """float64 or float32: the dtype of the track buffers in `out` (they must agree; it picks the C entry point)."""
        kinds = {out[n].dtype for n in names if out.get(n) is not None}
        if len(kinds) > 1 or (kinds and next(iter(kinds)) not in (np.dtype(np.float64), np.dtype(np.float32))):
            raise TypeError("track buffers must be all float64 or all float32, got %s" % sorted(str(k) for k in kinds))
        return next(iter(kinds)) if kinds else np.dtype(np.float64)

    OCC_TRACKS = ("smoothed_vals", "smoothed_lower", "smoothed_upper", "vals", "lower_bound", "upper_bound", "cov")
    NUC_TRACKS = ("nuc_signal", "background", "norm_signal", "smoothed", "nuc_cov", "nfr_cov")

    def occ_download(self, h, out):
        """Enqueue the D2H of the occ results into `out` (float64 tracks: nb200_occ_download; float32 tracks:
        nb200_occ_download32, converted on the device).  Returns the bytes that cross the link."""
        if self._track_dtype(out, self.OCC_TRACKS) == np.float32:
            o = self._fill(L.OccOut32(), out)
            self.check(self.lib.nb200_occ_download32(self.h, h, C.byref(o)))
            return int(self.lib.nb200_occ_d2h_bytes32(h, C.byref(o)))
        o = self._fill(L.OccOut(), out)
        self.check(self.lib.nb200_occ_download(self.h, h, C.byref(o)))
        return int(self.lib.nb200_occ_d2h_bytes(h, C.byref(o)))

    def nuc_download(self, h, out):
        if self._track_dtype(out, self.NUC_TRACKS) == np.float32:
            o = self._fill(L.NucOut32(), out)
            self.check(self.lib.nb200_nuc_download32(self.h, h, C.byref(o)))
            return int(self.lib.nb200_nuc_d2h_bytes32(h, C.byref(o)))
        o = self._fill(L.NucOut(), out)
        self.check(self.lib.nb200_nuc_download(self.h, h, C.byref(o)))
        return int(self.lib.nb200_nuc_d2h_bytes(h, C.byref(o)))

    def h2d_bytes(self, h):
        return int(self.lib.nb200_batch_h2d_bytes(h))

    # convenience: whole batch, synchronous
    def process_occ(self, pb, raw=True):
        h = self.upload(pb)
        try:
            self.occ_run(h)
            out = self.occ_alloc(pb, raw=raw)
            self.occ_download(h, out)
            self.sync(h)
        finally:
            self.free_batch(h)
        return out

    def process_nuc(self, pb):
        h = self.upload(pb)
        try:
            self.nuc_run(h)
            out = self.nuc_alloc(pb)
            self.nuc_download(h, out)
            self.sync(h)
        finally:
            self.free_batch(h)
        return out

    # ------------------------------------------------------------------ timing / profiling
    def timer_start(self, h):
        self.check(self.lib.nb200_timer_start(self.h, h))

    def timer_stop(self, h):
        self.check(self.lib.nb200_timer_stop(self.h, h))

    def timer_ms(self, h):
        ms = C.c_float()
        self.check(self.lib.nb200_timer_elapsed_ms(self.h, h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self, h=None):
        self.check(self.lib.nb200_flush_l2(self.h, h))

    def profile(self, on=True):
        self.check(self.lib.nb200_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self.check(self.lib.nb200_profile_reset(self.h))

    def profile_report(self):
        n = self.lib.nb200_profile_count(self.h)
        rep = {}
        for i in range(n):
            name, cnt, ms = C.c_char_p(), C.c_int64(), C.c_double()
            self.check(self.lib.nb200_profile_get(self.h, i, C.byref(name), C.byref(cnt), C.byref(ms)))
            rep[name.value.decode()] = (int(cnt.value), float(ms.value))
        return rep

    # ------------------------------------------------------------------ end-of-run reductions (NCCL inside the library)
    def nccl_unique_id(self):
        buf = (C.c_ubyte * 128)()
        st = self.lib.nb200_nccl_unique_id(buf)
        if st != 0:
            raise L.NB200Error(st, self.lib.nb200_last_error(None).decode())
        return bytes(buf)

    def nccl_init(self, unique_id, rank, world):
        raw = (C.c_ubyte * 128)(*unique_id)
        self.check(self.lib.nb200_nccl_init(self.h, raw, int(rank), int(world)))

    def allreduce(self, arr):
        """In-place sum over the ranks of a float64 or int64 numpy array (fragment sizes, nuc_dist, V-plot sums)."""
        assert arr.flags["C_CONTIGUOUS"]
        if arr.dtype == np.float64:
            self.check(self.lib.nb200_allreduce_f64(self.h, L.ptr(arr, C.c_double), arr.size))
        elif arr.dtype == np.int64:
            self.check(self.lib.nb200_allreduce_i64(self.h, L.ptr(arr, C.c_int64), arr.size))
        else:
            raise TypeError("allreduce takes float64 or int64 arrays")
        return arr

    def nccl_finalize(self):
        self.check(self.lib.nb200_nccl_finalize(self.h))

    # ------------------------------------------------------------------ primitives (reference seams)
    def fragmat(self, pos, tlen, start, end, lower, upper, atac=True):
        pos, tlen = L.as_i32(pos), L.as_i32(tlen)
        out = np.empty((upper - lower, end - start), dtype=np.float64)
        self.check(self.lib.nb200_fragmat_build(self.h, L.ptr(pos, C.c_int32), L.ptr(tlen, C.c_int32), len(pos), start, end,
                                                lower, upper, int(bool(atac)), L.ptr(out, C.c_double)))
        return out

    def insertions(self, pos, tlen, start, end, lower, upper, atac=True):
        pos, tlen = L.as_i32(pos), L.as_i32(tlen)
        out = np.empty(end - start, dtype=np.float64)
        self.check(self.lib.nb200_insertions(self.h, L.ptr(pos, C.c_int32), L.ptr(tlen, C.c_int32), len(pos), start, end,
                                             lower, upper, int(bool(atac)), L.ptr(out, C.c_double)))
        return out

    def fragment_sizes(self, starts, ends, frag_off, pos, tlen, lower, upper, atac=True):
        starts, ends, frag_off = L.as_i32(starts), L.as_i32(ends), L.as_i64(frag_off)
        pos, tlen = L.as_i32(pos), L.as_i32(tlen)
        out = np.zeros(upper - lower, dtype=np.int64)
        self.check(self.lib.nb200_fragment_sizes(self.h, len(starts), L.ptr(starts, C.c_int32), L.ptr(ends, C.c_int32),
                                                 L.ptr(frag_off, C.c_int64), L.ptr(pos, C.c_int32), L.ptr(tlen, C.c_int32),
                                                 lower, upper, int(bool(atac)), L.ptr(out, C.c_int64)))
        return out

    def bias_track(self, seq):
        arr = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
        arr = np.ascontiguousarray(arr)
        out = np.empty(len(arr) - (self.pwm_up + self.pwm_down), dtype=np.float64)
        self.check(self.lib.nb200_bias_track(self.h, L.ptr(arr, C.c_uint8), len(arr), L.ptr(out, C.c_double)))
        return out

    def biasmat(self, bias_vals, lower, upper):
        b = L.as_f64(bias_vals)
        ncol = len(b) - (upper + (upper - 1) % 2) + 1
        if ncol < 1:
            raise Exception("Insufficient flanking region on bias track for the bias matrix")
        out = np.empty((upper - lower, ncol), dtype=np.float64)
        self.check(self.lib.nb200_biasmat_build(self.h, L.ptr(b, C.c_double), len(b), lower, upper, L.ptr(out, C.c_double)))
        return out

    def get_ins(self, mat, lower, upper):
        mat = L.as_f64(mat)
        n = mat.shape[1] - (upper + (upper - 1) % 2) + 1
        out = np.empty(n, dtype=np.float64)
        self.check(self.lib.nb200_get_ins(self.h, L.ptr(mat, C.c_double), lower, upper, mat.shape[1], L.ptr(out, C.c_double)))
        return out

    def xcor_dense(self, mat):
        mat = L.as_f64(mat)
        n = mat.shape[1] - self.vmat_shape[1] + 1
        if mat.shape[0] != self.vmat_shape[0]:
            raise Exception("mat rows do not match the VMat")
        if n < 1:
            raise Exception("Insufficient flanking region on mat to calculate signal")
        out = np.empty(n, dtype=np.float64)
        self.check(self.lib.nb200_xcor_dense(self.h, L.ptr(mat, C.c_double), mat.shape[1], L.ptr(out, C.c_double)))
        return out

    def coverage_dense(self, mat, row0, row1, window_len):
        mat = L.as_f64(mat)
        n = mat.shape[1] - window_len + 1
        if n < 1:
            raise Exception("Insufficient flanking region on mat to calculate coverage with desired window")
        out = np.empty(n, dtype=np.float64)
        self.check(self.lib.nb200_coverage_dense(self.h, L.ptr(mat, C.c_double), mat.shape[0], mat.shape[1], row0, row1,
                                                 window_len, L.ptr(out, C.c_double)))
        return out

    def coverage(self, pos, tlen, start, end, lower, upper, window_len, atac=True):
        """pyatac/get_cov.py:22-31 without the final scaling, straight from the reads (no dense matrix)."""
        pos, tlen = L.as_i32(pos), L.as_i32(tlen)
        out = np.empty(end - start, dtype=np.float64)
        self.check(self.lib.nb200_coverage(self.h, L.ptr(pos, C.c_int32), L.ptr(tlen, C.c_int32), len(pos), start, end, lower,
                                           upper, int(window_len), int(bool(atac)), L.ptr(out, C.c_double)))
        return out

    def vplot(self, centers, flips, frag_off, pos, tlen, flank, lower, upper, atac=True, scale=False):
        """Sum of the per-site V-plot matrices (pyatac/make_vplot.py:22-43): float64 [upper-lower, 2*flank+1]."""
        centers, flips, frag_off = L.as_i32(centers), L.as_i32(flips), L.as_i64(frag_off)
        pos, tlen = L.as_i32(pos), L.as_i32(tlen)
        out = np.empty((upper - lower, 2 * flank + 1), dtype=np.float64)
        self.check(self.lib.nb200_vplot(self.h, len(centers), L.ptr(centers, C.c_int32), L.ptr(flips, C.c_int32),
                                        L.ptr(frag_off, C.c_int64), L.ptr(pos, C.c_int32), L.ptr(tlen, C.c_int32), int(flank),
                                        int(lower), int(upper), int(bool(atac)), int(bool(scale)), L.ptr(out, C.c_double)))
        return out

    def smooth(self, sig, w, mode="same", norm=True):
        sig, w = L.as_f64(sig), L.as_f64(w)
        n = len(sig) if mode == "same" else len(sig) - len(w) + 1
        out = np.empty(n, dtype=np.float64)
        self.check(self.lib.nb200_smooth(self.h, L.ptr(sig, C.c_double), len(sig), L.ptr(w, C.c_double), len(w),
                                         int(mode == "same"), int(bool(norm)), L.ptr(out, C.c_double)))
        return out

    def call_peaks(self, sig, min_signal=0, sep=120, boundary=None, order=1):
        """pyatac/utils.py:82-102; `sig` (float64 array) is updated in place (NaN -> min) like the reference."""
        assert sig.dtype == np.float64 and sig.flags["C_CONTIGUOUS"]
        if boundary is None:
            boundary = sep // 2
        self.ensure_jitter(len(sig))
        cap = len(sig) // sep + 2
        idx = np.empty(cap, dtype=np.int32)
        n = C.c_int32()
        self.check(self.lib.nb200_call_peaks(self.h, L.ptr(sig, C.c_double), len(sig), float(min_signal), int(sep), int(boundary),
                                             int(order), L.ptr(idx, C.c_int32), cap, C.byref(n)))
        return idx[:n.value].astype(np.int64)

    def reduce_peaks(self, peaks, sig, sep):
        peaks, sig = L.as_i32(peaks), L.as_f64(sig)
        keep = np.zeros(len(peaks), dtype=np.int32)
        self.check(self.lib.nb200_reduce_peaks(self.h, L.ptr(peaks, C.c_int32), L.ptr(sig, C.c_double), len(peaks), int(sep),
                                               L.ptr(keep, C.c_int32)))
        return peaks[keep == 1].astype(np.int64)

    def calculate_occupancy(self, inserts, bias):
        inserts, bias = L.as_f64(inserts), L.as_f64(bias)
        out = np.empty(3, dtype=np.float64)
        self.check(self.lib.nb200_calculate_occupancy(self.h, L.ptr(inserts, C.c_double), L.ptr(bias, C.c_double), len(inserts),
                                                      L.ptr(out, C.c_double)))
        return tuple(out)

    def multinomial_cov(self, p, v, r):
        p, v = L.as_f64(p), L.as_f64(v)
        if p.ndim != 1 or v.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1)")
        if p.shape[0] != v.shape[0]:
            raise ValueError("p and v must be same shape")  # multinomial_cov.pyx:21-22
        out = C.c_double()
        self.check(self.lib.nb200_multinomial_cov(self.h, L.ptr(p, C.c_double), L.ptr(v, C.c_double), len(p), int(r), C.byref(out)))
        return float(out.value)


_default = {}


def default_engine(device=0):
    """Process-wide engine per device (the reference's functions are free functions)."""
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]

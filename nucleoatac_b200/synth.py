"""Deterministic synthetic workload for the occ+nuc scoring path (BASELINE.json configs).

Host-side, numpy only, shared by the parity tests, bench.py (GPU arm, CPU baseline and the
reference arm) so that all of them score the same reads.  Coordinates are per contig like the
reference's (pyatac/chunk.py:132-175): chunk k lies on virtual contig k // 100 000 at
[10 000 + 12 000 (k mod 100 000), +10 000), so any chunk count (configs[3]: 500 000) stays inside
int32; its reads depend only on (seed, k).

  reads      n ~ Poisson(density * L) fragments; insert size from the 3-component mixture
             0.55 * (Gamma(2.2, 22) + 38)  +  0.35 * N(188, 18)  +  0.10 * N(370, 30), rounded, kept if
             0 <= size < 2000; centre ~ U[start-600, end+600); stored as raw BAM fields
             (pos = left - 4, tlen = size + 8) so the ATAC shift of fragments.pyx:28-31 is exercised.
             15 % of the chunks carry a planted nucleosome array (dyads every 165 +- 10 bp; fragments of
             size N(170, 18) centred N(0, 4 + |size-147|/3) around a dyad).
  sequence   i.i.d. A/C/G/T with p = .3/.2/.2/.3 and 0.1 % N, over [start - 400, end + 400).
  VMat       R x W template (default 251 x 251, sizes [0, 251)): V-shaped ridge, gaussian smoothed,
             normalised like VMat.norm (pyatac/VMat.py:98-103).
  occ model  nfr / nuc size distributions = the analytic mixture components on [0, upper).
"""
import os

import numpy as np

from .engine import PackedBatch

SEED = 20261017
CHUNK_LEN = 10000
CHUNK_STRIDE = 12000
CHUNK0 = 10000
CHUNKS_PER_CONTIG = 100000   # 1.2 Gbp per virtual contig: the largest coordinate stays below 2^31
SEQ_MARGIN = 400
HERE = os.path.dirname(os.path.abspath(__file__))


def chunk_contig(k):
    """Name of the virtual contig chunk k lies on."""
    return "synth%d" % (k // CHUNKS_PER_CONTIG)


def chunk_span(k, length=CHUNK_LEN):
    s = CHUNK0 + (k % CHUNKS_PER_CONTIG) * CHUNK_STRIDE
    return s, s + length


def _checked_i32(a, what):
    a = np.asarray(a, dtype=np.int64)
    if a.size and (a.min() < -2 ** 31 or a.max() >= 2 ** 31):
        raise OverflowError("%s does not fit int32 (BAM coordinates are int32 per contig)" % what)
    return a.astype(np.int32)


def _sizes(rng, n):
    comp = rng.random(n)
    size = np.where(comp < 0.55, rng.gamma(2.2, 22.0, n) + 38.0,
                    np.where(comp < 0.90, rng.normal(188.0, 18.0, n), rng.normal(370.0, 30.0, n)))
    return np.rint(size).astype(np.int64)


def make_chunk(k, length=CHUNK_LEN, density=0.25, seed=SEED, with_seq=True, seq_margin=SEQ_MARGIN):
    """-> (start, end, pos int32[], tlen int32[], seq uint8[] | None, seq_start)."""
    rng = np.random.default_rng([seed, k])
    s, e = chunk_span(k, length)
    n = rng.poisson(density * length)
    size = _sizes(rng, n)
    centre = rng.integers(s - 600, e + 600, n)
    if rng.random() < 0.15:  # planted nucleosome array
        dyads = []
        p = s + int(rng.integers(60, 200))
        while p < e:
            dyads.append(p)
            p += int(rng.integers(155, 176))
        m = rng.poisson(0.35 * length)
        nsize = np.rint(rng.normal(170.0, 18.0, m)).astype(np.int64)
        d = np.asarray(dyads)[rng.integers(0, len(dyads), m)]
        ncentre = d + np.rint(rng.normal(0.0, 1.0, m) * (4.0 + np.abs(nsize - 147) / 3.0)).astype(np.int64)
        size = np.concatenate([size, nsize])
        centre = np.concatenate([centre, ncentre])
    ok = (size >= 0) & (size < 2000)
    size, centre = size[ok], centre[ok]
    left = centre - (size - 1) // 2
    order = np.argsort(left, kind="stable")  # BAM order: coordinate sorted
    pos = _checked_i32(left[order] - 4, "read position")
    tlen = _checked_i32(size[order] + 8, "template length")
    seq = None
    if with_seq:
        codes = rng.choice(5, size=length + 2 * seq_margin, p=[0.2997, 0.1998, 0.1998, 0.2997, 0.001])
        seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    return s, e, pos, tlen, seq, s - seq_margin


def make_batch(k0, n, length=CHUNK_LEN, density=0.25, seed=SEED, with_seq=True, seq_margin=SEQ_MARGIN):
    """PackedBatch of chunks k0 .. k0+n-1."""
    return PackedBatch.from_chunks([make_chunk(k, length, density, seed, with_seq, seq_margin) for k in range(k0, k0 + n)])


def size_mixture(upper):
    """Analytic pmf-like components of the read-size mixture on sizes [0, upper): (nfr, nuc, all)."""
    from math import gamma as G
    x = np.arange(upper, dtype=np.float64)
    xm = np.maximum(x - 38.0, 0.0)
    nfr = 0.55 * xm ** 1.2 * np.exp(-xm / 22.0) / (22.0 ** 2.2 * G(2.2))
    nuc = 0.35 * np.exp(-0.5 * ((x - 188.0) / 18.0) ** 2) / (18.0 * np.sqrt(2 * np.pi))
    di = 0.10 * np.exp(-0.5 * ((x - 370.0) / 30.0) ** 2) / (30.0 * np.sqrt(2 * np.pi))
    floor = 1e-7
    return nfr + floor, nuc + floor, nfr + nuc + di + 3 * floor


def make_vmat(R=251, W=251, lower=0):
    """Synthetic positive V-plot template: ridge at |offset| ~ (size-147)/2, smoothed, VMat.norm'ed."""
    w = W // 2
    k = np.arange(W, dtype=np.float64) - w
    sizes = np.arange(lower, lower + R, dtype=np.float64)
    _, nuc, _ = size_mixture(lower + R)
    mat = np.empty((R, W))
    for r, sz in enumerate(sizes):
        off = max(sz - 147.0, 0.0) / 2.0
        sd = 6.0 + abs(sz - 147.0) / 6.0
        ridge = np.exp(-0.5 * ((np.abs(k) - off) / sd) ** 2)
        mat[r] = (0.02 + ridge) * (nuc[lower + r] + 2e-4)
    # VMat.norm, pyatac/VMat.py:98-103
    tmp1 = mat / np.sum(mat)
    tmp2 = np.ones(mat.shape) * (1.0 / mat.size)
    mat = mat / (np.sum(mat * tmp1) - np.sum(mat * tmp2))
    return (mat / mat.shape[1]) * 10.0, lower, lower + R


def read_pwm(name="Human"):
    """PWM.open, pyatac/bias.py:47-76 (name = bundled PWM or a path)."""
    path = name if os.path.exists(name) else os.path.join(HERE, "pwm", name + ".PWM.txt")
    state, mat, up, down, nucs = "", [], None, None, None
    with open(path) as fh:
        for line in fh:
            if "#up" in line:
                state = "up"
            elif "#down" in line:
                state = "down"
            elif "#mat" in line:
                state = "mat"
            elif "#nucleotides" in line:
                state = "nucleotides"
            elif state == "up":
                up = int(line.strip("\n"))
            elif state == "down":
                down = int(line.strip("\n"))
            elif state == "nucleotides":
                nucs = line.strip("\n").split()
            elif state == "mat":
                mat.append([float(x) for x in line.strip("\n").split("\t")])
    if up is None or down is None or nucs is None:
        raise Exception("PWM decriptor file appeas to be missing some needed components")
    return np.array(mat), up, down, nucs


class Workload:
    """Run constants of the synthetic configuration (BASELINE.json configs[1..4])."""

    def __init__(self, R=251, W=251, upper=251, lower=0):
        self.upper = max(upper, lower + R)
        self.vmat, self.v_lower, self.v_upper = make_vmat(R, W, lower)
        nfr, nuc, allp = size_mixture(self.upper)
        self.nfr_probs = nfr / nfr.sum()
        self.nuc_probs = nuc / nuc.sum()
        self.fragmentsizes = allp / allp.sum()
        self.pwm, self.pwm_up, self.pwm_down, self.nucleotides = read_pwm("Human")

    def configure(self, eng, use_bias=True, xcor_mode=0, sd=10):
        eng.set_pwm(self.pwm, self.pwm_up, self.pwm_down, self.nucleotides)
        eng.set_fragment_sizes(self.fragmentsizes)  # before the VMat: the scaled template needs sizes up to vmat.upper
        eng.set_vmat(self.vmat, self.v_lower, self.v_upper)
        eng.set_occ_model(self.nuc_probs, self.nfr_probs)
        eng.configure_nuc(sd=sd, use_bias=use_bias, xcor_mode=xcor_mode)
        eng.configure_occ(upper=self.upper, use_bias=use_bias)

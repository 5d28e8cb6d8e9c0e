"""Loaders for the committed golden fixtures (tests/golden, made by make_golden.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Example:
    """Inputs of the reference's bundled example (reads, sequence, PWM, VMat, fits)."""

    def __init__(self, z):
        self.z = z
        self.chrom_names = [str(x) for x in z["chrom_names"]]
        self.n_chunks = len(z["chunk_start"])
        self.pwm = z["pwm"]
        self.pwm_up = int(z["pwm_up"])
        self.pwm_down = int(z["pwm_down"])
        self.nucleotides = [str(x) for x in z["pwm_nucleotides"]]
        self.vmat = (z["vmat"], int(z["vmat_lower"]), int(z["vmat_upper"]))
        self.vmat_example = (z["vmat_example"], int(z["vmat_example_lower"]), int(z["vmat_example_upper"]))
        self.fragmentsizes = z["fragmentsizes"]
        self.occ_fit = z["occ_fit"]

    def chunk(self, i):
        z = self.z
        return self.chrom_names[int(z["chunk_chrom"][i])], int(z["chunk_start"][i]), int(z["chunk_end"][i])

    def reads(self, i):
        z = self.z
        a, b = int(z["frag_off"][i]), int(z["frag_off"][i + 1])
        return z["frag_pos"][a:b], z["frag_tlen"][a:b]

    def sequence(self, i):
        """(bytes uint8[], genomic coordinate of the first base)."""
        z = self.z
        a, b = int(z["seq_off"][i]), int(z["seq_off"][i + 1])
        return z["seq"][a:b], int(z["seq_start"][i])

    def seq_slice(self, i, start, end):
        seq, s0 = self.sequence(i)
        assert start >= s0 and end <= s0 + len(seq)
        return bytes(seq[start - s0:end - s0]).decode()


def load_example():
    return Example(np.load(os.path.join(GOLD, "example_inputs.npz")))


def load_golden():
    return np.load(os.path.join(GOLD, "example_golden.npz"))


def track_close(gold, mine, sig=12, slack=4.0, atol=0.0):
    """Compare a track with values parsed from the reference's '%.12g' text output.

    Returns (ok, worst) where worst = max (|gold-mine| - atol) / (half unit in the 12th digit of gold);
    ``atol`` absorbs float64 cancellation noise of values that are differences of larger terms."""
    gold = np.asarray(gold, dtype=np.float64)
    mine = np.asarray(mine, dtype=np.float64)
    if gold.shape != mine.shape:
        return False, float("inf")
    gn, mn = np.isnan(gold), np.isnan(mine)
    if not np.array_equal(gn, mn):
        return False, float("inf")
    g, m = gold[~gn], mine[~gn]
    if g.size == 0:
        return True, 0.0
    mag = np.maximum(np.abs(g), 1e-300)
    unit = 10.0 ** (np.floor(np.log10(mag)) - (sig - 1))
    unit = np.maximum(unit, 1e-16)  # values printed like 1.2e-17 keep 12 digits; floor for exact zeros
    worst = float(np.max(np.maximum(np.abs(g - m) - atol, 0.0) / (0.5 * unit)))
    return worst <= slack, worst

"""Tn5 PWM + insertion-bias track (pyatac/bias.py:16-107); the PWM scan runs on the device."""
import os

import numpy as np

from . import seq
from .engine import default_engine
from .tracks import Track

_PWM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pwm")


class PWM:
    def __init__(self, mat, up, down, nucleotides):
        self.mat, self.up, self.down, self.nucleotides = mat, up, down, nucleotides

    def save(self, filename):
        with open(filename, "w") as fh:
            fh.write("#PWM Descriptor File\n#Contains PWM and pertinent information\n")
            fh.write("#up\n%d\n#down\n%d\n#nucleotides\n%s\n#mat\n" % (self.up, self.down, "\t".join(self.nucleotides)))
            for row in self.mat:
                fh.write("\t".join(repr(float(x)) for x in row) + "\n")

    @staticmethod
    def open(name):
        """A bundled PWM name ('Human') or a path to a PWM descriptor file (bias.py:47-76)."""
        path = name if os.path.exists(name) else os.path.join(_PWM_DIR, name + ".PWM.txt")
        if not os.path.exists(path):
            raise Exception("PWM %s not found" % name)
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
                elif line.startswith("#"):
                    continue
                elif state == "up":
                    up = int(line.strip("\n"))
                elif state == "down":
                    down = int(line.strip("\n"))
                elif state == "nucleotides":
                    nucs = line.strip("\n").split()
                elif state == "mat":
                    mat.append([float(x) for x in line.strip("\n").split("\t")])
        if up is None or down is None or nucs is None or not mat:
            raise Exception("PWM decriptor file appeas to be missing some needed components")
        return PWM(np.array(mat), up, down, nucs)


class InsertionBiasTrack(Track):
    def __init__(self, chrom, start, end, log=True):
        Track.__init__(self, chrom, start, end, name="insertion bias", log=log)

    def computeBias(self, fasta, chromDict, pwm):
        """log-bias[p] = sum_j log PWM[nuc(seq[p-up+j]), j] (bias.py:85-92), nb200_bias_track on the device."""
        self.slop(chromDict, up=pwm.up, down=pwm.down)
        sequence = seq.get_sequence(self, fasta)
        eng = default_engine()
        eng.set_pwm(pwm.mat, pwm.up, pwm.down, pwm.nucleotides)
        self.vals = eng.bias_track(sequence)
        self.start += pwm.up
        self.end -= pwm.down

    def get(self, start=None, end=None, pos=None, log=None):
        out = Track.get(self, start, end, pos)
        if log is None or bool(log) == bool(self.log):
            return out
        return np.log(out) if log else np.exp(out)

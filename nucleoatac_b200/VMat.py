"""VMat: the V-plot template (pyatac/VMat.py:14-218) -- container, text IO and the once-per-run processing
steps of `vprocess` (trim / symmetrize / norm_y / smooth / norm; host numpy+scipy like the reference)."""
import numpy as np

from .utils import fmt12


class VMat_Error(Exception):
    def __init__(self, value):
        self.value = value

    def __str__(self):
        return repr(self.value)


class VMat:
    def __init__(self, mat, lower, upper):
        mat = np.asarray(mat, dtype=np.float64)
        if mat.ndim != 2 or mat.shape[0] != upper - lower:
            raise VMat_Error("mat shape is not consistent with insert limits")
        self.mat, self.upper, self.lower = mat, upper, lower
        self.w = mat.shape[1] // 2

    def trim(self, lower, upper, w):
        up, dn = upper - self.lower, lower - self.lower
        left, right = self.w - w, self.w + w + 1
        if up > self.mat.shape[0] or dn < 0 or left < 0 or right > self.mat.shape[1]:
            raise VMat_Error("Mat is smaller than desired trim")
        self.mat = np.array(self.mat[dn:up, left:right])
        self.lower, self.upper, self.w = lower, upper, w

    def symmetrize(self):
        """Odd sizes mirror about the centre column, even sizes about centre - 0.5 (VMat.py:55-64)."""
        w = self.w
        for j in range(self.lower, self.upper):
            row = self.mat[j - self.lower]
            if j % 2 == 1:
                half = (row[:w + 1] + row[w:][::-1]) * 0.5
                self.mat[j - self.lower] = np.hstack((half, half[:-1][::-1]))
            else:
                half = (row[w:-1] + row[:w][::-1]) * 0.5
                self.mat[j - self.lower] = np.hstack((half[::-1], half, row[-1]))

    def smooth(self, sd=1):
        from scipy import ndimage
        self.mat = ndimage.gaussian_filter(self.mat, sd, mode="constant")

    def smooth1d(self, sd=1, axis=1):
        from scipy import ndimage
        self.mat = ndimage.gaussian_filter1d(self.mat, sd, axis, mode="nearest")

    def norm(self):
        """Scale so that signal minus an even background is 10 / window width (VMat.py:98-103)."""
        total = np.sum(self.mat)
        self.mat = self.mat / (np.sum(self.mat * (self.mat / total)) - np.sum(self.mat * (1.0 / self.mat.size)))
        self.mat = (self.mat / self.mat.shape[1]) * 10.0

    def norm_y(self, dist):
        for i in range(self.mat.shape[0]):
            self.mat[i] = self.mat[i] * (dist.get(size=i + self.lower) / np.sum(self.mat[i]))

    def save(self, filename):
        with open(filename, "w") as out:
            out.write("#VMat Descriptor File\n#Contains VMat and pertinent information\n")
            out.write("#lower\n%s\n#upper\n%s\n#mat\n" % (self.lower, self.upper))
            for row in self.mat:
                out.write("\t".join(fmt12(v) for v in row) + "\n")

    @staticmethod
    def open(filename):
        state, mat, lower, upper = "", [], None, None
        with open(filename) as fh:
            for line in fh:
                if "#lower" in line:
                    state = "lower"
                elif "#upper" in line:
                    state = "upper"
                elif "#mat" in line:
                    state = "mat"
                elif "#" in line:
                    state = "other"
                elif state == "lower":
                    lower = int(line.strip("\n"))
                elif state == "upper":
                    upper = int(line.strip("\n"))
                elif state == "mat":
                    mat.append([float(x) for x in line.strip("\n").split("\t")])
        if lower is None or upper is None:
            raise VMat_Error("VMat decriptor file appeas to be missing some needed components")
        return VMat(np.array(mat), lower, upper)

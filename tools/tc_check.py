#!/usr/bin/env python
"""Compare the tcgen05 background xcor (xcor_mode 2) with the fp64 CUDA-core kernel (mode 1) on device."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nucleoatac_b200 import synth
from nucleoatac_b200.engine import Engine

def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 251
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 251
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    eng = Engine(0)
    wl = synth.Workload(R, W)
    pb = synth.make_batch(0, n, seq_margin=max(400, W + R // 2 + 24))
    res = {}
    for mode in (1, 2):
        wl.configure(eng, use_bias=True, xcor_mode=mode)
        t = time.time()
        out = eng.process_nuc(pb)
        res[mode] = out
        print("mode", mode, "wall %.3f s" % (time.time() - t))
    a, b = res[1]["background"], res[2]["background"]
    scale = np.abs(a).max()
    err = np.abs(a - b)
    print("background: max abs err %.3e  scale %.3e  -> rel-to-scale %.3e ; max pointwise rel %.3e" % (
        err.max(), scale, err.max() / scale, (err / np.maximum(np.abs(a), 1e-300)).max()))
    i = int(np.argmax(err))
    print("worst at", i, a[i], b[i], " first values", a[:3], b[:3])
    for k in ("norm_signal", "smoothed"):
        e = np.abs(res[1][k] - res[2][k])
        print(k, "max abs err %.3e" % np.nanmax(e))
    same = np.array_equal(res[1]["cand_pos"], res[2]["cand_pos"]) and np.array_equal(res[1]["cand_flag"], res[2]["cand_flag"])
    print("candidates identical:", same, "n cands", int(res[1]["cand_count"].sum()), int(res[2]["cand_count"].sum()))
    print("profile", {k: v for k, v in eng.profile_report().items() if "bx" in k or "emax" in k})

if __name__ == "__main__":
    main()

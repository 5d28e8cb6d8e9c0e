#!/usr/bin/env python
"""One profiled pass of the hot path for ncu: B synthetic chunks (default 400), 251x251 VMat, two warm-up passes, then
one pass of nuc_run + occ_run.  Prints the number of kernel launches per pass so that the capture can be limited to it:

    python tools/ncu_pass.py --count                      # -> launches per pass (K)
    ncu --set full --clock-control none --import-source on --launch-skip 2K --launch-count K -o gpurun_out/prof \
        python tools/ncu_pass.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nucleoatac_b200 import synth
from nucleoatac_b200.engine import Engine


def main():
    B = 400
    eng = Engine(0)
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True, xcor_mode=0)
    pb = synth.make_batch(0, B)
    h = None
    n0 = 0
    for it in range(3):
        h = eng.upload(pb, h)
        eng.nuc_run(h)
        eng.occ_run(h)
        eng.sync(h)
        n = sum(v[0] for v in eng.profile_report().values())
        if it == 1:
            per_pass = n - n0
        n0 = n
    if "--count" in sys.argv:
        print(per_pass)
    eng.free_batch(h)
    eng.close()


if __name__ == "__main__":
    main()

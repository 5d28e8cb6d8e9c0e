"""Interleaved A/B of the forms of k_occ_mle (lanes per window x resident blocks) on the bench workload, inside one process.

    python tools/mle_ab.py [--batch 2000] [--reps 4]

Every form scores the same uploaded batch; the three occupancy grids (vals, lower, upper) of each form are compared bit for
bit with the first one's (4 M windows at the default size), then the forms are timed round-robin (CUDA events of the
library's own per-kernel profile).  One JSON line per form."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _concat(_, synth, PackedBatch, n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 2)) as pool:
        return PackedBatch.from_chunks(pool.map(synth.make_chunk, range(n), chunksize=16))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--forms", default="8:4:fast:1,8:4:fast:0,8:4:canonical:1,16:6:fast:1", help="lanes per window : resident blocks per SM : epilogue : staged inputs")
    args = ap.parse_args()
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import Engine, PackedBatch
    forms = [tuple(f.split(":")) for f in args.forms.split(",")]
    eng = Engine(0)
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True, xcor_mode=0)
    pb = PackedBatch.from_chunks([synth.make_chunk(k) for k in range(args.batch)]) if args.batch < 64 else _concat(None, synth, PackedBatch, args.batch)
    h = eng.upload(pb)
    eng.profile(True)
    ref = None
    times = {f: [] for f in forms}
    for rep in range(args.reps + 1):
        for f in forms:
            os.environ["NB200_MLE_GL"], os.environ["NB200_MLE_LB"], os.environ["NB200_MLE_EPI"], os.environ["NB200_MLE_STAGE"] = f
            eng.profile_reset()
            eng.occ_run(h)
            eng.sync(h)
            if rep == 0:   # warm-up round doubles as the identity check
                out = eng.occ_alloc(pb)
                eng.occ_download(h, out)
                eng.sync(h)
                got = [out[k].copy() for k in ("vals", "lower_bound", "upper_bound", "smoothed_vals")]
                if ref is None:
                    ref = got
                else:
                    for a, b in zip(ref, got):
                        assert np.array_equal(a, b, equal_nan=True), "form %s differs from form %s" % (f, forms[0])
            else:
                times[f].append(eng.profile_report()["k_occ_mle"][1])
    for f in forms:
        print(json.dumps(dict(kernel="k_occ_mle", lanes_per_window=int(f[0]), blocks_per_sm=int(f[1]), epilogue=f[2], staged=int(f[3]), ms=sorted(times[f]), median_ms=float(np.median(times[f])),
                              batch=args.batch, identical_to_first=True)))
    eng.free_batch(h)
    eng.close()


if __name__ == "__main__":
    main()

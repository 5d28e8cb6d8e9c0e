#!/usr/bin/env python
"""BASELINE.json configs[4]: VMat size sweep R = W in 101 ... 501 at fixed 10 000 x 10 kb chunks on one B200 (nuc path).
Prints one JSON line per size: nuc bp/s, the dense-contraction kernel's ms, useful TFLOP/s and fraction of the measured
bf16 peak, and the HBM-side figure (algorithmic bytes of the non-contraction stages / their time)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nucleoatac_b200 import synth
from nucleoatac_b200.engine import Engine


def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [101, 151, 201, 251, 301, 401, 501]
    n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    B = 1000
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
    eng = Engine(0)
    for R in sizes:
        W = R
        wl = synth.Workload(R, W)
        wl.configure(eng, use_bias=True, xcor_mode=0)
        margin = max(400, W + R // 2 + 24)
        batches = [synth.make_batch(i * B, B, seq_margin=margin) for i in range(2)]
        h = None
        for pb in batches:  # warm-up
            h = eng.upload(pb, h)
            eng.nuc_run(h)
            eng.sync(h)
        eng.profile_reset()
        eng.profile(True)
        steps = max(2, n_chunks // B)
        total = 0.0
        for i in range(steps):
            h = eng.upload(batches[i % 2], h)
            eng.sync(h)
            eng.flush_l2(h)
            eng.timer_start(h)
            eng.nuc_run(h)
            eng.timer_stop(h)
            total += eng.timer_ms(h)
        prof = eng.profile_report()
        eng.profile(False)
        bp = steps * batches[0].total_len
        kname = next((k for k in ("k_nuc_bx_ts", "k_nuc_bx_tc") if k in prof), "k_nuc_bx_fp64")
        kms = prof[kname][1]
        useful = 2.0 * R * W * bp / (kms * 1e-3) / 1e12
        other_ms = total - kms
        # non-contraction stages: inputs 8 B/fragment + 1 B/base, outputs 6 f64 tracks (48 B/bp) -> ~50-60 B/bp
        alg_bytes = bp * (0.25 * 8 + 1.08 + 48.0)
        print(json.dumps(dict(R=R, W=W, chunks=steps * B, nuc_bp_per_s=bp / (total * 1e-3), ms_per_1000_chunks=total / steps,
                              xcor_kernel=kname, xcor_ms_per_1000_chunks=kms / steps, xcor_useful_tflops=useful,
                              xcor_frac_of_bf16_peak=useful / peaks["bf16_tflops_sustained"],
                              other_stages_gbs=alg_bytes / (other_ms * 1e-3) / 1e9,
                              other_stages_frac_of_hbm=alg_bytes / (other_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              per_kernel_ms={k: round(v[1] / steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})),
              flush=True)
        eng.free_batch(h)
    eng.close()


if __name__ == "__main__":
    main()

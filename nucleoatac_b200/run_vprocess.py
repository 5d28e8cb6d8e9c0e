"""`nucleoatac vprocess` (nucleoatac/run_vprocess.py:15-42): turn a raw V-plot into the template `nuc` slides:
trim -> symmetrize -> insert-size normalisation with the nucleosomal size distribution from `occ` -> gaussian
smoothing -> norm.  Host numpy/scipy, once per run (146 x 121 array)."""
import os

from .fragmentsizes import FragmentSizes
from .VMat import VMat

DEFAULT_VPLOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vplot", "standard_vplot.VMat")


def run_vprocess(args):
    vmat = VMat.open(args.vplot)
    vmat.trim(args.lower, args.upper, args.flank)
    vmat.symmetrize()
    if args.sizes is not None:
        vmat.norm_y(FragmentSizes.open(args.sizes))
    if args.smooth > 0:
        vmat.smooth(sd=args.smooth)
    vmat.norm()
    vmat.save(args.out + ".VMat")
    return vmat

"""Fuzziness fit of the called nucleosomes (Nucleosome.getFuzz, nucleoatac/NucleosomeCalling.py:137-194): one to three
gaussians fitted to the smoothed signal around a call by scipy's L-BFGS-B, as the reference does it (host
scipy by design).  A fit costs milliseconds of Python and there are thousands of calls per Mbp of real data, so the fits of
a scored batch are spread over a pool of worker processes -- the reference gets the same effect from running whole chunks
in `Pool(cores)` (run_nuc.py:164-187).  This module imports numpy / scipy only: workers start fast and never touch the GPU.

NB200_FUZZ_PROCS sets the number of workers (default: min(16, cores - 1); 0 or 1 = fit in the calling process)."""
import atexit
import os

import numpy as np

_pool = None
_pool_size = 0
_broken = False
MIN_JOBS_FOR_POOL = 64


def fit_fuzz(job):
    """job = (sig, means, smooth_sd): the signal slice [left, right) (negative values already clipped to 0), the expected
    peak offsets inside it (the call itself and its close neighbours) and the smoothing sd that seeds the variance.
    Returns (fuzz, weight, fitted offset of the first gaussian)."""
    from scipy import optimize
    sig, means, smooth_sd = job
    bounds, guesses = (), ()
    top = max(sig)
    for m in means:
        bounds += ((2 ** 2, 50 ** 2), (0.001, top * 1.1), (m - 10, m + 10))
        guesses += (smooth_sd ** 2, top * 0.9, m)
    xs = np.linspace(0, len(sig) - 1, len(sig))

    def err(pars, y):
        # the reference's objective value for value (sum of `norm` terms, then the squared error added up left to right)
        # without its per-element Python loops: max(n) -> n.max() (the same number), sum(d ** 2) -> the last element of a
        # running sum (the same additions in the same order).  L-BFGS-B calls this a few hundred times per nucleosome.
        fit = np.zeros(len(y))
        for j in range(len(pars) // 3):
            v = pars[3 * j]
            n = 1.0 / np.sqrt(2 * np.pi * v) * np.exp(-(xs - pars[3 * j + 2]) ** 2 / (2 * v))
            fit += n * (pars[3 * j + 1] / n.max())
        d = fit - y
        return np.cumsum(d ** 2)[-1]

    res = optimize.minimize(err, guesses, args=(sig,), bounds=bounds, method="L-BFGS-B")
    return float(np.sqrt(res["x"][0])), float(res["x"][1]), float(res["x"][2])


def n_procs():
    env = os.environ.get("NB200_FUZZ_PROCS")
    if env is not None:
        return max(0, int(env))
    n = min(16, (os.cpu_count() or 1) - 1)
    host = os.environ.get("NB200_HOST_WORKERS")   # set by the drivers from `--cores N`
    if host:
        n = min(n, max(1, int(host)))
    return max(0, n)


def _get_pool(n):
    global _pool, _pool_size
    if _pool is not None and _pool_size == n:
        return _pool
    close_pool()
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    # forkserver: the workers descend from a clean server process, not from this one (which holds a CUDA context and threads)
    # and single-threaded numeric libraries in the workers (n workers x a BLAS pool each oversubscribes the cores: measured 6x
    # slower than no pool at all); the variables are read when a worker first imports numpy
    keys = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")
    saved = {k: os.environ.get(k) for k in keys}
    os.environ.update({k: "1" for k in keys})
    try:
        ctx = mp.get_context("forkserver")
        _pool = ProcessPoolExecutor(n, mp_context=ctx)
        list(_pool.map(abs, range(n)))   # start the server and the workers while the variables are set
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    _pool_size = n
    return _pool


def close_pool():
    global _pool, _pool_size
    if _pool is not None:
        _pool.shutdown(wait=False, cancel_futures=True)
    _pool, _pool_size = None, 0


atexit.register(close_pool)


def _fit_here(jobs):
    """The fits in the calling process, with the BLAS pool limited to one thread while they run (L-BFGS-B's tiny vector
    operations are slower on a pool of threads than on one: 2x measured)."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        return [fit_fuzz(j) for j in jobs]
    with threadpool_limits(limits=1, user_api="blas"):
        return [fit_fuzz(j) for j in jobs]


def fit_many(jobs):
    """Results of fit_fuzz for every job, in order; on the worker pool when there are enough of them to pay for it.  A pool
    whose workers cannot start (e.g. a main module that cannot be re-imported) is reported once and the fits run here."""
    global _broken
    n = n_procs()
    if n <= 1 or len(jobs) < MIN_JOBS_FOR_POOL or _broken:
        return _fit_here(jobs)
    from concurrent.futures.process import BrokenProcessPool
    try:
        return list(_get_pool(n).map(fit_fuzz, jobs, chunksize=max(1, min(32, len(jobs) // (4 * n)))))
    except BrokenProcessPool as ex:
        import sys
        sys.stderr.write("nucleoatac_b200.fuzz: worker pool unavailable (%s); fitting in the calling process\n" % ex)
        _broken = True
        close_pool()
        return _fit_here(jobs)

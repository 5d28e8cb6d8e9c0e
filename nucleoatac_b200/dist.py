"""Multi-GPU plumbing: one process per GPU, BED chunk k -> rank k mod N (round-robin, no data-path collective).
The only exchanges are the end-of-run reductions of small vectors (fragment-size histogram, nuc_dist, V-plot sums:
fragments.pyx:122-145, run_occ.py:117-121, make_vplot.py:70-73) and putting the per-rank output files back into
chunk order.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used purely as transport."""
import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(items, rank, world):
    """The chunks of this rank, in order: k = rank, rank + world, ..."""
    return [x for k, x in enumerate(items) if k % world == rank]


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def allreduce_sum(arr):
    """Sum a numpy array over all ranks (identity without an initialised process group)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return arr
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def barrier():
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.barrier()


class ShardWriter:
    """Per-rank text output + the byte offset after every chunk, so that rank 0 can interleave the ranks' blocks
    back into global chunk order (the reference's pool.map is order preserving; tabix needs sorted output)."""

    def __init__(self, path, rank, world):
        self.final, self.world, self.rank = path, world, rank
        self.path = path if world == 1 else "%s.rank%d" % (path, rank)
        self.fh = open(self.path, "w")
        self.offsets = []

    def write(self, text):
        self.fh.write(text)

    def end_chunk(self):
        self.offsets.append(self.fh.tell())

    def close(self):
        self.fh.close()
        if self.world > 1:
            np.savetxt(self.path + ".idx", np.asarray(self.offsets, dtype=np.int64), fmt="%d")

    @staticmethod
    def merge(path, world, n_chunks):
        """Interleave `path.rank<r>` blocks into `path` (chunk k is block k // world of rank k % world)."""
        if world == 1:
            return
        fhs = [open("%s.rank%d" % (path, r), "rb") for r in range(world)]
        idx = [np.atleast_1d(np.loadtxt("%s.rank%d.idx" % (path, r), dtype=np.int64)) if os.path.getsize("%s.rank%d.idx" % (path, r)) else
               np.zeros(0, np.int64) for r in range(world)]
        with open(path, "wb") as out:
            for k in range(n_chunks):
                r, j = k % world, k // world
                lo = int(idx[r][j - 1]) if j > 0 else 0
                hi = int(idx[r][j])
                fhs[r].seek(lo)
                out.write(fhs[r].read(hi - lo))
        for r, fh in enumerate(fhs):
            fh.close()
            os.remove("%s.rank%d" % (path, r))
            os.remove("%s.rank%d.idx" % (path, r))

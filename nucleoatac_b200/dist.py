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
    """torch.distributed if a process group is up.  A group can only exist if somebody imported torch.distributed: a
    single-rank run never pays for the import (several seconds on a cold start, more than scoring a genome)."""
    import sys
    dist = sys.modules.get("torch.distributed")
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist
    return None


def init_from_env(device=None):
    """Join the torchrun rendezvous (env://): NCCL with this process bound to its own GPU (LOCAL_RANK), gloo without a
    GPU.  The device is set BEFORE the group exists so that every collective's tensors land on this rank's GPU."""
    import torch
    import torch.distributed as td
    rank, world, local = env_rank_world()
    if td.is_initialized():
        return rank, world, local
    if torch.cuda.is_available():
        dev = local if device is None else int(device)
        torch.cuda.set_device(dev)
        td.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        td.init_process_group("gloo")
    return rank, world, local


def require_group(world):
    """A sharded run (world > 1) needs the reductions: refuse to hand back per-shard partial sums."""
    if world > 1 and _dist() is None:
        if "MASTER_ADDR" in os.environ and "RANK" in os.environ and "WORLD_SIZE" in os.environ:
            init_from_env()
            return
        raise RuntimeError("--world %d without a process group: launch with `python -m torch.distributed.run --nproc-per-node %d "
                           "--master-addr 127.0.0.1 ...` (or export MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE) so that fragment sizes, "
                           "nuc_dist and V-plot sums are reduced over the shards" % (world, world))


def allreduce_sum(arr, world=None):
    """Sum a numpy array over all ranks.  Identity for a single-rank run; a sharded run (`world` > 1) without an
    initialised process group is an error, not a silent partial sum."""
    dist = _dist()
    if dist is None:
        if world is not None and world > 1:
            require_group(world)
            dist = _dist()
        if dist is None:
            return arr
    if dist.get_world_size() == 1:
        return arr
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.cuda(torch.cuda.current_device())   # the device init_from_env bound this rank to
    dist.all_reduce(t)
    return t.cpu().numpy()


def barrier(world=None):
    dist = _dist()
    if dist is None and world is not None and world > 1:
        require_group(world)
        dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            import torch
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


class ShardWriter:
    """Per-rank text output + the byte offset after every chunk, so that rank 0 can interleave the ranks' blocks
    back into global chunk order (the reference's pool.map is order preserving; tabix needs sorted output)."""

    def __init__(self, path, rank, world):
        self.final, self.world, self.rank = path, world, rank
        self.path = path if world == 1 else "%s.rank%d" % (path, rank)
        self.fh = open(self.path, "wb")
        self.offsets = []
        self.pos = 0

    def write(self, text):
        self.write_bytes(text.encode())

    def write_bytes(self, data):
        self.fh.write(data)
        self.pos += len(data)

    def end_chunk(self):
        self.offsets.append(self.pos)

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


class BatchWriter:
    """Host side of the drivers' chunk loop: the outputs of a scored batch are formatted on a pool of threads (the native
    formatter releases the GIL) and written in chunk order by one background thread while the next batch is being scored.
    The reference does the same with writer processes behind queues (run_occ.py:103-116, run_nuc.py:166-182).  At most one
    batch is being written while one is scored; an exception of the writer surfaces at the next submit() or at close().
    The background thread must not touch the device engine (calls on a context are serialised by the caller)."""

    def __init__(self, threads=None):
        from concurrent.futures import ThreadPoolExecutor
        from . import hostio
        self.fmt = ThreadPoolExecutor(max(1, threads or hostio.host_workers()))
        self.bg = ThreadPoolExecutor(1)
        self.prev = None

    def map(self, fn, jobs):
        return list(self.fmt.map(fn, jobs))

    def submit(self, fn, *args):
        self.drain()
        self.prev = self.bg.submit(fn, *args)

    def drain(self):
        prev, self.prev = self.prev, None
        if prev is not None:
            prev.result()

    def close(self):
        try:
            self.drain()
        finally:
            self.bg.shutdown(wait=True)
            self.fmt.shutdown(wait=True)


def bind_to_gpu_numa(device=0):
    """Pin this process (CPU affinity + preferred memory node) to the NUMA node the GPU hangs off, before any pinned host
    buffer is allocated: on a two-socket host a D2H copy into memory of the other socket crosses the inter-socket link and
    runs at roughly half the PCIe rate (32 vs 55 GB/s measured on the B200 boxes).  Best effort: returns a dict describing
    what was done, never raises."""
    info = dict(gpu_numa_node=None, cpus_bound=None, mempolicy=None)
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = device
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[device])
            except Exception:
                idx = device
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:   # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
        mine = os.sched_getaffinity(0) & cpus
        if mine:
            os.sched_setaffinity(0, mine)
            info["cpus_bound"] = len(mine)
        try:   # set_mempolicy(MPOL_PREFERRED, {node}): pinned buffers land next to the GPU even if no local CPU is ours
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(238, 1, ctypes.byref(mask), 16 * 64 + 1)   # x86-64: __NR_set_mempolicy = 238, MPOL_PREFERRED = 1
            info["mempolicy"] = "preferred" if rc == 0 else "errno %d" % ctypes.get_errno()
        except Exception as ex:
            info["mempolicy"] = "failed: %s" % ex
    except Exception as ex:
        info["error"] = str(ex)[:200]
    return info


def reset_mempolicy():
    """Back to the default memory policy (undoes the preference set by bind_to_gpu_numa)."""
    try:
        import ctypes
        ctypes.CDLL(None, use_errno=True).syscall(238, 0, None, 0)
    except Exception:
        pass

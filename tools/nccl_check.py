#!/usr/bin/env python
"""2+ GPU check of the library's own NCCL reductions (nb200_allreduce_f64 / _i64): run under torchrun; the NCCL unique
id travels over a gloo process group (CPU), the reductions themselves go through libnucleo_b200 -> libnccl."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from nucleoatac_b200 import _lib
from nucleoatac_b200.engine import Engine


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    eng = Engine(local)
    lib = eng.lib
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        assert lib.nb200_nccl_unique_id(buf) == 0, lib.nb200_last_error(None)
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    dist.broadcast(uid, 0)
    raw = (C.c_ubyte * 128)(*uid.tolist())
    eng.check(lib.nb200_nccl_init(eng.h, raw, rank, world))
    x = np.arange(251, dtype=np.float64) * (rank + 1) + 0.25
    eng.check(lib.nb200_allreduce_f64(eng.h, _lib.ptr(x, C.c_double), len(x)))
    exp = np.arange(251, dtype=np.float64) * sum(r + 1 for r in range(world)) + 0.25 * world
    assert np.array_equal(x, exp), (rank, x[:4], exp[:4])
    y = (np.arange(251, dtype=np.int64) + 2 ** 40) * (rank + 1)
    eng.check(lib.nb200_allreduce_i64(eng.h, _lib.ptr(y, C.c_int64), len(y)))
    assert np.array_equal(y, (np.arange(251, dtype=np.int64) + 2 ** 40) * sum(r + 1 for r in range(world)))
    eng.check(lib.nb200_nccl_finalize(eng.h))
    eng.close()
    dist.barrier()
    if rank == 0:
        print("nb200 NCCL all-reduce f64 / i64 ok on %d GPUs" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

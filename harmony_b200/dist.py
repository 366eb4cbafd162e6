"""Multi-GPU plumbing (one process per GPU): cells are sharded contiguously across
ranks, the library all-reduces the small K x B / K x J x (d+1) statistics itself (NCCL), and
``torch.distributed`` (nccl or gloo backend) only carries the communicator id and test data.
"""
import ctypes

import numpy as np


def shard_bounds(N, world_size, rank=None):
    """Contiguous, balanced split of cells [0, N): rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(N), int(world_size))
    b = np.zeros(world_size + 1, dtype=np.int64)
    for r in range(world_size):
        b[r + 1] = b[r] + base + (1 if r < rem else 0)
    if rank is None:
        return b
    return int(b[rank]), int(b[rank + 1])


def broadcast_bytes(payload, nbytes, src=0, device=None):
    """Broadcast a fixed-size byte string from ``src`` over the default process group."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if device is not None:
        dev = device
    if dist.get_rank() == src:
        t = torch.tensor(list(payload[:nbytes]), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def make_comm(N_global, id_source=None):
    """Returns the ``comm`` tuple the ``harmony`` constructor takes:
    (rank, world_size, communicator id, N_global, this rank's cell offset).  Needs an initialised
    torch.distributed process group; rank 0 creates the NCCL unique id through the C ABI."""
    import torch.distributed as dist
    from . import _lib
    rank, world = dist.get_rank(), dist.get_world_size()
    if id_source is None:
        def id_source():
            buf = ctypes.create_string_buffer(128)
            st = _lib.lib().hb_comm_unique_id(buf)
            if st != 0:
                raise RuntimeError("hb_comm_unique_id failed (libnccl.so.2 not loadable?)")
            return buf.raw
    uid = id_source() if rank == 0 else b"\0" * 128
    uid = broadcast_bytes(uid, 128, 0)
    lo, _ = shard_bounds(N_global, world, rank)
    return (rank, world, uid, int(N_global), lo)

"""Frame sharding across ranks (one process per GPU) and the ordered host-side reassembly of the stream.

VC-2 is intra-only: picture n is an independent unit, so a sequence is dealt to the ranks in
contiguous batches and no collective sits on the data path.  The one ordered step is on the host: rank 0
collects the slice payloads in picture order and writes the data units, whose parse-info headers chain the
previous / next offsets (reference: src/Library/src/DataUnit.cpp:112-123, 236-266, 359-368).  The gather of
the (compressed) payload bytes uses torch.distributed object collectives on the host - gloo or nccl alike.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def shard_pictures(n_pictures, rank, world, batch):
    """Indices of the pictures rank `rank` codes: batches of `batch` consecutive pictures dealt round robin
    (batch b -> rank b % world), the order EncodeStream/DecodeStream --gpus use."""
    out = []
    for b0 in range(0, n_pictures, batch):
        if (b0 // batch) % world == rank:
            out.extend(range(b0, min(b0 + batch, n_pictures)))
    return out


def _host_lib():
    lib = C.CDLL(os.path.join(_HERE, "libvc2host.so"))
    lib.vc2host_wrap_hq_stream.restype = C.c_longlong
    lib.vc2host_wrap_hq_stream.argtypes = [C.c_int] * 13 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    return lib


def wrap_hq_stream(fmt, geom, payloads):
    """Sequence header + one HQ picture data unit per payload + end of sequence (host code, libvc2host.so).
    fmt: dict(height, width, chroma (0/1/2), frame_rate, top_field_first, bitdepth); geom: a vc2 Geom."""
    bufs = [np.frombuffer(p, np.uint8) for p in payloads]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * len(bufs))(*[b.size for b in bufs])
    out = np.zeros(sum(b.size for b in bufs) + 64 * (len(bufs) + 2), np.uint8)
    n = _host_lib().vc2host_wrap_hq_stream(fmt["height"], fmt["width"], fmt["chroma"], fmt["frame_rate"], int(fmt["top_field_first"]),
                                           fmt["bitdepth"], geom.kernel, geom.depth, geom.slices_x, geom.slices_y, geom.prefix, geom.scalar,
                                           len(bufs), ptrs, lens, out.ctypes.data, out.size)
    if n < 0:
        raise ValueError("stream framing failed (%d)" % n)
    return out[:n].tobytes()


def gather_stream(local_payloads, local_indices, n_pictures, fmt, geom, group=None):
    """Every rank passes the payloads of the pictures it coded; rank 0 returns the whole stream, the others None."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = list(zip(local_indices, local_payloads))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    ordered = [None] * n_pictures
    for part in gathered:
        for i, p in part:
            if ordered[i] is not None:
                raise ValueError("picture %d coded twice" % i)
            ordered[i] = p
    missing = [i for i, p in enumerate(ordered) if p is None]
    if missing:
        raise ValueError("pictures %s were not coded by any rank" % missing[:8])
    return wrap_hq_stream(fmt, geom, ordered)

"""CPU-only, world size 2 over gloo: the multi-GPU host logic (vc2_reference_b200/sharding.py) - dealing pictures to
ranks and the ordered reassembly of the stream on rank 0.  The per-rank "encoder" here is the reference's own
payload bytes (tests/golden/framing_*.npz), so the reassembled stream must equal the reference's stream."""
import json
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import vc2_reference_b200 as vc2
from vc2_reference_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "md5.json")))


def test_shard_pictures_partition():
    for n, world, batch in [(240, 8, 8), (10, 2, 4), (7, 4, 2), (3, 8, 1), (0, 2, 4)]:
        parts = [sharding.shard_pictures(n, r, world, batch) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        for r, p in enumerate(parts):
            assert all((i // batch) % world == r for i in p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, reps, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = GOLD[name]["params"]
        z = np.load(os.path.join(HERE, "golden", "framing_%s.npz" % name))
        packaged = z["packaged"].tobytes()
        g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
        n_slices = g.slices_x * g.slices_y
        pics, pos = [], 0
        for _ in range(c["frames"]):
            off = vc2.hq_index_slices(packaged[pos:], n_slices, c["P"], c["S"])
            pics.append(packaged[pos:pos + int(off[-1])])
            pos += int(off[-1])
        n = c["frames"] * reps                       # the clip repeated: picture i carries payload i % frames
        mine = sharding.shard_pictures(n, rank, world, batch)
        fmt = dict(height=c["h"], width=c["w"], chroma={"444": 0, "422": 1, "420": 2}[c["fmt"]], frame_rate=c["r"], top_field_first=True,
                   bitdepth=c["bits"])
        stream = sharding.gather_stream([pics[i % c["frames"]] for i in mine], mine, n, fmt, g)
        if rank == 0:
            q.put(stream)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,reps,batch", [("S07_LeGall_d1_444", 1, 1), ("S01_LeGall_d3_422", 3, 2)])
def test_two_rank_reassembly_equals_reference_stream(name, reps, batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, reps, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    stream = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    c = GOLD[name]["params"]
    z = np.load(os.path.join(HERE, "golden", "framing_%s.npz" % name))
    ref = z["stream"].tobytes()
    if reps == 1:
        assert stream == ref
    else:
        # the reference stream of the 2-frame clip is a prefix pattern: same sequence header, and unit k of our stream
        # equals unit (k % frames) of the reference except for the picture number and the previous-offset chain
        import hostapi
        ours, theirs = hostapi.parse_units(stream), hostapi.parse_units(ref)
        assert len(ours) == c["frames"] * reps + 2
        assert stream[:int(ours[1][1])] == ref[:int(theirs[1][1])]
        for k, u in enumerate(ours[1:-1]):
            t = theirs[1 + k % c["frames"]]
            assert int(u[2]) == int(t[2])
            a, b = int(u[1]), int(t[1])
            assert stream[a + 17:a + int(u[2])] == ref[b + 17:b + int(t[2])]
            assert int.from_bytes(stream[a + 13:a + 17], "big") == k
        for a, b in zip(ours[:-1], ours[1:]):
            assert int(b[1]) == int(a[1]) + int(a[2]) and int(b[3]) == int(a[2])

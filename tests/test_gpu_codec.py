"""GPU parity of the fused batched codec against golden digests produced by running the reference
(tests/golden/md5.json, made by tests/golden/make_golden.py).  Bit exact: Transform, Quantised,
Indices, Packaged (slice payload bytes) and Decoded pictures."""
import hashlib
import json
import os

import numpy as np
import pytest

import gen
import vc2_reference_b200 as vc2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "md5.json")))
FULL = os.environ.get("VC2_TEST_FULL", "1") != "0"


def be32(planes):
    return b"".join(np.ascontiguousarray(p).astype(">i4").tobytes() for p in planes)


def run_case(ctx, name, batch):
    gold = GOLD[name]
    c = gold["params"]
    taps = gold["taps"]
    mode = c["mode"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"],
                      c["P"] if mode != "LD" else 0, c["S"] if mode != "LD" else 1)
    frames = [gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"]) for f in range(c["frames"])]
    assert hashlib.md5(b"".join(frames)).hexdigest() == taps["input"]["md5"]
    n = len(frames)
    if True:
        enc = vc2.Codec(ctx, g, mode, qindex=c["q"] or 0, picture_bytes=c["s"] or 0, luma_depth=c["bits"], max_pictures=batch)
        md = {k: hashlib.md5() for k in ("Transform", "Quantised", "Indices", "Packaged")}
        payloads = []
        for base in range(0, n, batch):
            m = min(batch, n - base)
            for i in range(m):
                enc.upload_picture(i, frames[base + i])
            enc.encode(m)
            for i in range(m):
                payload, qidx, off = enc.download_payload(i)
                payloads.append(payload)
                md["Transform"].update(be32(enc.read_transform(i)))
                md["Quantised"].update(be32(enc.read_quantised(i)))
                md["Indices"].update(qidx.astype(np.uint8).tobytes())
                md["Packaged"].update(payload)
                assert off[-1] == len(payload)
        for k, h in md.items():
            if "enc_" + k in taps:
                assert h.hexdigest() == taps["enc_" + k]["md5"], (name, k)
        assert sum(map(len, payloads)) == taps["enc_Packaged"]["bytes"]
        enc.close()
    if mode == "LD" and os.path.exists(os.path.join(HERE, "golden", name + ".ldpayload.npz")):
        # the slice payloads cut out of the reference-encoded stream (fixture): the LD encoder reproduces them
        assert payloads == load_ld_payloads(name, g, c)
    # decode (bit depth 8 streams decode to one byte per sample, DecodeStream.cpp:268-271)
    bps = 1 if c["bits"] == 8 else 2
    dec = vc2.Codec(ctx, g, mode if mode == "LD" else "HQ_ConstQ", qindex=0, picture_bytes=c["s"] or 0, bytes_per_sample=bps,
                    luma_depth=c["bits"], max_pictures=batch)
    out = hashlib.md5()
    mdq, mdt = hashlib.md5(), hashlib.md5()
    for base in range(0, n, batch):
        m = min(batch, n - base)
        for i in range(m):
            dec.upload_payload(i, payloads[base + i])
        dec.decode(m)
        for i in range(m):
            dec.slot_status(i)
            out.update(dec.download_picture(i))
            mdt.update(be32(dec.read_transform(i)))
    assert mdt.hexdigest() == taps["dec_Transform"]["md5"], (name, "dec_Transform")
    assert out.hexdigest() == taps["dec_Decoded"]["md5"], (name, "Decoded")
    dec.close()


def load_ld_payloads(name, g, c):
    path = os.path.join(HERE, "golden", name + ".ldpayload.npz")
    z = np.load(path)
    return [z["p%d" % i].tobytes() for i in range(c["frames"])]


SMALL = sorted(k for k in GOLD if k[0] in "SBLWDR" and not GOLD[k]["params"].get("extra"))


@pytest.mark.parametrize("name", SMALL)
def test_small_cases(ctx, name):
    run_case(ctx, name, batch=2)


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C5"])
def test_baseline_configs(ctx, name):
    run_case(ctx, name, batch=2)


@pytest.mark.skipif(not FULL, reason="VC2_TEST_FULL=0")
@pytest.mark.parametrize("name", ["C4a", "C4b"])
def test_baseline_configs_8k(ctx, name):
    run_case(ctx, name, batch=1)


def test_batch_slots_are_independent(ctx):
    """a batch of 3 pictures gives the same bytes as three single-picture runs"""
    c = GOLD["S08_DD137_d4_422"]["params"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
    frames = [gen.frame_bytes(5, f, c["w"], c["h"], c["fmt"], c["bits"]) for f in range(3)]
    one = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=9, luma_depth=c["bits"], max_pictures=1)
    singles = []
    for f in frames:
        one.upload_picture(0, f)
        one.encode(1)
        singles.append(one.download_payload(0)[0])
    many = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=9, luma_depth=c["bits"], max_pictures=3)
    for i, f in enumerate(frames):
        many.upload_picture(i, f)
    many.encode(3)
    for i in range(3):
        assert many.download_payload(i)[0] == singles[i]
    # end-to-end host API gives the same bytes and round-trips
    pics = [np.frombuffer(f, np.uint8).copy() for f in frames]
    bufs = [np.zeros(many.payload_capacity, np.uint8) for _ in frames]
    lens = many.encode_host(pics, bufs)
    for i in range(3):
        assert bufs[i][:lens[i]].tobytes() == singles[i]
    outs = [np.zeros(many.picture_bytes, np.uint8) for _ in frames]
    many.decode_host(bufs, lens, outs)
    many.upload_payload(0, singles[0])
    many.decode(1)
    assert outs[0].tobytes() == many.download_picture(0)


@pytest.mark.parametrize("name", ["S01_LeGall_d3_422", "S05_Fidelity_d2_422", "S11_DD97_d4_422", "C1"])
def test_host_entry_points_vs_golden(ctx, name):
    """vc2_codec_encode_host / _decode_host (device-side slice index, several chunks per call) reproduce the
    reference's payload and decoded bytes; prefix / scalar variants included"""
    c = GOLD[name]["params"]
    taps = GOLD[name]["taps"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
    reps = 3                                   # 2 golden frames x 3 -> 6 pictures through a 4-slot codec: two chunks
    frames = [np.frombuffer(gen.frame_bytes(c["seed"], f % c["frames"], c["w"], c["h"], c["fmt"], c["bits"]), np.uint8).copy()
              for f in range(c["frames"] * reps)]
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=c["q"], luma_depth=c["bits"], max_pictures=4)
    bufs = [np.zeros(k.payload_capacity, np.uint8) for _ in frames]
    lens = k.encode_host(frames, bufs)
    for r in range(reps):
        md = hashlib.md5()
        for f in range(c["frames"]):
            i = r * c["frames"] + f
            md.update(bufs[i][:lens[i]].tobytes())
        assert md.hexdigest() == taps["enc_Packaged"]["md5"], (name, "Packaged", r)
    outs = [np.zeros(k.picture_bytes, np.uint8) for _ in frames]
    k.decode_host(bufs, lens, outs)
    for r in range(reps):
        md = hashlib.md5()
        for f in range(c["frames"]):
            md.update(outs[r * c["frames"] + f].tobytes())
        assert md.hexdigest() == taps["dec_Decoded"]["md5"], (name, "Decoded", r)
    k.close()


def test_decode_host_reports_truncated_payload(ctx):
    """a payload that ends inside a slice is a stream error (the reference's reader would run off the end)"""
    c = GOLD["S08_DD137_d4_422"]["params"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=c["q"], luma_depth=c["bits"], max_pictures=2)
    frames = [np.frombuffer(gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"]), np.uint8).copy() for f in range(2)]
    bufs = [np.zeros(k.payload_capacity, np.uint8) for _ in frames]
    lens = k.encode_host(frames, bufs)
    outs = [np.zeros(k.picture_bytes, np.uint8) for _ in frames]
    with pytest.raises(vc2.Vc2Error) as e:
        k.decode_host(bufs, [lens[0], lens[1] // 2], outs)
    assert e.value.status == -9       # VC2_ERR_STREAM
    k.decode_host(bufs, lens, outs)      # the codec is still usable afterwards
    k.close()


def test_device_decode_indexes_the_payload_itself(ctx):
    """decode_dev rebuilds the slice offsets from the length bytes of the payload (Slices.cpp:544-605): a second codec
    that only ever receives payload bytes + length decodes to the same picture, and a cut payload is a stream error"""
    c = GOLD["S08_DD137_d4_422"]["params"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
    n = 3
    frames = [gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"]) for f in range(n)]
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=c["q"], luma_depth=c["bits"], max_pictures=n)
    for i, f in enumerate(frames):
        k.upload_picture(i, f)
    k.encode(n)
    k.decode(n)      # payload and payload length come straight from the encoder, on the device
    pays = [k.download_payload(i) for i in range(n)]
    pics = [k.download_picture(i) for i in range(n)]
    d = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=0, luma_depth=c["bits"], max_pictures=n)
    for i in range(n):
        d.upload_payload(i, pays[i][0])
    d.decode(n)
    for i in range(n):
        d.slot_status(i)
        assert d.download_picture(i) == pics[i]
        assert np.array_equal(d.download_payload(i)[2], pays[i][2])     # the rebuilt offset table = the encoder's
    d.upload_payload(1, pays[1][0][: len(pays[1][0]) // 2])
    d.decode(n)
    d.slot_status(0)
    with pytest.raises(vc2.Vc2Error) as e:
        d.slot_status(1)
    assert e.value.status == -9       # VC2_ERR_STREAM
    d.close()
    k.close()


def test_pipelined_mode_is_bit_identical(ctx):
    """vc2_codec_set_pipelined only changes the ordering between sub-batches of consecutive calls, never the bytes"""
    c = GOLD["S08_DD137_d4_422"]["params"]
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["wdepth"], c["u"], c["a"], c["P"], c["S"])
    n = 8
    frames = [gen.frame_bytes(7, f, c["w"], c["h"], c["fmt"], c["bits"]) for f in range(2 * n)]
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=c["q"], luma_depth=c["bits"], max_pictures=n)
    want = []
    for base in (0, n):
        for i in range(n):
            k.upload_picture(i, frames[base + i])
        k.encode(n)
        k.decode(n)
        want.append([(k.download_payload(i)[0], k.download_picture(i)) for i in range(n)])
    k.set_pipelined(True)
    for rep in range(2):
        for half, base in enumerate((0, n)):
            for i in range(n):
                k.upload_picture(i, frames[base + i])
            for _ in range(3):          # back-to-back calls without a host sync in between
                k.encode(n)
                k.decode(n)
            got = [(k.download_payload(i)[0], k.download_picture(i)) for i in range(n)]
            assert got == want[half]
    k.close()

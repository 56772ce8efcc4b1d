"""GPU: the fused codec against the plain-C oracle (oracle/vc2_oracle.c, itself pinned to the compiled reference) on seeded
random geometries - kernel, depth, colour format, bit depth, slice shape, prefix / scalar, quantiser, sizes with and
without padding, narrow pictures (general loop only) and pictures wide enough for the lifting kernels' fast loop."""
import numpy as np
import pytest

import gen
import orcapi as orc
import vc2_reference_b200 as vc2

pytestmark = pytest.mark.gpu
KERNELS = ["DD97", "LeGall", "DD137", "Haar0", "Haar1", "Fidelity", "Daub97"]


def cases():
    rng = np.random.default_rng(20261017)
    out = []
    for i in range(28):
        depth = int(rng.integers(1, 5))
        fmt = ["444", "422", "420"][int(rng.integers(0, 3))]
        cell = 1 << depth
        # slice shape in cells; 4:2:x halves the chroma, so the luma slice stays even in that direction
        u = int(rng.integers(1, 3)) * (2 if fmt == "420" else 1)
        a = int(rng.integers(1, 3)) * (2 if fmt != "444" else 1)
        # every fourth case is wide enough for interior strips (the fast loop of the lifting kernels): >= 1000 columns
        want_w = int(rng.integers(1000, 1500)) if i % 4 == 0 else int(rng.integers(40, 400))
        want_h = int(rng.integers(120, 320)) if i % 4 == 0 else int(rng.integers(24, 200))
        nx, ny = max(1, want_w // (a * cell)), max(1, want_h // (u * cell))
        ph, pw = ny * u * cell, nx * a * cell
        # unpadded size: up to one cell short of the padded size, even so that 4:2:x chroma stays integral
        h = max(2, ph - 2 * int(rng.integers(0, cell // 2 + 1))) if cell > 2 else ph
        w = max(2, pw - 2 * int(rng.integers(0, cell // 2 + 1))) if cell > 2 else pw
        out.append(dict(id="z%02d" % i, w=w, h=h, fmt=fmt, bits=[8, 10, 12, 16][int(rng.integers(0, 4))], kernel=KERNELS[i % 7], depth=depth,
                        u=u, a=a, q=int(rng.integers(4, 40)), P=int(rng.integers(0, 3)), S=int(rng.integers(2, 9)) * (u * a), seed=900 + i))
    return out


@pytest.mark.parametrize("c", cases(), ids=lambda c: "%s_%s_d%d_%s_%dx%d" % (c["id"], c["kernel"], c["depth"], c["fmt"], c["w"], c["h"]))
def test_codec_vs_oracle(ctx, c):
    try:
        g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["depth"], c["u"], c["a"], c["P"], c["S"])
    except vc2.Vc2Error:
        pytest.skip("sliceSizeIsValid rejects this combination")
    ch, cw = gen.chroma_dims(c["w"], c["h"], c["fmt"])
    frames = [gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], smooth=(c["seed"] % 2 == 0)) for f in range(2)]
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=c["q"], luma_depth=c["bits"], max_pictures=2)
    for i, f in enumerate(frames):
        k.upload_picture(i, f)
    k.encode(2)
    kern = vc2.KERNELS[c["kernel"]]
    for i, f in enumerate(frames):
        try:
            want = orc.encode_picture_hq_constq(f, c["h"], c["w"], ch, cw, c["bits"], kern, c["depth"], g.slices_y, g.slices_x, c["q"], c["P"], c["S"])
        except orc.OrcError as e:       # e.g. slice scalar too small for this content: the codec must say the same
            with pytest.raises(vc2.Vc2Error) as ge:
                k.download_payload(i)
            assert ge.value.status == e.status
            k.close()
            return
        got = k.download_payload(i)[0]
        assert got == want, (c, i, len(got), len(want))
    k.decode(2)
    for i, f in enumerate(frames):
        k.slot_status(i)
        want = orc.encode_picture_hq_constq(f, c["h"], c["w"], ch, cw, c["bits"], kern, c["depth"], g.slices_y, g.slices_x, c["q"], c["P"], c["S"])
        pic = orc.decode_picture_hq(want, c["h"], c["w"], ch, cw, c["bits"], kern, c["depth"], g.slices_y, g.slices_x, c["P"], c["S"])
        assert k.download_picture(i) == pic, (c, i)
    k.close()


def _noise_frame(seed, w, h, fmt, bits):
    """full-range noise, MSB justified big-endian words: coefficients far beyond 15 bits at q = 0"""
    ch, cw = gen.chroma_dims(w, h, fmt)
    rng = np.random.default_rng(seed)
    v = rng.integers(0, 1 << bits, size=w * h + 2 * ch * cw, dtype=np.uint32) << (16 - bits)
    return v.astype(">u2").tobytes()


@pytest.mark.parametrize("kernel,depth,bits", [("DD137", 3, 13), ("Daub97", 3, 13), ("LeGall", 2, 10)])
def test_narrow_block_overflow_falls_back_to_32_bit(ctx, kernel, depth, bits):
    """HQ_ConstQ encodes and HQ decodes keep 16-bit quantised coefficients between the lifting kernels and the slice
    coders; q = 0 on full-range noise overflows that block (the LeGall case stays inside it), and the slot has to come out
    exactly as through the 32-bit path: device-resident calls, the status / download entry points, the host-buffer calls"""
    w, h, fmt, q, P, S = 416, 256, "422", 0, 1, 16
    g = vc2.make_geom(h, w, fmt, kernel, depth, 1, 2, P, S)
    ch, cw = gen.chroma_dims(w, h, fmt)
    kern = vc2.KERNELS[kernel]
    frames = [_noise_frame(77 + i, w, h, fmt, bits) if i != 1 else gen.frame_bytes(5, i, w, h, fmt, bits) for i in range(3)]
    want, pics = [], []
    for f in frames:
        try:
            want.append(orc.encode_picture_hq_constq(f, h, w, ch, cw, bits, kern, depth, g.slices_y, g.slices_x, q, P, S))
        except orc.OrcError:
            pytest.skip("the reference rejects this content at q = 0 (slice scalar)")
        pics.append(orc.decode_picture_hq(want[-1], h, w, ch, cw, bits, kern, depth, g.slices_y, g.slices_x, P, S))
    k = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=q, luma_depth=bits, max_pictures=4)
    for i, f in enumerate(frames):
        k.upload_picture(i, f)
    k.encode(3)
    k.decode(3)                       # round trip in the slot, then ask for the results in both orders
    assert k.download_picture(2) == pics[2]
    for i in range(3):
        assert k.download_payload(i)[0] == want[i], (kernel, i)
        assert k.download_picture(i) == pics[i], (kernel, i)
    # a separate decoder fed with the payloads
    d = vc2.Codec(ctx, g, "HQ_ConstQ", qindex=q, luma_depth=bits, max_pictures=4)
    for i in range(3):
        d.upload_payload(i, want[i])
    d.decode(3)
    for i in range(3):
        d.slot_status(i)
        assert d.download_picture(i) == pics[i], (kernel, i)
    # host-buffer calls, more pictures than slots
    src = [np.frombuffer(frames[i % 3], np.uint8).copy() for i in range(7)]
    bufs = [np.zeros(k.payload_capacity, np.uint8) for _ in src]
    lens = k.encode_host(src, bufs)
    for i in range(7):
        assert bufs[i][:lens[i]].tobytes() == want[i % 3], (kernel, "encode_host", i)
    outs = [np.zeros(k.picture_bytes, np.uint8) for _ in src]
    d.decode_host(bufs, lens, outs)
    for i in range(7):
        assert outs[i].tobytes() == pics[i % 3], (kernel, "decode_host", i)
    k.close()
    d.close()

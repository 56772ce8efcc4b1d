"""CPU-only: pins the plain-C restatement (oracle/vc2_oracle.c) against (1) the reference's own known
answers, (2) the compiled unmodified reference (oracle/_ref/libvc2ref.so) on random inputs, and
(3) the golden digests made by running the reference command lines (tests/golden/md5.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

import gen
import orcapi as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "md5.json")))
KN = ["DD97", "LeGall", "DD137", "Haar0", "Haar1", "Fidelity", "Daub97"]


def rnd(shape, lo, hi, seed):
    return np.random.default_rng(seed).integers(lo, hi + 1, size=shape, dtype=np.int64).astype(np.int32)


# ---- (1) known answers --------------------------------------------------------------------------------
def test_reference_unit_test_vectors():
    # /root/reference/tests/Quantisation.cpp:30-36
    assert orc.quant(12, 0) == 12 and orc.quant(12, 2) == 8 and orc.quant(-12, 2) == -8 and orc.quant(-12, -2) == -12
    # :6-12 - index beyond the table is an error
    with pytest.raises(orc.OrcError):
        orc.quant(1, 130)


def test_survey_known_answers():
    qs = [0, 1, 2, 3, 4, 5, 8, 12, 20, 33]
    assert [orc.quant(1000, q) for q in qs] == [1000, 800, 666, 571, 500, 400, 250, 125, 31, 3]
    assert [orc.quant(-37, q) for q in qs] == [-37, -29, -24, -21, -18, -14, -9, -4, -1, 0]
    assert [orc.scale(5, q) for q in qs] == [5, 7, 8, 10, 11, 14, 22, 44, 176, 1675]
    assert [orc.scale(-1, q) for q in qs] == [-1, -2, -2, -3, -3, -4, -6, -12, -48, -457]
    assert orc.scale(0, 17) == 0
    vlc = {0: (1, 0x1), 1: (4, 0x2), -1: (4, 0x3), 2: (4, 0x6), -2: (4, 0x7), 3: (6, 0x2), 4: (6, 0x6), 7: (8, 0x2), 8: (8, 0x6),
           -15: (10, 0x3), 16: (10, 0x6), 100: (14, 0x1046), -1000: (20, 0x55107)}
    for v, want in vlc.items():
        assert orc.signed_vlc(v) == want, v
    assert orc.slice_bytes(3, 4, 1000, 1).ravel().tolist() == [83, 83, 84, 83, 83, 84, 83, 83, 84, 83, 83, 84]
    assert (orc.slice_bytes(135, 120, 2073600, 1) == 128).all()
    assert orc.quant_matrix(5, 4).tolist() == [0, 4, 4, 8, 8, 8, 12, 13, 13, 17, 17, 17, 21]
    assert orc.quant_matrix(3, 3).tolist() == [16, 12, 12, 8, 8, 8, 4, 4, 4, 0]


DWT_KAT = {
    0: ([-101, -2, -80, -13, 28, 141, -30, 83], [0, 0, 0, 0, 25, 101, 25, 0]),
    1: ([-103, -6, -79, -6, 31, 146, -30, 78], [0, 0, 0, 0, 25, 101, 25, 0]),
    2: ([-102, -2, -85, -13, 29, 143, -26, 83], [0, 0, -3, 0, 28, 101, 28, 0]),
    3: ([-39, 3, -21, 15, 21, 27, -14, 39], [17, 0, 17, 0, 17, 0, 17, 0]),
    4: ([-80, 6, -44, 30, 40, 54, -30, 78], [34, 0, 34, 0, 34, 0, 34, 0]),
    5: ([-168, 9, -139, -17, 45, 58, -23, 26], [-4, 3, -13, -4, 16, 21, 15, -10]),
    6: ([-136, 0, -108, -20, 33, 128, -30, 84], [2, 0, 1, -3, 26, 70, 27, -7]),
}


@pytest.mark.parametrize("k", range(7))
def test_dwt_known_answers(k):
    y, x = np.mgrid[0:4, 0:8]
    a = ((3 * x * x + 17 * y) % 101 - 50).astype(np.int32)
    t = orc.dwt_forward(a, k, 1)
    assert t[0].tolist() == DWT_KAT[k][0] and t[1].tolist() == DWT_KAT[k][1]
    assert np.array_equal(orc.dwt_inverse(t, k, 1, (4, 8)), a)


# ---- (2) against the compiled reference ------------------------------------------------------------------
def test_tables_vs_reference(ref):
    for q in range(-3, 120):
        assert orc.quant_factor(q) == ref.quant_factor(q)
        assert orc.quant_offset(q) == ref.quant_offset(q)
    for k in range(7):
        for d in range(0, 7):
            assert orc.quant_matrix(k, d).tolist() == ref.quant_matrix(k, d).tolist()
    for v in list(range(-70, 71)) + [255, -256, 1023, -4097, 65534, -65534]:
        assert orc.signed_vlc(v) == ref.signed_vlc(v)
    rng = np.random.default_rng(3)
    for _ in range(300):
        v, q = int(rng.integers(-500000, 500000)), int(rng.integers(0, 60))
        assert orc.quant(v, q) == ref.quant(v, q)
        assert orc.scale(orc.quant(v, q), q) == ref.scale(ref.quant(v, q), q)
    for d in range(1, 6):
        for size in (1, 7, 8, 1080, 1081):
            assert orc.padded_size(size, d) == ref.padded_size(size, d)
        for luma, chroma in ((1080, 1080), (1920, 960), (188, 94), (100, 100)):
            for n in range(0, 10):
                assert orc.slice_size_is_valid(d, luma, chroma, n) == ref.slice_size_is_valid(d, luma, chroma, n)
    for a in [(3, 4, 1000, 1), (7, 5, 9999, 3), (17, 11, 123457, 4), (2, 2, 64, 8)]:
        assert (orc.slice_bytes(*a) == ref.slice_bytes(*a)).all()


@pytest.mark.parametrize("k", range(7))
@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_dwt_vs_reference(ref, k, depth):
    for i, (h, w) in enumerate([(4, 8), (16, 16), (37, 53), (66, 130), (135, 240)]):
        if min(h, w) < (1 << depth) // 2:
            continue
        a = rnd((h, w), -2048, 2047, 10 * depth + i)
        want = ref.dwt_forward(a, k, depth)
        assert np.array_equal(orc.dwt_forward(a, k, depth), want), (k, depth, h, w)
        c = rnd(want.shape, -3000, 3000, 99 + i)
        assert np.array_equal(orc.dwt_inverse(c, k, depth, (h, w)), ref.dwt_inverse(c, k, depth, (h, w)))


@pytest.mark.parametrize("depth,ny,nx,mh,mw", [(1, 3, 5, 2, 1), (2, 4, 3, 1, 2), (3, 5, 6, 1, 2), (4, 2, 3, 1, 1)])
def test_quantisers_vs_reference(ref, depth, ny, nx, mh, mw):
    ph, pw = (ny * mh) << depth, (nx * mw) << depth
    c = rnd((ph, pw), -20000, 20000, depth)
    qidx = rnd((ny, nx), 0, 40, 50 + depth)
    qm = ref.quant_matrix(depth % 7, depth)
    want = ref.quantise_np(c, qidx, qm)
    assert np.array_equal(orc.quantise_np(c, qidx, qm), want)
    assert np.array_equal(orc.dequantise_np(want, qidx, qm), ref.dequantise_np(want, qidx, qm))
    small = rnd((ph, pw), -60, 60, 7 + depth)
    assert np.array_equal(orc.dequantise_ld(small, qidx, qm), ref.dequantise_ld(small, qidx, qm))


def _planes(depth, ny, nx, fmt, seed, amp):
    lh, lw = (ny * 1) << depth, (nx * 2) << depth
    ch, cw = (lh, lw) if fmt == "444" else ((lh, lw // 2) if fmt == "422" else (lh // 2, lw // 2))
    if ch % (ny << depth) or cw % (nx << depth):
        lh, lw, ch, cw = lh * 2, lw * 2, ch * 2, cw * 2
    rng = np.random.default_rng(seed)
    mk = lambda h, w: (rng.laplace(0, amp, size=(h, w))).astype(np.int32)
    return mk(lh, lw), mk(ch, cw), mk(ch, cw)


@pytest.mark.parametrize("depth,fmt,scalar,prefix", [(2, "444", 1, 0), (3, "422", 2, 1), (4, "422", 4, 0), (3, "420", 3, 2)])
def test_hq_slices_vs_reference(ref, depth, fmt, scalar, prefix):
    ny, nx = 3, 4
    y, u, v = _planes(depth, ny, nx, fmt, depth, 6.0)
    y[0, :] = 0
    u[:, :] = 0          # an all-zero component: length byte 0
    qidx = rnd((ny, nx), 0, 63, 5)
    want = ref.pack_slices(y, u, v, depth, qidx, 0, prefix, scalar)
    got = orc.pack_slices(y, u, v, depth, qidx, 0, prefix, scalar)
    assert got == want
    a = orc.unpack_slices(got, y.shape[0], y.shape[1], u.shape[0], u.shape[1], depth, ny, nx, 0, prefix, scalar)
    b = ref.unpack_slices(want, y.shape[0], y.shape[1], u.shape[0], u.shape[1], depth, ny, nx, 0, prefix, scalar)
    for p, q, r in zip(a, b, (y, u, v, qidx)):
        assert np.array_equal(p, q) and np.array_equal(p, r)
    # CBR writer: generous budget, V takes the remainder
    sb = np.full((ny, nx), 4 + scalar * ((len(want) // (ny * nx)) // scalar + 40), np.int32)
    try:
        wantc = ref.pack_slices(y, u, v, depth, qidx, 1, prefix, scalar, sb)
    except ref.RefError:
        with pytest.raises(orc.OrcError):
            orc.pack_slices(y, u, v, depth, qidx, 1, prefix, scalar, sb)
    else:
        assert orc.pack_slices(y, u, v, depth, qidx, 1, prefix, scalar, sb) == wantc


@pytest.mark.parametrize("depth,fmt,total", [(2, "444", 500), (3, "422", 2500), (3, "420", 1800), (4, "422", 9000)])
def test_ld_encoder_vs_reference(ref, depth, fmt, total):
    """quantIndicesLD, the predictive quantiser and the LD slice writer against the compiled reference"""
    ny, nx = 3, 4
    y, u, v = _planes(depth, ny, nx, fmt, 20 + depth, 40.0)
    rng = np.random.default_rng(depth)
    for p in (y, u, v):      # a DC level, so that the prediction matters
        p[:: 1 << depth, :: 1 << depth] += rng.integers(-900, 900, size=p[:: 1 << depth, :: 1 << depth].shape).astype(np.int32)
    qm = ref.quant_matrix(depth % 7, depth)
    sb = ref.slice_bytes(ny, nx, total, 1)
    want = ref.ld_qindices(y, u, v, qm, sb)
    got = orc.ld_qindices(y, u, v, qm, sb)
    assert np.array_equal(got, want)
    assert want.min() < want.max()
    q = [ref.quantise_ld(p, want, qm) for p in (y, u, v)]
    for p, r in zip((y, u, v), q):
        assert np.array_equal(orc.quantise_ld(p, want, qm), r)
    assert orc.pack_slices_ld(q[0], q[1], q[2], depth, want, sb) == ref.pack_slices(q[0], q[1], q[2], depth, want, 2, 0, 1, sb)


def test_hq_scalar_too_small(ref):
    y, u, v = _planes(3, 2, 2, "422", 1, 4000.0)
    qidx = np.zeros((2, 2), np.int32)
    with pytest.raises(ref.RefError):
        ref.pack_slices(y, u, v, 3, qidx, 0, 0, 1)
    with pytest.raises(orc.OrcError, match="scalar is too small"):
        orc.pack_slices(y, u, v, 3, qidx, 0, 0, 1)


def test_ld_slices_vs_reference(ref):
    depth, ny, nx = 3, 3, 4
    y, u, v = _planes(depth, ny, nx, "422", 11, 1.2)
    qidx = rnd((ny, nx), 0, 50, 8)
    sb = ref.slice_bytes(ny, nx, ny * nx * 200 + 7, 1)
    stream = ref.pack_slices(y, u, v, depth, qidx, 2, 0, 1, sb)     # the reference LD writer makes the fixture
    a = orc.unpack_slices(stream, y.shape[0], y.shape[1], u.shape[0], u.shape[1], depth, ny, nx, 2, 0, 1, sb)
    b = ref.unpack_slices(stream, y.shape[0], y.shape[1], u.shape[0], u.shape[1], depth, ny, nx, 2, 0, 1, sb)
    for p, q in zip(a, b):
        assert np.array_equal(p, q)


@pytest.mark.parametrize("kernel,depth,fmt,scalar", [(0, 3, "422", 1), (2, 2, "444", 2), (1, 3, "420", 1), (1, 3, "420", 3), (6, 2, "422", 1)])
def test_cbr_rate_control_vs_reference(ref, kernel, depth, fmt, scalar):
    ny, nx = 3, 4
    y, u, v = _planes(depth, ny, nx, fmt, 20 + depth, 150.0)
    qm = ref.quant_matrix(kernel, depth)
    total = (y.size + 2 * u.size) // 3
    sb = ref.slice_bytes(ny, nx, total, scalar)
    try:
        want = ref.cbr_qindices(y, u, v, qm, sb, scalar)
    except ref.RefError as e:
        with pytest.raises(orc.OrcError, match=str(e)[:20]):
            orc.cbr_qindices(y, u, v, qm, sb, scalar)
    else:
        assert np.array_equal(orc.cbr_qindices(y, u, v, qm, sb, scalar), want)


# ---- (3) against golden digests of the reference command lines ---------------------------------------------
@pytest.mark.parametrize("name", ["S01_LeGall_d3_422", "S05_Fidelity_d2_422", "S08_DD137_d4_422", "S02_DD137_d2_444", "S04_Haar1_d4_420"])
def test_whole_picture_vs_golden(name):
    c = GOLD[name]["params"]
    taps = GOLD[name]["taps"]
    if c["bits"] == 8:
        pytest.skip("the whole-picture helpers read 16-bit samples")
    ch, cw = gen.chroma_dims(c["w"], c["h"], c["fmt"])
    d = c["wdepth"]
    ny = orc.slice_size_is_valid(d, c["h"], ch, c["u"])
    nx = orc.slice_size_is_valid(d, c["w"], cw, c["a"])
    k = KN.index(c["kernel"])
    pay, dec = hashlib.md5(), hashlib.md5()
    for f in range(c["frames"]):
        raw = gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"])
        p = orc.encode_picture_hq_constq(raw, c["h"], c["w"], ch, cw, c["bits"], k, d, ny, nx, c["q"], c["P"], c["S"])
        pay.update(p)
        dec.update(orc.decode_picture_hq(p, c["h"], c["w"], ch, cw, c["bits"], k, d, ny, nx, c["P"], c["S"]))
    assert pay.hexdigest() == taps["enc_Packaged"]["md5"]
    assert dec.hexdigest() == taps["dec_Decoded"]["md5"]

"""GPU parity: every Library-surface C-ABI entry point against the compiled reference, bit exact."""
import numpy as np
import pytest

import vc2_reference_b200 as vc2

pytestmark = pytest.mark.gpu

KN = ["DD97", "LeGall", "DD137", "Haar0", "Haar1", "Fidelity", "Daub97"]


def rnd(shape, lo, hi, seed):
    return np.random.default_rng(seed).integers(lo, hi + 1, size=shape, dtype=np.int64).astype(np.int32)


def same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        first = tuple(bad[0])
        raise AssertionError("%s: %d of %d differ, first at %s: got %d want %d"
                             % (what, len(bad), got.size, first, got[first], want[first]))


# survey Appendix B.4: depth-1 transform of a[y][x] = (3x^2+17y)%101-50 (4x8), rows 0 and 1
DWT_KAT = {
    "DD97": ([-101, -2, -80, -13, 28, 141, -30, 83], [0, 0, 0, 0, 25, 101, 25, 0]),
    "LeGall": ([-103, -6, -79, -6, 31, 146, -30, 78], [0, 0, 0, 0, 25, 101, 25, 0]),
    "DD137": ([-102, -2, -85, -13, 29, 143, -26, 83], [0, 0, -3, 0, 28, 101, 28, 0]),
    "Haar0": ([-39, 3, -21, 15, 21, 27, -14, 39], [17, 0, 17, 0, 17, 0, 17, 0]),
    "Haar1": ([-80, 6, -44, 30, 40, 54, -30, 78], [34, 0, 34, 0, 34, 0, 34, 0]),
    "Fidelity": ([-168, 9, -139, -17, 45, 58, -23, 26], [-4, 3, -13, -4, 16, 21, 15, -10]),
    "Daub97": ([-136, 0, -108, -20, 33, 128, -30, 84], [2, 0, 1, -3, 26, 70, 27, -7]),
}


@pytest.mark.parametrize("kernel", KN)
def test_dwt_known_answers(ctx, kernel):
    y, x = np.mgrid[0:4, 0:8]
    a = ((3 * x * x + 17 * y) % 101 - 50).astype(np.int32)
    t = ctx.waveletTransform(a, kernel, 1)
    assert t[0].tolist() == DWT_KAT[kernel][0]
    assert t[1].tolist() == DWT_KAT[kernel][1]
    same(ctx.inverseWaveletTransform(t, kernel, 1, (4, 8)), a, "inverse of KAT")


SHAPES = [(4, 8), (2, 2), (16, 16), (37, 53), (64, 128), (66, 130), (129, 257), (135, 240), (200, 300), (270, 481)]


@pytest.mark.parametrize("kernel", KN)
@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_dwt_forward_inverse_vs_reference(ctx, ref, kernel, depth):
    k = vc2.KERNELS[kernel]
    for i, (h, w) in enumerate(SHAPES):
        if min(h, w) < (1 << depth) // 2:
            continue
        a = rnd((h, w), -512, 511, 1000 * depth + i)
        want = ref.dwt_forward(a, k, depth)
        got = ctx.waveletTransform(a, kernel, depth)
        same(got, want, "forward %s d%d %dx%d" % (kernel, depth, h, w))
        # inverse on arbitrary (not necessarily reachable) coefficients, with crop
        c = rnd(want.shape, -3000, 3000, 77 + i)
        same(ctx.inverseWaveletTransform(c, kernel, depth, (h, w)), ref.dwt_inverse(c, k, depth, (h, w)),
             "inverse %s d%d %dx%d" % (kernel, depth, h, w))
        same(ctx.inverseWaveletTransform(got, kernel, depth, (h, w)), a, "round trip")


def test_dwt_large_values_fidelity(ctx, ref):
    # 12-bit content through Fidelity depth 4 reaches +-4.6e5: int32 everywhere
    a = rnd((144, 272), -2048, 2047, 5)
    same(ctx.waveletTransform(a, "Fidelity", 4), ref.dwt_forward(a, 5, 4), "fidelity 12 bit")


def _slice_cfg(depth, ny, nx, sy, sx):
    return (ny * sy << depth, nx * sx << depth)


@pytest.mark.parametrize("depth,ny,nx,mh,mw", [(1, 3, 5, 2, 1), (2, 4, 3, 1, 2), (3, 5, 6, 1, 2), (4, 2, 3, 1, 1)])
def test_quantise_dequantise_vs_reference(ctx, ref, depth, ny, nx, mh, mw):
    ph, pw = _slice_cfg(depth, ny, nx, mh, mw)
    coef = rnd((ph, pw), -70000, 70000, depth)
    coef[::3, ::5] = rnd(coef[::3, ::5].shape, -40, 40, 9)
    for kernel in ("LeGall", "Fidelity", "Haar0"):
        qm = vc2.quant_matrix(kernel, depth)
        qidx = rnd((ny, nx), 0, 60, 3 + depth)
        want = ref.quantise_np(coef, qidx, qm)
        same(ctx.quantise_transform_np(coef, qidx, qm), want, "quantise")
        same(ctx.inverse_quantise_transform_np(want, qidx, qm), ref.dequantise_np(want, qidx, qm), "dequantise")
        small = rnd((ph, pw), -300, 300, 11)
        same(ctx.inverse_quantise_transform(small, qidx, qm), ref.dequantise_ld(small, qidx, qm), "LD dequantise")


def test_quantise_all_indices_exact_division(ctx, ref):
    # every quantiser index 0..115 (beyond that the reference's int quant_factor is negative): exactness of the
    # multiply-shift division against the reference's integer division, incl. values around multiples
    qm = np.zeros(4, np.int32)
    for q in range(0, 116):
        qf = ref.quant_factor(q)
        base = np.arange(0, 64, dtype=np.int64)[:, None] * qf // 4
        vals = (base + np.arange(-2, 3)[None, :]).ravel()
        vals = np.concatenate([vals, -vals, rnd((192,), -(1 << 28), 1 << 28, q)]).astype(np.int64)
        vals = vals[np.abs(vals) < (1 << 29)][:512]
        coef = np.zeros((16, 32), np.int32)
        coef.ravel()[:vals.size] = vals
        qidx = np.full((1, 1), q, np.int32)
        same(ctx.quantise_transform_np(coef, qidx, qm), ref.quantise_np(coef, qidx, qm), "quant q=%d" % q)
    with pytest.raises(vc2.Vc2Error) as e:
        ctx.quantise_transform_np(np.zeros((2, 2), np.int32), np.full((1, 1), 120, np.int32), qm)
    assert "quantization index exceeds maximum implemented value" in str(e.value)


def _quantised_picture(ref, geom, kernel, seed, amp=400, q=None):
    depth = geom.depth
    (ph, pw), (ch, cw) = vc2.api.padded_dims(geom)
    ny, nx = geom.slices_y, geom.slices_x
    qm = vc2.quant_matrix(kernel, depth)
    planes = [rnd((ph, pw), -amp, amp, seed), rnd((ch, cw), -amp, amp, seed + 1), rnd((ch, cw), -amp, amp, seed + 2)]
    # make some slices sparse / all zero and some with long zero tails
    planes[0][: ph // ny, : pw // nx] = 0
    planes[1][: ch // ny] //= 64
    planes[2][:, : cw // 2] //= 16
    qidx = rnd((ny, nx), 0, 40, seed + 3) if q is None else np.full((ny, nx), q, np.int32)
    quant = [ref.quantise_np(p, qidx, qm) for p in planes]
    return planes, quant, qidx, qm


PACK_CASES = [
    # h, w, chroma, kernel, depth, u, a, prefix, scalar
    (64, 128, "422", "LeGall", 3, 1, 2, 0, 1),
    (48, 96, "444", "DD97", 2, 2, 3, 1, 2),
    (71, 87, "420", "Haar1", 1, 2, 2, 2, 1),
    (128, 256, "422", "DD137", 4, 1, 2, 0, 4),
    (96, 96, "444", "Fidelity", 2, 3, 3, 0, 3),
]


@pytest.mark.parametrize("case", PACK_CASES)
def test_hq_pack_unpack_vbr_vs_reference(ctx, ref, case):
    h, w, cf, kernel, depth, u, a, prefix, scalar = case
    g = vc2.make_geom(h, w, cf, kernel, depth, u, a, prefix, scalar)
    planes, quant, qidx, qm = _quantised_picture(ref, g, kernel, 42)
    want = ref.pack_slices(quant[0], quant[1], quant[2], depth, qidx, 0, prefix, scalar)
    got, off = ctx.hq_pack(quant[0], quant[1], quant[2], g, qidx, "HQ_VBR")
    assert len(got) == len(want)
    assert got == want, "first differing byte %d" % next(i for i in range(len(want)) if got[i] != want[i])
    assert off[-1] == len(want)
    assert off.tolist() == vc2.hq_index_slices(want, g.slices_x * g.slices_y, prefix, scalar).tolist()
    y, uu, v, q = ctx.hq_unpack(want, g)
    (ph, pw), (ch, cw) = vc2.api.padded_dims(g)
    ry, ru, rv, rq = ref.unpack_slices(want, ph, pw, ch, cw, depth, g.slices_y, g.slices_x, 0, prefix, scalar)
    same(q, rq, "qindex")
    same(y, ry, "Y")
    same(uu, ru, "U")
    same(v, rv, "V")
    same(y, quant[0], "Y vs source")


@pytest.mark.parametrize("case", PACK_CASES)
def test_cbr_rate_control_and_pack_vs_reference(ctx, ref, case):
    h, w, cf, kernel, depth, u, a, prefix, scalar = case
    g = vc2.make_geom(h, w, cf, kernel, depth, u, a, prefix, scalar)
    planes, _, _, qm = _quantised_picture(ref, g, kernel, 7, amp=900)
    ncoef = sum(p.size for p in planes)
    for ratio in (3, 6):
        total = ncoef * 10 // 8 // ratio
        sb = vc2.slice_bytes(g.slices_y, g.slices_x, total, scalar)
        want_q = ref.cbr_qindices(planes[0], planes[1], planes[2], qm, sb, scalar)
        got_q = ctx.quantIndicesCBR(planes[0], planes[1], planes[2], g, qm, sb)
        same(got_q, want_q, "CBR qindex ratio %d" % ratio)
        quant = [ref.quantise_np(p, want_q, qm) for p in planes]
        want = ref.pack_slices(quant[0], quant[1], quant[2], depth, want_q, 1, prefix, scalar, sb)
        got, off = ctx.hq_pack(quant[0], quant[1], quant[2], g, want_q, "HQ_CBR", sb)
        assert got == want
        y, uu, v, q = ctx.hq_unpack(want, g)
        same(q, want_q, "qindex")
        same(y, quant[0], "Y")
        same(uu, quant[1], "U")
        same(v, quant[2], "V")


def test_slice_scalar_too_small_is_reported(ctx, ref):
    g = vc2.make_geom(64, 128, "444", "LeGall", 2, 8, 16, 0, 1)   # one 32x64 slice per 32x64 block: > 255 bytes
    planes, quant, qidx, qm = _quantised_picture(ref, g, "LeGall", 3, amp=30000, q=0)
    with pytest.raises(Exception) as e_ref:
        ref.pack_slices(quant[0], quant[1], quant[2], 2, qidx, 0, 0, 1)
    with pytest.raises(vc2.Vc2Error) as e:
        ctx.hq_pack(quant[0], quant[1], quant[2], g, qidx, "HQ_VBR")
    assert "Slice scalar is too small" in str(e.value) and "Slice scalar is too small" in str(e_ref.value)


def test_ld_unpack_vs_reference(ctx, ref):
    depth, kernel = 3, "LeGall"
    g = vc2.make_geom(64, 128, "422", kernel, depth, 1, 2)
    (ph, pw), (ch, cw) = vc2.api.padded_dims(g)
    qm = vc2.quant_matrix(kernel, depth)
    planes = [rnd((ph, pw), -60, 60, 1), rnd((ch, cw), -60, 60, 2), rnd((ch, cw), -60, 60, 3)]
    qidx = rnd((g.slices_y, g.slices_x), 16, 30, 4)
    quant = [ref.quantise_ld(p, qidx, qm) for p in planes]
    sb = vc2.slice_bytes(g.slices_y, g.slices_x, 200 * g.slices_y * g.slices_x + 13, 1)
    data = ref.pack_slices(quant[0], quant[1], quant[2], depth, qidx, 2, 0, 1, sb)
    ry, ru, rv, rq = ref.unpack_slices(data, ph, pw, ch, cw, depth, g.slices_y, g.slices_x, 2, 0, 1, sb)
    y, u, v, q = ctx.ld_unpack(data, g, sb)
    same(q, rq, "qindex")
    same(y, ry, "Y")
    same(u, ru, "U")
    same(v, rv, "V")


@pytest.mark.parametrize("case", PACK_CASES[:4])
def test_slice_bits_and_component_bytes_vs_reference(ctx, ref, case):
    """vc2_slice_bits / vc2_hq_slice_sizes against luma_slice_bits, chroma_slice_bits and component_slice_bytes
    (Slices.cpp:51-119) called slice by slice on the compiled reference"""
    h, w, cf, kernel, depth, u, a, prefix, scalar = case
    g = vc2.make_geom(h, w, cf, kernel, depth, u, a, prefix, scalar)
    planes, quant, qidx, qm = _quantised_picture(ref, g, kernel, 11)
    ny, nx = g.slices_y, g.slices_x
    got_y = ctx.slice_bits(quant[0], None, depth, ny, nx)
    got_uv = ctx.slice_bits(quant[1], quant[2], depth, ny, nx)
    scalar_b = max(scalar, 4)
    got_bytes = ctx.component_slice_bytes(quant[0], depth, ny, nx, scalar_b)
    sh, sw = quant[0].shape[0] // ny, quant[0].shape[1] // nx
    ch, cw = quant[1].shape[0] // ny, quant[1].shape[1] // nx
    for sy in range(ny):
        for sx in range(nx):
            ys = quant[0][sy * sh:(sy + 1) * sh, sx * sw:(sx + 1) * sw]
            us = quant[1][sy * ch:(sy + 1) * ch, sx * cw:(sx + 1) * cw]
            vs = quant[2][sy * ch:(sy + 1) * ch, sx * cw:(sx + 1) * cw]
            assert got_y[sy, sx] == ref.slice_bits(ys, None, depth), (sy, sx)
            assert got_uv[sy, sx] == ref.slice_bits(us, vs, depth), (sy, sx)
            assert got_bytes[sy, sx] == ref.component_slice_bytes(ys, depth, scalar_b), (sy, sx)


def test_component_slice_bytes_scalar_too_small(ctx, ref):
    q = rnd((32, 64), -30000, 30000, 5)
    with pytest.raises(vc2.Vc2Error) as e:
        ctx.component_slice_bytes(q, 2, 1, 1, 1)
    assert "Slice scalar is too small" in str(e.value)


@pytest.mark.parametrize("depth,kernel,cf", [(3, "LeGall", "422"), (2, "DD97", "420"), (4, "Haar1", "444")])
def test_ld_quantise_and_pack_vs_reference(ctx, ref, depth, kernel, cf):
    """vc2_quantise_ld (quantise_transform with DC prediction) and vc2_ld_pack (the LD slice writer) against the reference"""
    g = vc2.make_geom(96, 192, cf, kernel, depth, 2 if cf == "420" else 1, 2 if cf != "444" else 1)
    (ph, pw), (ch, cw) = vc2.api.padded_dims(g)
    qm = vc2.quant_matrix(kernel, depth)
    planes = [rnd((ph, pw), -300, 300, 21), rnd((ch, cw), -300, 300, 22), rnd((ch, cw), -300, 300, 23)]
    for p in planes:       # a DC level, so that the prediction matters
        p[::1 << depth, ::1 << depth] += 700
    qidx = rnd((g.slices_y, g.slices_x), 10, 34, 24)
    want_q = [ref.quantise_ld(p, qidx, qm) for p in planes]
    got_q = [ctx.quantise_transform(p, qidx, qm) for p in planes]
    for a, b, n in zip(got_q, want_q, "YUV"):
        same(a, b, "quantise_transform " + n)
    same(ctx.inverse_quantise_transform(want_q[0], qidx, qm), ref.dequantise_ld(want_q[0], qidx, qm), "inverse")
    n = g.slices_y * g.slices_x
    sb = vc2.slice_bytes(g.slices_y, g.slices_x, 2 * sum(p.size for p in planes) + 7, 1)   # ample: two bytes per coefficient
    want = ref.pack_slices(want_q[0], want_q[1], want_q[2], depth, qidx, 2, 0, 1, sb)
    got = ctx.ld_pack(want_q[0], want_q[1], want_q[2], g, qidx, sb)
    assert got == want
    # a budget the chroma does not fit: the reference throws, so does the C-ABI
    tight = np.full((g.slices_y, g.slices_x), 12, np.int32)
    with pytest.raises(Exception) as e_ref:
        ref.pack_slices(want_q[0], want_q[1], want_q[2], depth, qidx, 2, 0, 1, tight)
    with pytest.raises(vc2.Vc2Error) as e:
        ctx.ld_pack(want_q[0], want_q[1], want_q[2], g, qidx, tight)
    assert "Too many bytes" in str(e_ref.value) and "Too many bytes" in str(e.value)


def test_cpp_library_mirror_vs_reference(ref):
    """every function of the C++ Library mirror (include/vc2/*.h) against the compiled reference, for all seven kernels:
    the test program is C++ (tests/cpp/test_library_mirror.cpp), its output lists each comparison"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "vc2_reference_b200", "bin", "test_library_mirror")
    r = subprocess.run([exe, ref.PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    out = r.stdout.decode()
    assert r.returncode == 0, [l for l in out.splitlines() if not l.endswith(" same")]
    assert out.count(" same") >= 7 * 18 and "DIFFERENT" not in out


@pytest.mark.parametrize("switch", ["VC2_SEARCH_SMEM", "VC2_SEARCH_WARP", "VC2_SEARCH_SUB"])
def test_optional_rate_control_kernels_are_bit_exact(switch):
    """the optional rate-control kernels (slices in shared memory / a warp per slice / eight lanes per slice with the slice in registers; all measured
    slower and off by default, the switch is read once per process): the CBR parity tests again in a process that has one switched on"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_library.py"), "-q", "-m", "gpu", "-k", "cbr_rate_control_and_pack"],
                       env=dict(os.environ, **{switch: "1"}), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, cwd=root)
    assert r.returncode == 0, r.stdout.decode()[-2000:]

"""CPU-only: host-side helpers of the C-ABI against the reference (live taps) and its known answers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import vc2_reference_b200 as vc2
from vc2_reference_b200._cabi import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# quantMatrix known answers dumped from the compiled reference
QM_KAT = {
    ("DD97", 4): [5, 3, 3, 0, 4, 4, 1, 5, 5, 2, 6, 6, 3],
    ("LeGall", 4): [4, 2, 2, 0, 4, 4, 2, 5, 5, 3, 7, 7, 5],
    ("DD137", 4): [5, 3, 3, 0, 4, 4, 1, 5, 5, 2, 6, 6, 3],
    ("Haar1", 4): [8, 4, 4, 0, 4, 4, 0, 4, 4, 0, 4, 4, 0],
    ("Haar0", 4): [20, 16, 16, 12, 12, 12, 8, 8, 8, 4, 4, 4, 0],
    ("Fidelity", 4): [0, 4, 4, 8, 8, 8, 12, 13, 13, 17, 17, 17, 21],
    ("Daub97", 4): [3, 1, 1, 0, 4, 4, 2, 6, 6, 5, 9, 9, 7],
}


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vc2_cabi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(vc2_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 40
    raw = C.CDLL(vc2.lib_path)
    missing = [n for n in sorted(names) if not hasattr(raw, n)]
    assert not missing, missing
    # and the ctypes binding declares all of them
    assert set(lib._vc2_symbols) == names


def test_quant_matrix_known_answers():
    for (k, d), want in QM_KAT.items():
        assert vc2.quant_matrix(k, d).tolist() == want
    assert vc2.quant_matrix("LeGall", 0).tolist() == [0]


def test_quant_matrix_vs_reference(ref):
    for k in range(7):
        for d in range(0, 7):
            assert vc2.quant_matrix(k, d).tolist() == ref.quant_matrix(k, d).tolist(), (k, d)


def test_quant_factor_table_vs_reference(ref):
    for q in range(-3, 116):
        assert lib.vc2_quant_factor(q) == ref.quant_factor(q), q
        assert lib.vc2_quant_offset(q) == ref.quant_offset(q), q
    # reference unit test (tests/Quantisation.cpp:6-12): index 130 is beyond the table
    assert ref.quant_factor(130) == -1


def test_quant_magic31_is_exact_below_2_31():
    """The slice coders divide by quant_factor with one multiply-high and one shift (cabi.cu vc2_quant_magic31):
    it must equal truncating division for EVERY dividend the reference can form, (abs(v) << 2) in int
    (Quantisation.cpp:69-76), i.e. below 2^31.  Checked on every dividend below 2^22, on the multiples of the
    divisor and their neighbours (where a rounding error would show first) up to 2^31, and on random dividends."""
    import ctypes as C
    rng = np.random.default_rng(31)
    dense = np.arange(1 << 22, dtype=np.uint64)
    rnd = rng.integers(0, 1 << 31, 1 << 20, dtype=np.uint64)
    top = np.uint64((1 << 31) - 1)
    for q in range(120):
        m, sh = C.c_uint32(), C.c_uint32()
        assert lib.vc2_quant_magic31(q, C.byref(m), C.byref(sh)) == 0
        d = np.uint64(lib.vc2_quant_factor(q) & 0xFFFFFFFF)
        mult = np.arange(1, 1 << 16, dtype=np.uint64) * np.uint64(max(1, ((1 << 31) // int(d)) >> 16)) * d
        mult = mult[mult <= top]
        edge = np.concatenate([mult, mult - np.uint64(1), np.minimum(mult + np.uint64(1), top), np.array([top, top - min(d, top), 0], np.uint64)])
        for a in (dense, rnd, edge):
            got = ((a * np.uint64(m.value)) >> np.uint64(32)) >> np.uint64(sh.value)
            assert np.array_equal(got, a // d), q
    assert lib.vc2_quant_magic31(120, C.byref(m), C.byref(sh)) != 0


def test_slice_bytes(ref):
    assert vc2.slice_bytes(3, 4, 1000, 1).ravel().tolist() == [83, 83, 84, 83, 83, 84, 83, 83, 84, 83, 83, 84]
    assert (vc2.slice_bytes(135, 120, 2073600, 1) == 128).all()
    for ny, nx, total, scalar in [(3, 4, 1000, 1), (135, 120, 2073600, 1), (7, 5, 9999, 3), (17, 11, 123457, 4), (2, 2, 64, 8)]:
        assert (vc2.slice_bytes(ny, nx, total, scalar) == ref.slice_bytes(ny, nx, total, scalar)).all()


def test_padding_and_slice_validity(ref):
    for size in (1, 7, 8, 1080, 1081, 2160, 4320):
        for d in range(1, 6):
            assert vc2.padded_size(size, d) == ref.padded_size(size, d)
    for d in range(1, 6):
        for luma, chroma in ((1080, 1080), (1920, 960), (188, 94), (174, 87), (352, 176), (100, 100)):
            for n in range(0, 12):
                assert lib.vc2_slice_size_is_valid(d, luma, chroma, n) == ref.slice_size_is_valid(d, luma, chroma, n)


def test_make_geom_rejects_invalid_slicing():
    g = vc2.make_geom(1080, 1920, "422", "LeGall", 3, 1, 2)
    assert (g.slices_y, g.slices_x, g.chroma_w) == (135, 120, 960)
    with pytest.raises(vc2.Vc2Error):
        vc2.make_geom(120, 200, "422", "LeGall", 3, 1, 2)


def test_hq_index_slices():
    # two slices, prefix 1, scalar 2: [pfx q ly Y.. lu U.. lv V..]
    s0 = bytes([9, 5, 1]) + b"ab" + bytes([0]) + bytes([2]) + b"cdef"
    s1 = bytes([9, 6, 0, 0, 0])
    off = vc2.hq_index_slices(s0 + s1, 2, 1, 2)
    assert off.tolist() == [0, len(s0), len(s0) + len(s1)]
    with pytest.raises(vc2.Vc2Error):
        vc2.hq_index_slices((s0 + s1)[:-1], 2, 1, 2)


def test_no_cpu_fallback_without_gpu():
    if lib.vc2_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vc2.Vc2Error):
        vc2.Context(0)


def test_command_line_pipeline_pieces(tmp_path):
    """host/pipeline.h without a GPU: the positional writer (multi-threaded pwrite, awkward piece sizes), the host buffer's
    fallback to ordinary memory, the round queue (tests/cpp/test_pipeline_host.cpp)"""
    import subprocess
    exe = os.path.join(ROOT, "vc2_reference_b200", "bin", "test_pipeline_host")
    r = subprocess.run([exe, str(tmp_path / "out.bin")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()


@pytest.mark.parametrize("args", [
    ["-x", "100", "-y", "60", "-f", "4:2:2", "-d", "3", "-u", "1", "-a", "1"],      # depth impossible: suggests depth and slice sizes
    ["-x", "1920", "-y", "1080", "-f", "4:2:2", "-d", "3", "-u", "5", "-a", "7"],   # slice sizes that do not divide: suggests others
    ["-x", "176", "-y", "144", "-f", "4:2:0", "-d", "2", "-u", "1", "-a", "1"],     # 4:2:0 needs even slice sizes
    ["-x", "64", "-y", "34", "-f", "4:4:4", "-d", "4", "-u", "3", "-a", "3"],
])
def test_parameter_advice_matches_reference(tmp_path, args):
    """bad wavelet depth / slice sizes: the advice on standard error (suggestWaveletDepth, suggestSliceSize,
    waveletTransformIsPossible, EncodeStream.cpp:379-405) and the error text are the reference's, for the drop-in command line and for
    the reference's own main over the drop-in Library bodies (no GPU is needed: nothing is transformed)"""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref", "EncodeStream")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    src = str(tmp_path / "in.yuv")
    open(src, "wb").write(b"\\0" * 200000)
    full = ["-m", "HQ_ConstQ", "-z", "10", "-k", "LeGall", "-q", "5"] + args + [src, str(tmp_path / "out")]

    def advice(exe):
        r = subprocess.run([exe] + full, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        lines = [l for l in r.stderr.decode().splitlines() if l.startswith("Consider") or l.startswith("It is not possible")]
        return r.returncode != 0, lines, r.stdout.decode().strip()
    want = advice(ref)
    assert want[0] and want[1], "the reference accepts these parameters"
    assert advice(os.path.join(ROOT, "vc2_reference_b200", "bin", "EncodeStream")) == want
    dropin = os.path.join(ROOT, "tests", "dropin", "_build", "EncodeStream")
    if os.path.exists(dropin):
        assert advice(dropin) == want


def test_vlc_value_classes_vs_reference():
    """include/vc2/VLC.h (UnsignedVLC / SignedVLC value classes, VLC.h:17-46) against SignedVLC of the compiled reference for 7000 values,
    and round trips through the MSB-first bit buffer (tests/cpp/test_vlc_host.cpp)"""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref", "libvc2ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([os.path.join(ROOT, "vc2_reference_b200", "bin", "test_vlc_host"), ref], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()

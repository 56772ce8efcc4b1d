"""Source-level drop-in: the reference's UNMODIFIED EncodeStream.cpp / DecodeStream.cpp, linked by
tests/dropin/build.sh against host/dropin/{WaveletTransform,Quantisation,Slices}.cpp (the replacement Library bodies
over the CUDA C-ABI) instead of the reference's own three files, must write the reference's bytes.
Golden digests: tests/golden/md5.json, made by the unmodified reference with the same flags."""
import hashlib
import json
import os
import subprocess

import pytest

import gen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "dropin", "_build")
GOLD = json.load(open(os.path.join(HERE, "golden", "md5.json")))
FMT = {"444": "4:4:4", "422": "4:2:2", "420": "4:2:0"}
built = pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "EncodeStream")),
                           reason="tests/dropin/_build not built (run tests/dropin/build.sh where /root/reference exists)")


def enc_args(c):
    a = ["-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", FMT[c["fmt"]], "-z", str(c["bits"]),
         "-k", c["kernel"], "-d", str(c["wdepth"]), "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"])]
    a += ["-q", str(c["q"])] if c["mode"] == "HQ_ConstQ" else ["-s", str(c["s"])]
    if c["mode"] != "LD":
        a += ["-S", str(c["S"]), "-P", str(c["P"])]
    return a + list(c.get("extra", []))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@built
def test_dropin_links_the_cuda_library_and_not_the_reference_bodies():
    """the binaries resolve the hot Library functions from host/dropin (over libvc2b200.so): the C-ABI entry points are
    undefined symbols of the executable, the lifting / slice coding internals of the reference are absent"""
    for exe in ("EncodeStream", "DecodeStream"):
        path = os.path.join(BUILD, exe)
        needed = subprocess.run(["readelf", "-d", path], stdout=subprocess.PIPE).stdout.decode()
        assert "libvc2b200.so" in needed
        syms = subprocess.run(["nm", "-C", path], stdout=subprocess.PIPE).stdout.decode()
        undefined = [l.split()[-1] for l in syms.splitlines() if " U " in l]
        for want in ("vc2_dwt_forward", "vc2_dwt_inverse", "vc2_quantise_np", "vc2_dequantise_np", "vc2_hq_pack", "vc2_hq_unpack",
                     "vc2_ld_pack", "vc2_ld_unpack", "vc2_slice_bits", "vc2_hq_slice_sizes"):
            assert want in undefined, (exe, want)
        for absent in ("waveletLevelDD97", "waveletLevelLeGall", "HQSliceIO_VBR", "LDSliceIO", "quantise_subbands"):
            assert absent not in syms, (exe, absent)


@built
def test_dropin_has_no_cpu_fallback(tmp_path):
    """without a CUDA device the drop-in fails with the library's error; it never computes on the host"""
    import ctypes
    try:
        have_gpu = ctypes.CDLL(os.path.join(ROOT, "vc2_reference_b200", "libvc2b200.so")).vc2_device_count() > 0
    except OSError:
        have_gpu = False
    if have_gpu:
        pytest.skip("a CUDA device is present")
    c = GOLD["S07_LeGall_d1_444"]["params"]
    src = str(tmp_path / "in.yuv")
    open(src, "wb").write(gen.frame_bytes(c["seed"], 0, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"]))
    r = subprocess.run([os.path.join(BUILD, "EncodeStream")] + enc_args(c) + [src, str(tmp_path / "out")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    assert b"no usable CUDA device" in r.stdout + r.stderr


@built
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["S01_LeGall_d3_422", "S05_Fidelity_d2_422", "S08_DD137_d4_422", "D5_DD97_d5_422",   # HQ_ConstQ
                                  "B00_DD97_d2_420", "B06_Daub97_d3_444",                                           # HQ_CBR: the reference's own search loop
                                  "F01_LeGall_d2_420_small", "I00_LeGall_d3_422_tff",                               # fragments, interlace
                                  "L00_LeGall_d3_420", "L06_DD97_d3_422_frag",                                      # LD encoder and decoder
                                  "C1"])                                                                              # 1080p
def test_unmodified_reference_mains_over_the_cabi(tmp_path, name):
    c, taps = GOLD[name]["params"], GOLD[name]["taps"]
    src = str(tmp_path / "in.yuv")
    with open(src, "wb") as f:
        for i in range(c["frames"]):
            f.write(gen.frame_bytes(c["seed"], i, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"]))
    small = not name.startswith("C")
    for tap in ["Stream"] + (["Transform", "Quantised", "Packaged"] if small else []):
        dst = str(tmp_path / ("enc_" + tap))
        r = subprocess.run([os.path.join(BUILD, "EncodeStream")] + enc_args(c) + ["-o", tap, src, dst], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, (tap, r.stdout[-300:], r.stderr[-300:])
        assert md5(dst) == taps["enc_" + tap]["md5"], (name, "enc", tap)
    stream = str(tmp_path / "enc_Stream")
    for tap in ["Decoded"] + (["Quantised", "Transform"] if small else []):
        if "md5" not in taps.get("dec_" + tap, {}):
            continue     # the reference's decoder rejects this stream itself (interlaced LD)
        dst = str(tmp_path / ("dec_" + tap))
        r = subprocess.run([os.path.join(BUILD, "DecodeStream"), "-o", tap, stream, dst], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, (tap, r.stdout[-300:], r.stderr[-300:])
        assert md5(dst) == taps["dec_" + tap]["md5"], (name, "dec", tap)

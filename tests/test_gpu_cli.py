"""GPU: the drop-in command lines (vc2_reference_b200/bin/EncodeStream, DecodeStream; host/*.cpp) run with the
reference's own flags and must produce the reference's bytes: stream, intermediate taps, decoded pictures
(golden digests in tests/golden/md5.json were made by the unmodified reference with the same flags)."""
import hashlib
import json
import os
import subprocess

import pytest

import gen

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "vc2_reference_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = json.load(open(os.path.join(HERE, "golden", "md5.json")))
FMT = {"444": "4:4:4", "422": "4:2:2", "420": "4:2:0"}


def enc_args(c):
    a = ["-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", FMT[c["fmt"]], "-z", str(c["bits"]),
         "-k", c["kernel"], "-d", str(c["wdepth"]), "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"])]
    a += ["-q", str(c["q"])] if c["mode"] == "HQ_ConstQ" else ["-s", str(c["s"])]
    if c["mode"] != "LD":
        a += ["-S", str(c["S"]), "-P", str(c["P"])]
    if c.get("nbytes"):
        a += ["-n", str(c["nbytes"])]
    return a + list(c.get("extra", []))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def widen(raw, nbytes):
    """16-bit big-endian MSB-justified words -> nbytes-wide ones (-n 3 / 4 input files)"""
    import numpy as np
    a = np.frombuffer(raw, np.uint8).reshape(-1, 2)
    out = np.zeros((a.shape[0], nbytes), np.uint8)
    out[:, :2] = a
    return out.tobytes()


def write_input(c, path):
    with open(path, "wb") as f:
        for i in range(c["frames"]):
            raw = gen.frame_bytes(c["seed"], i, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"])
            f.write(widen(raw, c["nbytes"]) if c.get("nbytes") else raw)


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, (cmd, r.stdout.decode(errors="replace")[-400:], r.stderr.decode(errors="replace")[-400:])
    return r


@pytest.mark.parametrize("name", ["S01_LeGall_d3_422", "S04_Haar1_d4_420", "S05_Fidelity_d2_422", "S08_DD137_d4_422",
                                  "B00_DD97_d2_420", "B06_Daub97_d3_444", "C1", "C2",
                                  # wavelet depths 5 and 6 (quantMatrix is computed for any depth, WaveletTransform.cpp:345-423)
                                  "D5_DD97_d5_422", "D6_LeGall_d6_444",
                                  # input words of 4 and 3 bytes (-n, Arrays.cpp:333-379)
                                  "N4_LeGall_d2_422", "N3_Haar1_d2_444_12b",
                                  # full-range noise at index 0: the payload is larger than the raw picture, the command line's
                                  # payload buffers (raw size to begin with) have to grow and the batch is encoded again
                                  "R00_noise_q0_444",
                                  # interlaced coding (two field pictures per frame) and fragmented pictures
                                  "I00_LeGall_d3_422_tff", "I01_DD137_d2_420_bff", "I02_Haar1_d3_444_tff",
                                  "F00_DD97_d3_422", "F01_LeGall_d2_420_small", "F02_Fidelity_d2_422_il",
                                  # the LD encoder; the reference's own decoder rejects its interlaced LD streams,
                                  # so those cases pin the encoder side only
                                  "L00_LeGall_d3_420", "L02_Haar1_d3_444", "L04_Fidelity_d2_422_il_frag", "L05_LeGall_d2_420_il",
                                  "L06_DD97_d3_422_frag"])
def test_command_lines_vs_golden(tmp_path, name):
    c, taps = GOLD[name]["params"], GOLD[name]["taps"]
    src = str(tmp_path / "in.yuv")
    write_input(c, src)
    # a batch smaller than the clip and (when there are several GPUs) two devices: chunking and reassembly
    extra = ["-B", "1", "-G", "2"] if name[0] in "SIDN" else ["-B", "3"] if name[0] in "FL" else []
    small = not name.startswith("C")       # the 1080p configs: stream and pictures only (each run pays a CUDA start-up)
    for tap in ["Stream"] + (["Packaged", "Transform", "Quantised"] if small else []) + (["Indices"] if c["mode"] != "HQ_ConstQ" else []):
        dst = str(tmp_path / ("enc_" + tap))
        run([os.path.join(BIN, "EncodeStream")] + enc_args(c) + extra + ["-o", tap, src, dst])
        assert md5(dst) == taps["enc_" + tap]["md5"], (name, tap)
    stream = str(tmp_path / "enc_Stream")
    for tap in ["Decoded"] + (["Transform", "Quantised", "Indices"] if small else []):
        if "md5" not in taps["dec_" + tap]:
            continue          # the reference decoder fails on this stream (interlaced LD): nothing to compare with
        dst = str(tmp_path / ("dec_" + tap))
        run([os.path.join(BIN, "DecodeStream")] + extra + ["-o", tap, stream, dst])
        assert md5(dst) == taps["dec_" + tap]["md5"], (name, tap)
    # EncodeStream -o PSNR: the reference's per-frame text report (quantiser statistics + PSNR of the local decode)
    if "enc_PSNR" in taps:
        dst = str(tmp_path / "enc_PSNR")
        run([os.path.join(BIN, "EncodeStream")] + enc_args(c) + extra + ["-o", "PSNR", src, dst])
        assert open(dst).read() == taps["enc_PSNR"]["text"], (name, "PSNR")
    # EncodeStream -o Decoded = the decoder's picture (local decode loop, EncodeStream.cpp:649-767)
    if not small or "md5" not in taps["dec_Decoded"]:
        return
    dst = str(tmp_path / "enc_Decoded")
    run([os.path.join(BIN, "EncodeStream")] + enc_args(c) + ["-o", "Decoded", src, dst])
    if c.get("nbytes"):    # the encoder's local decode keeps the input's word width: the decoder's 2-byte words, widened
        dec = open(str(tmp_path / "dec_Decoded"), "rb").read()
        assert open(dst, "rb").read() == widen(dec, c["nbytes"])
    elif c["bits"] != 8:   # the stand-alone decoder writes 8-bit streams as one byte per sample, the encoder keeps -n
        assert md5(dst) == taps["dec_Decoded"]["md5"]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "EncodeStream")), reason="oracle/_ref not built")
def test_decode_ld_stream_from_reference_encoder(tmp_path):
    """LD is decode-only here: the stream comes from the reference's encoder, both decoders must agree"""
    c = dict(w=352, h=288, fmt="420", bits=8, frames=3, seed=77, smooth=False)
    src = str(tmp_path / "in.yuv")
    write_input(c, src)
    stream = str(tmp_path / "ld.vc2")
    run([os.path.join(REF, "EncodeStream"), "-m", "LD", "-x", "352", "-y", "288", "-f", "4:2:0", "-z", "8", "-k", "LeGall", "-d", "3",
         "-u", "2", "-a", "2", "-r", "3", "-s", "40000", src, stream])
    for tap in ["Decoded", "Transform", "Quantised", "Indices"]:
        a, b = str(tmp_path / ("ref_" + tap)), str(tmp_path / ("our_" + tap))
        run([os.path.join(REF, "DecodeStream"), "-o", tap, stream, a])
        run([os.path.join(BIN, "DecodeStream"), "-o", tap, stream, b])
        assert md5(a) == md5(b), tap


def test_error_reporting_matches_reference(tmp_path):
    """bad parameters: 'Error: <reference text>' and a failure status (EncodeStream.cpp:250-256, 782-785)"""
    src = str(tmp_path / "in.yuv")
    open(src, "wb").write(b"\0" * 64)
    r = subprocess.run([os.path.join(BIN, "EncodeStream"), "-m", "HQ_ConstQ", "-x", "64", "-y", "32", "-f", "4:2:2", "-k", "LeGall", "-d", "2",
                        "-u", "1", "-a", "1", src, str(tmp_path / "o")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Quantisation index must be set in HQ_ConstQ mode" in r.stderr
    # a slice scalar that is too small for the content is the reference's logic_error, on standard output
    c = GOLD["C3"]["params"]
    one = dict(c, frames=1)
    big = str(tmp_path / "c3.yuv")
    write_input(one, big)
    a = enc_args(one)
    a[a.index("-S") + 1] = "1"
    r = subprocess.run([os.path.join(BIN, "EncodeStream")] + a + [big, str(tmp_path / "o2")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Error: Slice scalar is too small" in r.stdout, r.stdout[-300:]


def frame_args(c):
    a = ["-x", str(c["w"]), "-y", str(c["h"]), "-f", FMT[c["fmt"]], "-z", str(c["bits"]), "-k", c["kernel"], "-d", str(c["wdepth"]),
         "-u", str(c["u"]), "-a", str(c["a"]), "-m", "HQ", "-S", str(c["S"]), "-P", str(c["P"])]
    return a + [e for e in c.get("extra", []) if e in ("-i", "-b", "-t")]


@pytest.mark.parametrize("name", ["S01_LeGall_d3_422", "S05_Fidelity_d2_422", "B00_DD97_d2_420", "I00_LeGall_d3_422_tff", "I01_DD137_d2_420_bff"])
def test_decode_frame_vs_golden(tmp_path, name):
    """DecodeFrame (bare slice data, parameters on the command line; src/DecodeFrame) against digests made by the reference tool,
    including its habit of writing a zero frame behind every frame of -o Transform / Quantised / Indices output"""
    c, taps = GOLD[name]["params"], GOLD[name]["taps"]
    src, pk = str(tmp_path / "in.yuv"), str(tmp_path / "pk")
    write_input(c, src)
    run([os.path.join(BIN, "EncodeStream")] + enc_args(c) + ["-o", "Packaged", src, pk])
    assert md5(pk) == taps["enc_Packaged"]["md5"]
    for tap in ["Decoded", "Transform", "Quantised", "Indices"]:
        dst = str(tmp_path / ("f_" + tap))
        run([os.path.join(BIN, "DecodeFrame")] + frame_args(c) + ["-B", "3", "-o", tap, pk, dst])
        assert md5(dst) == taps["frame_" + tap]["md5"], (name, tap)
    # LD input: the reference tool fails on the first picture, and so does this one
    base = frame_args(c)
    i = base.index("-m")
    r = subprocess.run([os.path.join(BIN, "DecodeFrame"), "-m", "LD", "-s", "1000"] + base[:i] + base[i + 6:] + [pk, str(tmp_path / "x")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Failed to read the first compressed frame" in r.stderr

#!/usr/bin/env python3
"""Generate tests/golden/md5.json by RUNNING THE REFERENCE (oracle/_ref) here.

Runs in the build container only (needs oracle/_ref built from /root/reference
by oracle/build_ref.sh).  For every case it feeds the deterministic synthetic
clip of oracle/gen.py to the reference EncodeStream / DecodeStream and records
md5 + size of each tap (-o Transform/Quantised/Indices/Packaged/Stream, decoder
-o Quantised/Transform/Decoded).  The GPU parity tests regenerate the same clip
and must reproduce these digests bit for bit.

usage: make_golden.py [--only NAME ...] [--jobs N]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
FMT = {"444": "4:4:4", "422": "4:2:2", "420": "4:2:0"}


def case(name, w, h, fmt, depth_bits, frames, mode, kernel, wdepth, u, a, **kw):
    d = dict(name=name, w=w, h=h, fmt=fmt, bits=depth_bits, frames=frames, mode=mode, kernel=kernel,
             wdepth=wdepth, u=u, a=a, q=None, s=None, S=1, P=0, r=3, seed=1234, smooth=False, extra=[])
    d.update(kw)
    return d


CASES = [
    # the five BASELINE.json configs (frame counts bounded)
    case("C1", 1920, 1080, "422", 10, 2, "HQ_ConstQ", "LeGall", 3, 1, 2, q=12),
    case("C2", 1920, 1080, "422", 10, 2, "HQ_CBR", "DD97", 3, 1, 2, s=2073600, r=6),
    case("C3", 3840, 2160, "422", 10, 2, "HQ_ConstQ", "DD137", 4, 1, 2, q=16, r=6, S=4),
    case("C4a", 7680, 4320, "444", 12, 1, "HQ_ConstQ", "Haar1", 2, 4, 8, q=20, r=6, S=8),
    case("C4b", 7680, 4320, "444", 12, 1, "HQ_ConstQ", "Fidelity", 4, 1, 2, q=20, r=6, S=8),
    case("C5", 1920, 1080, "422", 10, 2, "LD", "LeGall", 3, 1, 2, s=2073600, r=6),
]
# small cases: every kernel, several depths / chroma formats, sizes that need padding
_small = [
    ("DD97", 2, "420", 174, 118, 8), ("LeGall", 3, "422", 188, 116, 10), ("DD137", 2, "444", 100, 60, 10),
    ("Haar0", 3, "422", 208, 104, 10), ("Haar1", 4, "420", 352, 288, 8), ("Fidelity", 2, "422", 144, 88, 12),
    ("Daub97", 3, "444", 120, 72, 10), ("LeGall", 1, "444", 66, 34, 10), ("DD137", 4, "422", 352, 240, 10),
    ("Fidelity", 3, "444", 256, 128, 12), ("Daub97", 2, "420", 182, 102, 10), ("DD97", 4, "422", 640, 368, 10),
]
for i, (k, d, f, w, h, b) in enumerate(_small):
    CASES.append(case("S%02d_%s_d%d_%s" % (i, k, d, f), w, h, f, b, 2, "HQ_ConstQ", k, d, 2 if f == "420" else 1, 2 if f != "444" else 1,
                      q=8 + 2 * i, S=1 + (i % 3), P=i % 2, seed=100 + i))
    if i % 3 == 0:
        ch, cw = gen.chroma_dims(w, h, f)
        CASES.append(case("B%02d_%s_d%d_%s" % (i, k, d, f), w, h, f, b, 2, "HQ_CBR", k, d, 2 if f == "420" else 1, 2 if f != "444" else 1,
                          s=(w * h + 2 * ch * cw) * b // 8 // 3, S=1 + (i % 2), seed=200 + i))


# interlaced coding (-i, field order -t / -b) and fragmented pictures (-F, HQ_CBR only); command-line tests only
CASES += [
    case("I00_LeGall_d3_422_tff", 188, 116, "422", 10, 3, "HQ_ConstQ", "LeGall", 3, 1, 2, q=10, S=2, P=1, seed=300, extra=["-i"]),
    case("I01_DD137_d2_420_bff", 176, 144, "420", 8, 2, "HQ_CBR", "DD137", 2, 2, 2, s=20000, seed=301, extra=["-i", "-b"]),
    case("I02_Haar1_d3_444_tff", 128, 96, "444", 12, 2, "HQ_CBR", "Haar1", 3, 1, 1, s=30000, S=2, seed=302, extra=["-i", "-t"]),
    case("F00_DD97_d3_422", 352, 240, "422", 10, 2, "HQ_CBR", "DD97", 3, 1, 2, s=60000, seed=310, extra=["-F", "1400"]),
    case("F01_LeGall_d2_420_small", 176, 144, "420", 8, 2, "HQ_CBR", "LeGall", 2, 2, 2, s=12000, P=2, seed=311, extra=["-F", "50"]),
    case("F02_Fidelity_d2_422_il", 144, 88, "422", 12, 3, "HQ_CBR", "Fidelity", 2, 1, 2, s=16000, S=2, seed=312, extra=["-F", "700", "-i"]),
]


# the LD encoder (rate control with DC prediction, LD slice writer, LD picture / fragment data units)
CASES += [
    case("L00_LeGall_d3_420", 352, 288, "420", 8, 2, "LD", "LeGall", 3, 2, 2, s=40000, seed=400),
    case("L01_DD97_d3_422_pad", 188, 116, "422", 10, 2, "LD", "DD97", 3, 1, 2, s=30000, seed=401),
    case("L02_Haar1_d3_444", 128, 96, "444", 12, 2, "LD", "Haar1", 3, 1, 1, s=20000, seed=402),
    case("L03_DD137_d2_422_lowrate", 176, 144, "422", 10, 2, "LD", "DD137", 2, 1, 2, s=6000, seed=403),
    case("L05_LeGall_d2_420_il", 176, 144, "420", 8, 2, "LD", "LeGall", 2, 2, 2, s=24000, seed=405, extra=["-i", "-b"]),
    case("L06_DD97_d3_422_frag", 352, 240, "422", 10, 2, "LD", "DD97", 3, 1, 2, s=60000, seed=406, extra=["-F", "1400"]),
    case("L04_Fidelity_d2_422_il_frag", 144, 88, "422", 12, 3, "LD", "Fidelity", 2, 1, 2, s=16000, seed=404, extra=["-F", "700", "-i"]),
]


# full-width pictures for every kernel family the BASELINE configs do not cover: the lifting kernels' fast loop only
# runs on strips that lie wholly inside the picture (>= 3 strips of 240 columns)
CASES += [
    case("W00_Daub97_d3_422_1080p", 1920, 1080, "422", 10, 1, "HQ_ConstQ", "Daub97", 3, 1, 2, q=10, S=2, seed=500),
    case("W01_Haar0_d4_422_1080p", 1920, 1080, "422", 10, 1, "HQ_ConstQ", "Haar0", 4, 1, 2, q=6, S=2, seed=501),
    case("W02_LeGall_d2_444_720p", 1280, 720, "444", 12, 1, "HQ_CBR", "LeGall", 2, 1, 1, s=900000, seed=502),
    case("W03_DD97_d3_420_1024p", 1920, 1024, "420", 8, 1, "HQ_ConstQ", "DD97", 3, 2, 2, q=14, P=1, seed=503),
    case("W04_Haar1_d3_422_odd", 1450, 810, "422", 10, 1, "HQ_ConstQ", "Haar1", 3, 1, 2, q=8, S=2, seed=504),
]

# wavelet depths beyond the four that the BASELINE configs use: the reference computes quantMatrix for any depth >= 0
# (WaveletTransform.cpp:345-423) and the kernels here take 1..6
CASES += [
    case("D5_DD97_d5_422", 640, 384, "422", 10, 2, "HQ_ConstQ", "DD97", 5, 1, 2, q=9, S=16, seed=600),
    case("D6_LeGall_d6_444", 512, 256, "444", 10, 2, "HQ_CBR", "LeGall", 6, 1, 1, s=200000, S=32, seed=601),
    case("D5_Haar0_d5_422_pad", 640, 330, "422", 8, 2, "HQ_ConstQ", "Haar0", 5, 1, 2, q=5, S=16, seed=602),
]

# sample words of 3 and 4 bytes in the input file (-n, Arrays.cpp:333-379): same pictures, wider words
CASES += [
    case("N4_LeGall_d2_422", 176, 96, "422", 10, 2, "HQ_ConstQ", "LeGall", 2, 1, 2, q=7, seed=700, nbytes=4),
    case("N3_Haar1_d2_444_12b", 128, 64, "444", 12, 2, "HQ_CBR", "Haar1", 2, 2, 2, s=9000, seed=701, nbytes=3),
]

# an incompressible picture: full-range noise at index 0 codes to more bytes than the raw picture has (the command line's payload
# buffers start at the raw size and must grow)
CASES += [
    case("R00_noise_q0_444", 1024, 512, "444", 10, 2, "HQ_ConstQ", "LeGall", 2, 2, 4, q=0, S=8, seed=800, smooth="noise"),
]


def widen(raw, nbytes):
    """16-bit big-endian MSB-justified words -> nbytes-wide ones (the sample stays MSB justified)"""
    import numpy as np
    a = np.frombuffer(raw, np.uint8).reshape(-1, 2)
    out = np.zeros((a.shape[0], nbytes), np.uint8)
    out[:, :2] = a
    return out.tobytes()


def md5_file(path):
    h = hashlib.md5()
    n = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 22)
            if not b:
                break
            h.update(b)
            n += len(b)
    return {"md5": h.hexdigest(), "bytes": n}


def enc_args(c):
    a = ["-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", FMT[c["fmt"]], "-z", str(c["bits"]),
         "-k", c["kernel"], "-d", str(c["wdepth"]), "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"])]
    if c["mode"] == "HQ_ConstQ":
        a += ["-q", str(c["q"])]
    else:
        a += ["-s", str(c["s"])]
    if c["mode"] != "LD":
        a += ["-S", str(c["S"]), "-P", str(c["P"])]
    if c.get("nbytes"):
        a += ["-n", str(c["nbytes"])]
    return a + list(c.get("extra", []))


def run_case(c):
    out = {"params": c, "taps": {}}
    with tempfile.TemporaryDirectory(prefix="vc2gold_") as td:
        src = os.path.join(td, "in.yuv")
        with open(src, "wb") as fo:
            for f in range(c["frames"]):
                raw = gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], c["smooth"])
                fo.write(widen(raw, c["nbytes"]) if c.get("nbytes") else raw)
        out["taps"]["input"] = md5_file(src)
        taps = ["Transform", "Quantised", "Packaged", "Stream"]
        if c["mode"] != "HQ_ConstQ":
            taps.insert(1, "Indices")
        for tap in taps:
            dst = os.path.join(td, "enc_" + tap)
            r = subprocess.run([os.path.join(REF, "EncodeStream")] + enc_args(c) + ["-o", tap, src, dst],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if r.returncode != 0:
                out["taps"]["enc_" + tap] = {"error": r.stdout.decode(errors="replace").strip()[-300:]}
            else:
                out["taps"]["enc_" + tap] = md5_file(dst)
        stream = os.path.join(td, "enc_Stream")
        if "error" not in out["taps"]["enc_Stream"]:
            for tap in ["Quantised", "Indices", "Transform", "Decoded"]:
                dst = os.path.join(td, "dec_" + tap)
                r = subprocess.run([os.path.join(REF, "DecodeStream"), "-o", tap, stream, dst],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                if r.returncode != 0:
                    out["taps"]["dec_" + tap] = {"error": r.stdout.decode(errors="replace").strip()[-300:]}
                else:
                    out["taps"]["dec_" + tap] = md5_file(dst)
    print("done", c["name"], flush=True)
    return c["name"], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--jobs", type=int, default=6)
    a = ap.parse_args()
    path = os.path.join(HERE, "md5.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    cases = [c for c in CASES if not a.only or c["name"] in a.only]
    with ThreadPoolExecutor(a.jobs) as ex:
        for name, res in ex.map(run_case, cases):
            gold[name] = res
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

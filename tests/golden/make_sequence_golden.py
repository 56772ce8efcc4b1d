#!/usr/bin/env python3
"""Golden digests for BASELINE.json config 3 as a whole job: the 240-frame C3 clip (3840x2160 4:2:2 10 bit, DD137 depth 4,
HQ_ConstQ q16, -u1 -a2 -S4) encoded and decoded by the UNMODIFIED reference (oracle/_ref), one process, about half an hour.
Build container only.  Writes tests/golden/sequence.json; tests/test_gpu_sequence.py regenerates the clip on the GPU box
(oracle/gen.py is deterministic) and compares the digests of the drop-in command lines' output.

usage: make_sequence_golden.py [--frames 240] [--dir /tmp/vc2seq]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen  # noqa: E402

SEQ = dict(name="C3x240", w=3840, h=2160, fmt="422", bits=10, seed=4242, mode="HQ_ConstQ", kernel="DD137", wdepth=4, u=1, a=2, q=16, S=4, P=0, r=6)


def enc_args(c):
    return ["-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", "4:2:2", "-z", str(c["bits"]), "-k", c["kernel"], "-d", str(c["wdepth"]),
            "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"]), "-q", str(c["q"]), "-S", str(c["S"]), "-P", str(c["P"])]


def _frame(f):
    c = SEQ
    return gen.frame_bytes(c["seed"], f, c["w"], c["h"], c["fmt"], c["bits"], False)


def write_clip(path, frames, jobs):
    """the clip, frame f from oracle/gen.py with seed SEQ['seed']; returns its md5"""
    h = hashlib.md5()
    with open(path, "wb") as out, ProcessPoolExecutor(jobs) as ex:
        for b in ex.map(_frame, range(frames), chunksize=2):
            out.write(b)
            h.update(b)
    return h.hexdigest()


def md5_file(path):
    h = hashlib.md5()
    n = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
            n += len(b)
    return {"md5": h.hexdigest(), "bytes": n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--dir", default="/tmp/vc2seq")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    src, stream, dec = (os.path.join(a.dir, n) for n in ("in.yuv", "ref.vc2", "ref.dec"))
    res = {"params": dict(SEQ, frames=a.frames)}
    t = time.time()
    res["input_md5"] = write_clip(src, a.frames, os.cpu_count() or 4)
    print("clip written", round(time.time() - t), "s", flush=True)
    t = time.time()
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "EncodeStream")] + enc_args(SEQ) + [src, stream])
    res["reference_encode_s"] = round(time.time() - t, 1)
    res["stream"] = md5_file(stream)
    print("reference encode", res["reference_encode_s"], "s", res["stream"], flush=True)
    t = time.time()
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "DecodeStream"), stream, dec], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res["reference_decode_s"] = round(time.time() - t, 1)
    res["decoded"] = md5_file(dec)
    print("reference decode", res["reference_decode_s"], "s", res["decoded"], flush=True)
    json.dump(res, open(os.path.join(HERE, "sequence.json"), "w"), indent=1, sort_keys=True)
    for p in (src, stream, dec):
        os.remove(p)


if __name__ == "__main__":
    main()

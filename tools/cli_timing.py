#!/usr/bin/env python3
"""Where the wall-clock time of the drop-in command lines goes: N C3 frames (oracle/gen.py), EncodeStream and DecodeStream with
VC2_CLI_TIMING=1 for a few batch sizes.  usage: cli_timing.py [frames] [gpus]"""
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen  # noqa: E402

BIN = os.path.join(ROOT, "vc2_reference_b200", "bin")
ARGS = ["-m", "HQ_ConstQ", "-x", "3840", "-y", "2160", "-f", "4:2:2", "-z", "10", "-k", "DD137", "-d", "4", "-u", "1", "-a", "2", "-r", "6", "-q", "16", "-S", "4", "-P", "0"]


def frame(f):
    return gen.frame_bytes(4242, f, 3840, 2160, "422", 10, False)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    gpus = sys.argv[2] if len(sys.argv) > 2 else "1"
    td = tempfile.mkdtemp(prefix="vc2cli_")
    src, stream, dec = (os.path.join(td, x) for x in ("in.yuv", "s.vc2", "d.yuv"))
    with open(src, "wb") as out, ProcessPoolExecutor(min(32, os.cpu_count() or 4)) as ex:
        for b in ex.map(frame, range(n), chunksize=2):
            out.write(b)
    env = dict(os.environ, VC2_CLI_TIMING="1")
    for batch in ("8", "4", "16"):
        t = time.time()
        r = subprocess.run([os.path.join(BIN, "EncodeStream")] + ARGS + ["-G", gpus, "-B", batch, src, stream], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        print("encode -B %s -G %s: %.2f s wall, %.1f frames/s | %s" % (batch, gpus, time.time() - t, n / (time.time() - t), r.stderr.decode().strip()[-300:]), flush=True)
        t = time.time()
        r = subprocess.run([os.path.join(BIN, "DecodeStream"), "-G", gpus, "-B", batch, stream, dec], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        print("decode -B %s -G %s: %.2f s wall, %.1f frames/s | %s" % (batch, gpus, time.time() - t, n / (time.time() - t), r.stderr.decode().strip()[-300:]), flush=True)
    for p in (src, stream, dec):
        os.remove(p)


if __name__ == "__main__":
    main()

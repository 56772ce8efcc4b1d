#!/usr/bin/env python3
"""Generates the stream-framing fixtures from the compiled, unmodified reference (oracle/_ref/EncodeStream):

  seqhdr.json        sequence header data units for a sweep of picture formats.  The reference writes the
                     sequence header before it reads the first frame (EncodeStream.cpp:437-447), so an empty
                     input file yields exactly that data unit.
  framing_*.npz      for a few small cases: the reference's -o Stream bytes and -o Packaged bytes of the same
                     input, so the host writer / reader can be checked without a GPU.
Run in the build container (needs /root/reference compiled by oracle/build_ref.sh)."""
import itertools
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen  # noqa: E402

ENC = os.path.join(ROOT, "oracle", "_ref", "EncodeStream")


def seq_header(w, h, fmt, bits, rate, depth=1, bff=False):
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "empty"), os.path.join(d, "out")
        open(src, "wb").close()
        cmd = [ENC, "-m", "HQ_ConstQ", "-q", "10", "-x", str(w), "-y", str(h), "-f", fmt, "-l", str(bits), "-n", "1" if bits == 8 else "2",
               "-k", "LeGall", "-d", str(depth), "-u", "1", "-a", "1", "-r", str(rate), src, dst]
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return open(dst, "rb").read().hex() if os.path.exists(dst) else None


def main():
    if "--framing-only" not in sys.argv:
        sequence_headers()
    framing_cases()


def sequence_headers():
    cases = []
    sizes = [(176, 120), (176, 144), (352, 240), (352, 288), (704, 480), (704, 576), (720, 480), (720, 576), (720, 486), (1280, 720),
             (1920, 1080), (2048, 1080), (4096, 2160), (3840, 2160), (7680, 4320), (640, 480), (64, 32), (1000, 600)]
    for (w, h), fmt, bits, rate in itertools.product(sizes, ["4:2:0", "4:2:2", "4:4:4"], [8, 10, 12, 16], [1, 2, 3, 4, 6, 7, 9, 10, 11, 12, 16]):
        hx = seq_header(w, h, fmt, bits, rate)
        if hx:
            cases.append({"w": w, "h": h, "fmt": fmt, "bits": bits, "rate": rate, "hex": hx})
    json.dump(cases, open(os.path.join(HERE, "seqhdr.json"), "w"), indent=0)
    print(len(cases), "sequence headers")


def framing_cases():
    gold = json.load(open(os.path.join(HERE, "md5.json")))
    for name in ["S07_LeGall_d1_444", "S01_LeGall_d3_422", "B00_DD97_d2_420", "S00_DD97_d2_420"]:
        c = gold[name]["params"]
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "in.yuv")
            with open(src, "wb") as f:
                for i in range(c["frames"]):
                    f.write(gen.frame_bytes(c["seed"], i, c["w"], c["h"], c["fmt"], c["bits"]))
            fmt = {"444": "4:4:4", "422": "4:2:2", "420": "4:2:0"}[c["fmt"]]
            base = [ENC, "-m", c["mode"], "-x", str(c["w"]), "-y", str(c["h"]), "-f", fmt, "-z", str(c["bits"]),
                    "-k", c["kernel"], "-d", str(c["wdepth"]), "-u", str(c["u"]), "-a", str(c["a"]), "-r", str(c["r"]), "-S", str(c["S"]),
                    "-P", str(c["P"])]
            base += ["-q", str(c["q"])] if c["mode"] == "HQ_ConstQ" else ["-s", str(c["s"])]
            out = {}
            for tap in ("Stream", "Packaged"):
                dst = os.path.join(d, tap)
                subprocess.run(base + ["-o", tap, src, dst], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                out[tap] = np.frombuffer(open(dst, "rb").read(), np.uint8)
            np.savez_compressed(os.path.join(HERE, "framing_%s.npz" % name), stream=out["Stream"], packaged=out["Packaged"])
            print(name, len(out["Stream"]), len(out["Packaged"]))


if __name__ == "__main__":
    main()

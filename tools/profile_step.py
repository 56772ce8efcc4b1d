#!/usr/bin/env python3
"""Minimal driver for ncu: a few encode+decode steps of a BASELINE config, nothing else.
usage: profile_step.py [C1|C2|C3] [steps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import gen  # noqa: E402
import vc2_reference_b200 as vc2  # noqa: E402

CFG = {
    "C1": dict(w=1920, h=1080, fmt="422", bits=10, kernel="LeGall", depth=3, u=1, a=2, mode="HQ_ConstQ", q=12, s=0, S=1),
    "C2": dict(w=1920, h=1080, fmt="422", bits=10, kernel="DD97", depth=3, u=1, a=2, mode="HQ_CBR", q=0, s=2073600, S=1),
    "C3": dict(w=3840, h=2160, fmt="422", bits=10, kernel="DD137", depth=4, u=1, a=2, mode="HQ_ConstQ", q=16, s=0, S=4),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    c = CFG[name]
    ctx = vc2.Context(0)
    g = vc2.make_geom(c["h"], c["w"], c["fmt"], c["kernel"], c["depth"], c["u"], c["a"], 0, c["S"])
    codec = vc2.Codec(ctx, g, c["mode"], qindex=c["q"], picture_bytes=c["s"], luma_depth=c["bits"], max_pictures=B)
    for i in range(B):
        codec.upload_picture(i, gen.frame_bytes(1234, i, c["w"], c["h"], c["fmt"], c["bits"]))
    for _ in range(2):      # warm-up: lazy module loading, first-touch
        codec.encode(B)
        codec.decode(B)
    ctx.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(steps):
        codec.encode(B)
        codec.decode(B)
    prof = ctx.profile_read()
    for k, (ms, n) in prof.items():
        if n:
            print("%-10s %8.3f ms/step (%d launches/step)" % (k, ms / steps, n // steps))
    codec.close()
    ctx.close()


if __name__ == "__main__":
    main()

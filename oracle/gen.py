#!/usr/bin/env python3
"""Deterministic synthetic picture generator shared by the oracle, tests and bench.

Integer-only (murmur3 finaliser hash), reproducible in numpy / C++ / CUDA.  The
file format is the reference EncodeStream's own input format: planar Y, C1, C2
per frame, 16-bit big-endian words, sample value MSB-justified (<< (16-depth)),
offset binary (/root/reference/src/Library/src/Arrays.cpp:333-379).

usage: gen.py OUT WIDTH HEIGHT {444|422|420} DEPTH FRAMES SEED [--smooth]
"""
import sys
import numpy as np


def fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def plane(seed, c, f, H, W, depth, smooth=False):
    """One plane of unsigned samples in [0, 2**depth) as uint16 (H, W)."""
    y = np.arange(H, dtype=np.uint32)[:, None]
    x = np.arange(W, dtype=np.uint32)[None, :]
    with np.errstate(over="ignore"):
        k = (np.uint32(seed) ^ (np.uint32(c) * np.uint32(0x9E3779B1)) ^ (np.uint32(f) * np.uint32(0x85EBCA77))
             ^ (y * np.uint32(0xC2B2AE3D)) ^ (x * np.uint32(0x27D4EB2F)))
        h = fmix32(k)
    if isinstance(smooth, str) and smooth == "noise":   # full-range white noise: incompressible pictures
        return (h >> np.uint32(32 - depth)).astype(np.uint16)
    noise = (h >> np.uint32(26)).astype(np.int32) - 32
    if smooth:
        noise = noise >> 3
    full = 1 << depth
    ramp = (((x >> np.uint32(1)) + (y >> np.uint32(1)) + np.uint32(4 * f + 37 * c)).astype(np.int64) % full).astype(np.int32)
    checker = ((((x >> np.uint32(6)) ^ (y >> np.uint32(6))) & np.uint32(1)).astype(np.int32)) * (full >> 3)
    v = (ramp >> 1) + (full >> 2) + checker + (noise * (full >> 10) if full >= 1024 else noise >> 2)
    return np.clip(v, 0, full - 1).astype(np.uint16)


def chroma_dims(W, H, fmt):
    cw = W if fmt == "444" else W // 2
    ch = H // 2 if fmt == "420" else H
    return ch, cw


def frame_planes(seed, f, W, H, fmt, depth, smooth=False):
    ch, cw = chroma_dims(W, H, fmt)
    return [plane(seed, c, f, h, w, depth, smooth) for c, (h, w) in enumerate(((H, W), (ch, cw), (ch, cw)))]


def frame_bytes(seed, f, W, H, fmt, depth, smooth=False):
    """File bytes of one frame (16-bit BE, MSB justified)."""
    return b"".join((p << np.uint16(16 - depth)).astype(">u2").tobytes() for p in frame_planes(seed, f, W, H, fmt, depth, smooth))


def main(argv):
    smooth = "--smooth" in argv
    argv = [a for a in argv if a != "--smooth"]
    out, W, H, fmt, depth, frames, seed = argv[1], int(argv[2]), int(argv[3]), argv[4], int(argv[5]), int(argv[6]), int(argv[7])
    with open(out, "wb") as fo:
        for f in range(frames):
            fo.write(frame_bytes(seed, f, W, H, fmt, depth, smooth))


if __name__ == "__main__":
    main(sys.argv)

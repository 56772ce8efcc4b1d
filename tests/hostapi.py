"""ctypes binding of libvc2host.so's C-ABI (include/vc2_host.h): host-side stream framing, no GPU needed."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "vc2_reference_b200", "libvc2host.so")
lib = C.CDLL(PATH)
lib.vc2host_sequence_header.restype = C.c_int
lib.vc2host_sequence_header.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_int]
lib.vc2host_wrap_hq_stream.restype = C.c_longlong
lib.vc2host_wrap_hq_stream.argtypes = [C.c_int] * 13 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
lib.vc2host_parse_units.restype = C.c_int
lib.vc2host_parse_units.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
lib.vc2host_read_sequence_header.restype = C.c_int
lib.vc2host_read_sequence_header.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
lib.vc2host_read_picture_header.restype = C.c_int
lib.vc2host_read_picture_header.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p]

SYMBOLS = ["vc2host_sequence_header", "vc2host_wrap_hq_stream", "vc2host_parse_units", "vc2host_read_sequence_header",
           "vc2host_read_picture_header"]
CF = {"4:4:4": 0, "4:2:2": 1, "4:2:0": 2, "444": 0, "422": 1, "420": 2}


def sequence_header(hq, h, w, fmt, interlace, rate, tff, bits):
    out = np.zeros(256, np.uint8)
    n = lib.vc2host_sequence_header(int(hq), h, w, CF[fmt], int(interlace), rate, int(tff), bits, out.ctypes.data, out.size)
    if n < 0:
        raise ValueError("vc2host_sequence_header failed: %d" % n)
    return out[:n].tobytes()


def wrap_hq_stream(h, w, fmt, rate, tff, bits, kernel, depth, sx, sy, prefix, scalar, payloads):
    bufs = [np.frombuffer(p, np.uint8) for p in payloads]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * len(bufs))(*[b.size for b in bufs])
    out = np.zeros(sum(b.size for b in bufs) + 64 * (len(bufs) + 2), np.uint8)
    n = lib.vc2host_wrap_hq_stream(h, w, CF[fmt], rate, int(tff), bits, kernel, depth, sx, sy, prefix, scalar, len(bufs), ptrs, lens,
                                   out.ctypes.data, out.size)
    if n < 0:
        raise ValueError("vc2host_wrap_hq_stream failed: %d" % n)
    return out[:n].tobytes()


def parse_units(stream):
    buf = np.frombuffer(stream, np.uint8)
    units = np.zeros((4096, 4), np.int64)
    n = lib.vc2host_parse_units(buf.ctypes.data, buf.size, units.shape[0], units.ctypes.data)
    if n < 0:
        raise ValueError("vc2host_parse_units failed")
    return units[:n]


def read_sequence_header(stream, offset):
    buf = np.frombuffer(stream, np.uint8)
    f = np.zeros(10, np.int32)
    if lib.vc2host_read_sequence_header(buf.ctypes.data, buf.size, offset, f.ctypes.data) != 0:
        raise ValueError("vc2host_read_sequence_header failed")
    return dict(zip(["major", "profile", "height", "width", "cf", "interlace", "rate", "tff", "bits", "consumed"], map(int, f)))


def read_picture_header(stream, offset, ld, major):
    buf = np.frombuffer(stream, np.uint8)
    f = np.zeros(9, np.int64)
    if lib.vc2host_read_picture_header(buf.ctypes.data, buf.size, offset, int(ld), major, f.ctypes.data) != 0:
        raise ValueError("vc2host_read_picture_header failed")
    return dict(zip(["picnum", "kernel", "depth", "sx", "sy", "p5", "p6", "consumed"], map(int, f[:8])))

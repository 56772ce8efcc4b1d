"""ctypes access to the compiled reference (oracle/_ref/libvc2ref.so).  Test oracle only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libvc2ref.so")
ENCODE = os.path.join(ROOT, "oracle", "_ref", "EncodeStream")
DECODE = os.path.join(ROOT, "oracle", "_ref", "DecodeStream")
_lib = None


class RefError(RuntimeError):
    pass


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        _lib.ref_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _chk(rc):
    if rc != 0:
        raise RefError(lib().ref_last_error().decode())


def padded_size(size, depth):
    return lib().ref_padded_size(size, depth)


def slice_size_is_valid(depth, luma, chroma, n):
    return lib().ref_slice_size_is_valid(depth, luma, chroma, n)


def quant_matrix(kernel, depth):
    out = np.zeros(3 * depth + 1, np.int32)
    _chk(lib().ref_quant_matrix(kernel, depth, _p(out)))
    return out


def quant(v, q):
    o = C.c_int(0)
    _chk(lib().ref_quant(int(v), int(q), C.byref(o)))
    return o.value


def scale(v, q):
    o = C.c_int(0)
    _chk(lib().ref_scale(int(v), int(q), C.byref(o)))
    return o.value


def quant_factor(q):
    return lib().ref_quant_factor(int(q))


def quant_offset(q):
    return lib().ref_quant_offset(int(q))


def signed_vlc(v):
    n, c = C.c_uint(0), C.c_uint(0)
    _chk(lib().ref_signed_vlc(int(v), C.byref(n), C.byref(c)))
    return n.value, c.value


def dwt_forward(src, kernel, depth):
    src = _i32(src)
    h, w = src.shape
    dst = np.empty((padded_size(h, depth), padded_size(w, depth)), np.int32)
    _chk(lib().ref_dwt_forward(_p(src), h, w, kernel, depth, _p(dst)))
    return dst


def dwt_inverse(src, kernel, depth, shape):
    src = _i32(src)
    dst = np.empty(shape, np.int32)
    _chk(lib().ref_dwt_inverse(_p(src), src.shape[0], src.shape[1], kernel, depth, _p(dst), shape[0], shape[1]))
    return dst


def _quant(fn, coef, qidx, qmatrix):
    coef, qidx, qmatrix = _i32(coef), _i32(qidx), _i32(qmatrix)
    out = np.empty_like(coef)
    _chk(fn(_p(coef), coef.shape[0], coef.shape[1], _p(qidx), qidx.shape[0], qidx.shape[1], _p(qmatrix), qmatrix.size, _p(out)))
    return out


def quantise_np(coef, qidx, qmatrix):
    return _quant(lib().ref_quantise_np, coef, qidx, qmatrix)


def dequantise_np(coef, qidx, qmatrix):
    return _quant(lib().ref_dequantise_np, coef, qidx, qmatrix)


def quantise_ld(coef, qidx, qmatrix):
    return _quant(lib().ref_quantise_ld, coef, qidx, qmatrix)


def dequantise_ld(coef, qidx, qmatrix):
    return _quant(lib().ref_dequantise_ld, coef, qidx, qmatrix)


def slice_bytes(ny, nx, total, scalar):
    out = np.zeros((ny, nx), np.int32)
    _chk(lib().ref_slice_bytes(ny, nx, total, scalar, _p(out)))
    return out


def component_slice_bytes(sl, depth, scalar):
    sl = _i32(sl)
    o = C.c_int(0)
    _chk(lib().ref_component_slice_bytes(_p(sl), sl.shape[0], sl.shape[1], depth, scalar, C.byref(o)))
    return o.value


def slice_bits(u, v, depth):
    u = _i32(u)
    v = _i32(v) if v is not None else None
    o = C.c_int(0)
    _chk(lib().ref_slice_bits(_p(u), _p(v), u.shape[0], u.shape[1], depth, C.byref(o)))
    return o.value


def cbr_qindices(y, u, v, qmatrix, sbytes, scalar):
    y, u, v, qmatrix, sbytes = _i32(y), _i32(u), _i32(v), _i32(qmatrix), _i32(sbytes)
    out = np.empty_like(sbytes)
    _chk(lib().ref_cbr_qindices(_p(y), _p(u), _p(v), y.shape[0], y.shape[1], u.shape[0], u.shape[1], _p(qmatrix), qmatrix.size,
                                _p(sbytes), sbytes.shape[0], sbytes.shape[1], scalar, _p(out)))
    return out


def ld_qindices(y, u, v, qmatrix, sbytes):
    y, u, v, qmatrix, sbytes = _i32(y), _i32(u), _i32(v), _i32(qmatrix), _i32(sbytes)
    out = np.empty_like(sbytes)
    _chk(lib().ref_ld_qindices(_p(y), _p(u), _p(v), y.shape[0], y.shape[1], u.shape[0], u.shape[1], _p(qmatrix), qmatrix.size,
                               _p(sbytes), sbytes.shape[0], sbytes.shape[1], _p(out)))
    return out


def pack_slices(y, u, v, depth, qidx, mode, prefix, scalar, sbytes=None):
    y, u, v, qidx = _i32(y), _i32(u), _i32(v), _i32(qidx)
    sb = _i32(sbytes) if sbytes is not None else None
    cap = 5 * (y.size + u.size + v.size) + 64 * qidx.size + 4096
    if sb is not None:
        cap = max(cap, int(sb.sum()) + (prefix + 8) * qidx.size)
    out = np.zeros(cap, np.uint8)
    ln = C.c_long(0)
    _chk(lib().ref_pack_slices(_p(y), _p(u), _p(v), y.shape[0], y.shape[1], u.shape[0], u.shape[1], depth, _p(qidx),
                               qidx.shape[0], qidx.shape[1], mode, prefix, scalar, _p(sb), _p(out), C.c_long(cap), C.byref(ln)))
    return out[:ln.value].tobytes()


def unpack_slices(data, lh, lw, ch, cw, depth, ny, nx, mode, prefix, scalar, sbytes=None):
    buf = np.frombuffer(data, np.uint8)
    sb = _i32(sbytes) if sbytes is not None else None
    y, u, v = np.empty((lh, lw), np.int32), np.empty((ch, cw), np.int32), np.empty((ch, cw), np.int32)
    q = np.empty((ny, nx), np.int32)
    _chk(lib().ref_unpack_slices(_p(buf), C.c_long(buf.size), lh, lw, ch, cw, depth, ny, nx, mode, prefix, scalar, _p(sb),
                                 _p(y), _p(u), _p(v), _p(q)))
    return y, u, v, q

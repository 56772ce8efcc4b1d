"""Thin numpy-facing wrappers over the C-ABI (tests and bench only; no codec arithmetic here).

Function names mirror the reference Library functions they reach through the C-ABI
(/root/reference/src/Library/{WaveletTransform,Quantisation,Slices}.h).
"""
import ctypes as C
import weakref

import numpy as np

from ._cabi import CodecParams, Geom, SampleFormat, Vc2Error, lib

KERNELS = {"DD97": 0, "LeGall": 1, "DD137": 2, "Haar0": 3, "Haar1": 4, "Fidelity": 5, "Daub97": 6}
MODES = {"HQ_ConstQ": 0, "HQ_VBR": 0, "HQ_CBR": 1, "LD": 2}
CHROMA = {"444": 0, "422": 1, "420": 2, "4:4:4": 0, "4:2:2": 1, "4:2:0": 2}


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _check(st, ctx=None):
    if st != 0:
        msg = lib.vc2_last_error(ctx).decode() if ctx else lib.vc2_status_message(st).decode()
        raise Vc2Error(st, msg or lib.vc2_status_message(st).decode())


# ---- host-only helpers ---------------------------------------------------------------------------
def padded_size(size, depth):
    return lib.vc2_padded_size(size, depth)


def quant_matrix(kernel, depth):
    out = np.zeros(3 * depth + 1, np.int32)
    _check(lib.vc2_quant_matrix(KERNELS.get(kernel, kernel), depth, out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def slice_bytes(ny, nx, total, scalar):
    out = np.zeros((ny, nx), np.int32)
    _check(lib.vc2_slice_bytes(ny, nx, total, scalar, out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def make_geom(height, width, chroma, kernel, depth, v_slice, h_slice, prefix=0, scalar=1):
    g = Geom()
    _check(lib.vc2_make_geom(height, width, CHROMA.get(chroma, chroma), KERNELS.get(kernel, kernel), depth, v_slice, h_slice,
                             prefix, scalar, C.byref(g)))
    return g


def hq_index_slices(payload, n_slices, prefix, scalar):
    buf = np.frombuffer(payload, np.uint8)
    out = np.zeros(n_slices + 1, np.uint32)
    _check(lib.vc2_hq_index_slices(_p(buf), buf.size, n_slices, prefix, scalar, out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


def padded_dims(g):
    d = g.depth
    return (padded_size(g.luma_h, d), padded_size(g.luma_w, d)), (padded_size(g.chroma_h, d), padded_size(g.chroma_w, d))


# ---- context: Library-surface operations on host arrays --------------------------------------------
class Context:
    def __init__(self, device=0, stream=None):
        self.h = lib.vc2_create(device)
        if not self.h:
            raise Vc2Error(-2, "vc2_create failed: no usable CUDA device (there is no CPU fallback)")
        self._codecs = weakref.WeakSet()   # live codecs: destroyed before the context they point into
        if stream is not None:
            _check(lib.vc2_set_stream(self.h, C.c_void_p(stream)), self.h)

    def close(self):
        if self.h:
            for k in list(self._codecs):
                k.close()
            lib.vc2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib.vc2_synchronize(self.h), self.h)

    def kernel_launches(self, reset=False):
        return lib.vc2_kernel_launches(self.h, 1 if reset else 0)

    STAGES = ["dwt_l0", "dwt_deep", "pack", "unpack", "idwt_deep", "idwt_l0", "ld_dc", "assemble", "index", "search"]

    def profile_enable(self, on=True):
        _check(lib.vc2_profile_enable(self.h, 1 if on else 0), self.h)

    def profile_read(self):
        n = len(self.STAGES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        _check(lib.vc2_profile_read(self.h, ms, cnt, n), self.h)
        return {s: (ms[i], cnt[i]) for i, s in enumerate(self.STAGES)}

    # waveletTransform / inverseWaveletTransform
    def waveletTransform(self, picture, kernel, depth):
        src = _i32(picture)
        h, w = src.shape
        dst = np.empty((padded_size(h, depth), padded_size(w, depth)), np.int32)
        _check(lib.vc2_dwt_forward(self.h, _p(src), h, w, KERNELS.get(kernel, kernel), depth, _p(dst)), self.h)
        return dst

    def inverseWaveletTransform(self, transform, kernel, depth, shape):
        src = _i32(transform)
        ph, pw = src.shape
        dst = np.empty(shape, np.int32)
        _check(lib.vc2_dwt_inverse(self.h, _p(src), ph, pw, KERNELS.get(kernel, kernel), depth, _p(dst), shape[0], shape[1]), self.h)
        return dst

    def _quant(self, fn, coef, qidx, qmatrix):
        src = _i32(coef)
        q = _i32(qidx)
        m = _i32(qmatrix)
        depth = (m.size - 1) // 3
        out = np.empty_like(src)
        _check(fn(self.h, _p(src), src.shape[0], src.shape[1], depth, _p(m), _p(q), q.shape[0], q.shape[1], _p(out)), self.h)
        return out

    def quantise_transform_np(self, coef, qidx, qmatrix):
        return self._quant(lib.vc2_quantise_np, coef, qidx, qmatrix)

    def inverse_quantise_transform_np(self, coef, qidx, qmatrix):
        return self._quant(lib.vc2_dequantise_np, coef, qidx, qmatrix)

    def inverse_quantise_transform(self, coef, qidx, qmatrix):   # LD, DC predicted
        return self._quant(lib.vc2_dequantise_ld, coef, qidx, qmatrix)

    def quantise_transform(self, coef, qidx, qmatrix):   # LD, DC predicted
        return self._quant(lib.vc2_quantise_ld, coef, qidx, qmatrix)

    def slice_bits(self, q, q2, depth, ny, nx):
        """luma_slice_bits (q2 None) / chroma_slice_bits of every slice of the quantised in-place plane(s)"""
        q = _i32(q)
        q2 = _i32(q2) if q2 is not None else None
        out = np.empty((ny, nx), np.int32)
        _check(lib.vc2_slice_bits(self.h, _p(q), _p(q2), q.shape[0], q.shape[1], depth, ny, nx, _p(out)), self.h)
        return out

    def component_slice_bytes(self, q, depth, ny, nx, scalar):
        q = _i32(q)
        out = np.empty((ny, nx), np.int32)
        _check(lib.vc2_hq_slice_sizes(self.h, _p(q), q.shape[0], q.shape[1], depth, ny, nx, scalar, _p(out)), self.h)
        return out

    def ld_pack(self, y, u, v, g, qidx, slice_bytes_):
        y, u, v, q, sb = _i32(y), _i32(u), _i32(v), _i32(qidx), _i32(slice_bytes_)
        cap = int(sb.sum()) + 64
        out = np.zeros(cap, np.uint8)
        ln = C.c_size_t(0)
        _check(lib.vc2_ld_pack(self.h, _p(y), _p(u), _p(v), C.byref(g), _p(q), _p(sb), _p(out), cap, C.byref(ln)), self.h)
        return out[:ln.value].tobytes()

    def hq_pack(self, y, u, v, g, qidx, mode="HQ_VBR", slice_bytes_=None, cap=None):
        y, u, v, q = _i32(y), _i32(u), _i32(v), _i32(qidx)
        sb = _i32(slice_bytes_) if slice_bytes_ is not None else None
        n = g.slices_x * g.slices_y
        if cap is None:
            cap = 4 * (y.size + u.size + v.size) + (g.prefix + 4) * n + 1024
            if sb is not None:
                cap = max(cap, int(sb.sum()) + g.prefix * n + 1024)
        out = np.zeros(cap, np.uint8)
        ln = C.c_size_t(0)
        off = np.zeros(n + 1, np.uint32)
        _check(lib.vc2_hq_pack(self.h, _p(y), _p(u), _p(v), C.byref(g), _p(q), MODES.get(mode, mode), _p(sb), _p(out), cap,
                               C.byref(ln), _p(off)), self.h)
        return out[:ln.value].tobytes(), off

    def hq_unpack(self, data, g):
        (ph, pw), (ch, cw) = padded_dims(g)
        buf = np.frombuffer(data, np.uint8)
        y, u, v = np.empty((ph, pw), np.int32), np.empty((ch, cw), np.int32), np.empty((ch, cw), np.int32)
        q = np.empty((g.slices_y, g.slices_x), np.int32)
        _check(lib.vc2_hq_unpack(self.h, _p(buf), buf.size, C.byref(g), _p(y), _p(u), _p(v), _p(q)), self.h)
        return y, u, v, q

    def ld_unpack(self, data, g, slice_bytes_):
        (ph, pw), (ch, cw) = padded_dims(g)
        buf = np.frombuffer(data, np.uint8)
        sb = _i32(slice_bytes_)
        y, u, v = np.empty((ph, pw), np.int32), np.empty((ch, cw), np.int32), np.empty((ch, cw), np.int32)
        q = np.empty((g.slices_y, g.slices_x), np.int32)
        _check(lib.vc2_ld_unpack(self.h, _p(buf), buf.size, C.byref(g), _p(sb), _p(y), _p(u), _p(v), _p(q)), self.h)
        return y, u, v, q

    def quantIndicesCBR(self, y, u, v, g, qmatrix, slice_bytes_):
        y, u, v, m, sb = _i32(y), _i32(u), _i32(v), _i32(qmatrix), _i32(slice_bytes_)
        q = np.empty((g.slices_y, g.slices_x), np.int32)
        flags = np.zeros(g.slices_y * g.slices_x, np.uint32)
        _check(lib.vc2_cbr_qindices(self.h, _p(y), _p(u), _p(v), C.byref(g), _p(m), _p(sb), _p(q), _p(flags)), self.h)
        return q


# ---- fused batched codec -------------------------------------------------------------------------
class Codec:
    def __init__(self, ctx, g, mode="HQ_ConstQ", qindex=0, picture_bytes=0, bytes_per_sample=2, luma_depth=10, chroma_depth=None,
                 max_pictures=1):
        self.ctx = ctx
        self.g = g
        p = CodecParams()
        p.geom = g
        p.fmt = SampleFormat(bytes_per_sample, luma_depth, chroma_depth or luma_depth)
        p.mode = MODES.get(mode, mode)
        p.qindex = qindex
        p.picture_bytes = picture_bytes
        p.max_pictures = max_pictures
        self.max_pictures = max_pictures
        self.h = lib.vc2_codec_create(ctx.h, C.byref(p))
        if not self.h:
            raise Vc2Error(-1, lib.vc2_last_error(ctx.h).decode())
        ctx._codecs.add(self)
        self.picture_bytes = lib.vc2_codec_picture_in_bytes(self.h)
        self.payload_capacity = lib.vc2_codec_payload_capacity(self.h)
        self.n_slices = g.slices_x * g.slices_y

    def close(self):
        if self.h:
            if self.ctx.h:   # a closed context has already destroyed its codecs (Context.close)
                lib.vc2_codec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_picture(self, slot, raw):
        buf = np.frombuffer(raw, np.uint8)
        assert buf.size == self.picture_bytes, (buf.size, self.picture_bytes)
        _check(lib.vc2_codec_upload_picture(self.h, slot, _p(buf)), self.ctx.h)
        self.ctx.synchronize()

    def download_picture(self, slot):
        out = np.empty(self.picture_bytes, np.uint8)
        _check(lib.vc2_codec_download_picture(self.h, slot, _p(out)), self.ctx.h)
        return out.tobytes()

    def upload_payload(self, slot, payload):
        buf = np.frombuffer(payload, np.uint8)
        _check(lib.vc2_codec_upload_payload(self.h, slot, _p(buf), buf.size), self.ctx.h)
        self.ctx.synchronize()

    def encode(self, n):
        _check(lib.vc2_codec_encode_dev(self.h, n), self.ctx.h)

    def decode(self, n):
        _check(lib.vc2_codec_decode_dev(self.h, n), self.ctx.h)

    def set_pipelined(self, on):
        _check(lib.vc2_codec_set_pipelined(self.h, 1 if on else 0), self.ctx.h)

    def slot_status(self, slot):
        _check(lib.vc2_codec_slot_status(self.h, slot), self.ctx.h)

    def download_payload(self, slot):
        out = np.empty(self.payload_capacity, np.uint8)
        ln = C.c_size_t(0)
        q = np.empty((self.g.slices_y, self.g.slices_x), np.int32)
        off = np.empty(self.n_slices + 1, np.uint32)
        _check(lib.vc2_codec_download_payload(self.h, slot, _p(out), out.size, C.byref(ln), _p(q), _p(off)), self.ctx.h)
        return out[:ln.value].tobytes(), q, off

    def _planes(self, fn, slot):
        (ph, pw), (ch, cw) = padded_dims(self.g)
        y, u, v = np.empty((ph, pw), np.int32), np.empty((ch, cw), np.int32), np.empty((ch, cw), np.int32)
        _check(fn(self.h, slot, _p(y), _p(u), _p(v)), self.ctx.h)
        return y, u, v

    def read_transform(self, slot):
        return self._planes(lib.vc2_codec_read_transform, slot)

    def read_quantised(self, slot):
        return self._planes(lib.vc2_codec_read_quantised, slot)

    def read_indices(self, slot):
        q = np.empty((self.g.slices_y, self.g.slices_x), np.int32)
        _check(lib.vc2_codec_read_indices(self.h, slot, _p(q)), self.ctx.h)
        return q

    # end to end with host buffers (numpy uint8 arrays; pinned memory recommended)
    def encode_host(self, pictures, payload_bufs):
        n = len(pictures)
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in pictures])
        oo = (C.c_void_p * n)(*[p.ctypes.data for p in payload_bufs])
        lens = (C.c_size_t * n)()
        cap = min(p.size for p in payload_bufs)
        _check(lib.vc2_codec_encode_host(self.h, n, pp, oo, cap, lens), self.ctx.h)
        return list(lens)

    def decode_host(self, payload_bufs, lens, pictures):
        n = len(pictures)
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in pictures])
        ii = (C.c_void_p * n)(*[p.ctypes.data for p in payload_bufs])
        ll = (C.c_size_t * n)(*lens)
        _check(lib.vc2_codec_decode_host(self.h, n, ii, ll, pp), self.ctx.h)

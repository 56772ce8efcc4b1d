"""ctypes declarations for include/vc2_cabi.h.  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libvc2b200.so")


class Vc2Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__("vc2 status %d: %s" % (status, message))
        self.status = status
        self.message = message


class Geom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("luma_h", "luma_w", "chroma_h", "chroma_w", "kernel", "depth",
                                         "slices_y", "slices_x", "prefix", "scalar")]


class SampleFormat(C.Structure):
    _fields_ = [("bytes_per_sample", C.c_int32), ("luma_depth", C.c_int32), ("chroma_depth", C.c_int32)]


class CodecParams(C.Structure):
    _fields_ = [("geom", Geom), ("fmt", SampleFormat), ("mode", C.c_int32), ("qindex", C.c_int32),
                ("picture_bytes", C.c_int32), ("max_pictures", C.c_int32)]


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(
            "libvc2b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or `make`. "
            "There is no CPU fallback for the hot path." % lib_path)
    L = C.CDLL(lib_path)
    vp, i32p, u32p, u8p, szp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
    gp = C.POINTER(Geom)
    sig = {
        "vc2_create": (vp, [C.c_int]),
        "vc2_destroy": (None, [vp]),
        "vc2_set_stream": (C.c_int, [vp, vp]),
        "vc2_synchronize": (C.c_int, [vp]),
        "vc2_last_error": (C.c_char_p, [vp]),
        "vc2_status_message": (C.c_char_p, [C.c_int]),
        "vc2_device_count": (C.c_int, []),
        "vc2_kernel_launches": (C.c_int, [vp, C.c_int]),
        "vc2_profile_enable": (C.c_int, [vp, C.c_int]),
        "vc2_profile_read": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
        "vc2_padded_size": (C.c_int, [C.c_int, C.c_int]),
        "vc2_slice_size_is_valid": (C.c_int, [C.c_int] * 4),
        "vc2_quant_matrix": (C.c_int, [C.c_int, C.c_int, i32p]),
        "vc2_slice_bytes": (C.c_int, [C.c_int] * 4 + [i32p]),
        "vc2_quant_factor": (C.c_int, [C.c_int]),
        "vc2_quant_offset": (C.c_int, [C.c_int]),
        "vc2_quant_magic31": (C.c_int, [C.c_int, u32p, u32p]),
        "vc2_make_geom": (C.c_int, [C.c_int] * 9 + [gp]),
        "vc2_hq_index_slices": (C.c_int, [vp, C.c_size_t, C.c_int, C.c_int, C.c_int, u32p]),
        "vc2_dwt_forward": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "vc2_dwt_inverse": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]),
        "vc2_quantise_np": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
        "vc2_dequantise_np": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
        "vc2_dequantise_ld": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
        "vc2_hq_pack": (C.c_int, [vp, vp, vp, vp, gp, vp, C.c_int, vp, vp, C.c_size_t, szp, vp]),
        "vc2_hq_unpack": (C.c_int, [vp, vp, C.c_size_t, gp, vp, vp, vp, vp]),
        "vc2_ld_unpack": (C.c_int, [vp, vp, C.c_size_t, gp, vp, vp, vp, vp, vp]),
        "vc2_cbr_qindices": (C.c_int, [vp, vp, vp, vp, gp, vp, vp, vp, vp]),
        "vc2_ld_pack": (C.c_int, [vp, vp, vp, vp, gp, vp, vp, vp, C.c_size_t, szp]),
        "vc2_quantise_ld": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
        "vc2_slice_bits": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "vc2_hq_slice_sizes": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "vc2_host_alloc": (vp, [C.c_size_t]),
        "vc2_bind_thread_to_device": (C.c_int, [C.c_int]),
        "vc2_host_free": (None, [vp]),
        "vc2_codec_create": (vp, [vp, C.POINTER(CodecParams)]),
        "vc2_codec_destroy": (None, [vp]),
        "vc2_codec_picture_in_bytes": (C.c_size_t, [vp]),
        "vc2_codec_payload_capacity": (C.c_size_t, [vp]),
        "vc2_codec_encode_dev": (C.c_int, [vp, C.c_int]),
        "vc2_codec_decode_dev": (C.c_int, [vp, C.c_int]),
        "vc2_codec_set_pipelined": (C.c_int, [vp, C.c_int]),
        "vc2_codec_samples_dev": (vp, [vp, C.c_int]),
        "vc2_codec_recon_dev": (vp, [vp, C.c_int]),
        "vc2_codec_payload_dev": (vp, [vp, C.c_int]),
        "vc2_codec_coeffs_dev": (vp, [vp, C.c_int]),
        "vc2_codec_slice_offsets_dev": (vp, [vp, C.c_int]),
        "vc2_codec_upload_picture": (C.c_int, [vp, C.c_int, vp]),
        "vc2_codec_download_picture": (C.c_int, [vp, C.c_int, vp]),
        "vc2_codec_upload_payload": (C.c_int, [vp, C.c_int, vp, C.c_size_t]),
        "vc2_codec_download_payload": (C.c_int, [vp, C.c_int, vp, C.c_size_t, szp, vp, vp]),
        "vc2_codec_read_transform": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "vc2_codec_read_quantised": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "vc2_codec_read_indices": (C.c_int, [vp, C.c_int, vp]),
        "vc2_codec_slot_status": (C.c_int, [vp, C.c_int]),
        "vc2_codec_encode_host": (C.c_int, [vp, C.c_int, vp, vp, C.c_size_t, szp]),
        "vc2_codec_decode_host": (C.c_int, [vp, C.c_int, vp, szp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)   # AttributeError here = header/library mismatch
        f.restype = res
        f.argtypes = args
    L._vc2_symbols = sorted(sig)
    return L


lib = _load()

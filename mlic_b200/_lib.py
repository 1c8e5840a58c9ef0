"""ctypes binding of include/mlic_b200.h.  There is no fallback: if the shared library is missing
or a call fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLIC_LIB") or os.path.join(_HERE, "libmlic_b200.so")      # MLIC_LIB: development builds (tools/build_variant.sh)

KIND_BASE, KIND_SD, KIND_VBR = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1
MODE_FORWARD, MODE_COMPRESS, MODE_DECODER, MODE_DECOMPRESS = 0, 1, 2, 3

# every symbol include/mlic_b200.h declares (tests check the library exports exactly these)
EXPORTS = (
    "mlic_engine_create", "mlic_engine_destroy", "mlic_engine_set_param", "mlic_engine_finalize",
    "mlic_engine_set_option", "mlic_engine_set_option_f", "mlic_workspace_bytes", "mlic_run", "mlic_run_host", "mlic_last_launch_count",
    "mlic_profile_read", "mlic_profile_read_top", "mlic_trace_dump", "mlic_conv2d_nhwc", "mlic_dwconv3x3_nhwc", "mlic_dsconv_nhwc", "mlic_ds_gdn_nhwc", "mlic_final_subpel", "mlic_local_attn", "mlic_lin_attn", "mlic_chain3", "mlic_ga_head", "mlic_gaussian_conditional", "mlic_last_error", "mlic_version",
    "mlic_engine_set_cdf", "mlic_decompress", "mlic_pmf_to_quantized_cdf", "mlic_rans_encode_bound", "mlic_rans_encode",
    "mlic_rans_decoder_create", "mlic_rans_decoder_destroy", "mlic_rans_decode_stream",
)


class Buffers(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_hat", C.c_void_p), ("y_likelihoods", C.c_void_p), ("z_likelihoods", C.c_void_p),
        ("symbols", C.c_void_p), ("indexes", C.c_void_p), ("z_symbols", C.c_void_p),
        ("y", C.c_void_p), ("y_hat", C.c_void_p), ("rd_sums", C.c_void_p),
    ]


class MlicError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libmlic_b200.so (built by `python -m mlic_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MlicError(f"{LIB_PATH} not found: build the CUDA extension first (python -m mlic_b200.build); "
                        "mlic_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
    L.mlic_engine_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    L.mlic_engine_destroy.argtypes = [vp]
    L.mlic_engine_destroy.restype = None
    L.mlic_engine_set_param.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
    L.mlic_engine_finalize.argtypes = [vp]
    L.mlic_engine_set_option.argtypes = [vp, C.c_char_p, i32]
    L.mlic_engine_set_option_f.argtypes = [vp, C.c_char_p, f32]
    L.mlic_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(sz)]
    L.mlic_run.argtypes = [vp, i32, i32, i32, i32, i32, f32, C.POINTER(Buffers), vp, sz, vp]
    L.mlic_run_host.argtypes = [vp, i32, i32, i32, i32, i32, f32, C.POINTER(Buffers), i32]
    L.mlic_last_launch_count.argtypes = [vp]
    L.mlic_last_launch_count.restype = i64
    L.mlic_profile_read.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.mlic_profile_read_top.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.mlic_trace_dump.argtypes = [vp, C.c_char_p]
    L.mlic_conv2d_nhwc.argtypes = [i32, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, i32,
                                   C.POINTER(f32), vp]
    L.mlic_dwconv3x3_nhwc.argtypes = [i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, i32, C.POINTER(f32), vp]
    L.mlic_dsconv_nhwc.argtypes = [i32, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_ds_gdn_nhwc.argtypes = [i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_final_subpel.argtypes = [i32, vp, i32, i32, i32, i32, vp, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_local_attn.argtypes = [i32, vp, i32, i32, i32, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_lin_attn.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, C.POINTER(f32), vp]
    L.mlic_chain3.argtypes = [i32, vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_ga_head.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, C.POINTER(f32), vp]
    L.mlic_gaussian_conditional.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, vp, vp]
    L.mlic_engine_set_cdf.argtypes = [vp, vp, i32, vp, vp, i32]
    L.mlic_decompress.argtypes = [vp, i32, i32, i32, i32, f32, vp, sz, vp, vp, vp, vp, sz, vp]
    L.mlic_pmf_to_quantized_cdf.argtypes = [vp, i32, vp]
    L.mlic_rans_encode_bound.argtypes = [sz]
    L.mlic_rans_encode_bound.restype = sz
    L.mlic_rans_encode.argtypes = [vp, vp, sz, vp, i32, vp, vp, i32, vp, sz, C.POINTER(sz)]
    L.mlic_rans_decoder_create.argtypes = [vp, sz]
    L.mlic_rans_decoder_create.restype = vp
    L.mlic_rans_decoder_destroy.argtypes = [vp]
    L.mlic_rans_decoder_destroy.restype = None
    L.mlic_rans_decode_stream.argtypes = [vp, vp, sz, vp, i32, vp, vp, i32, vp]
    L.mlic_last_error.restype = C.c_char_p
    L.mlic_version.restype = C.c_char_p
    _lib = L
    return L


def check(status):
    if status != 0:
        raise MlicError(lib().mlic_last_error().decode("utf-8", "replace"))

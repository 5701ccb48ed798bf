"""ctypes binding of the C-ABI CUDA library (include/remora_b200.h).

There is deliberately no fallback: if librb200.so is missing or a call fails, a RemoraError is
raised.  Build the library with ``python -m remora_b200.build_native`` (or
``__graft_entry__.build()``)."""
import ctypes
import os

from . import RemoraError

MAX_CONVS = 4
ARCH_CONVLSTM_W_REF = 1
ARCH_CONV_W_REF = 2
IMPL_AUTO, IMPL_LAYERS, IMPL_FUSED, IMPL_FUSED_TC, IMPL_TILED, IMPL_FUSED_MEGA, IMPL_FUSED_BF16 = 0, 1, 2, 3, 4, 5, 6

# RB200_LIB: load another build of the same library (kernel experiments, scripts/build_variant.sh)
LIB_PATH = os.environ.get("RB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                        "librb200.so")

# every symbol include/remora_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "rb200_version", "rb200_last_error", "rb200_create", "rb200_destroy", "rb200_set_impl",
    "rb200_last_impl", "rb200_launch_count", "rb200_set_debug", "rb200_debug_tensor",
    "rb200_encode_dense", "rb200_forward_dense", "rb200_forward_compact", "rb200_infer_host",
    "rb200_softmax_ml", "rb200_set_profile", "rb200_get_profile", "rb200_infer_host_async", "rb200_chunk_plan", "rb200_chunk_fill",
    "rb200_refine_normalize", "rb200_refine_scratch_bytes", "rb200_refine_dp", "rb200_svb16_decode",
    "rb200_get_flags", "rb200_forward_compact_gather", "rb200_svb16_scratch_bytes",
    "rb200_forward_compact_ship",
)


class ConvDesc(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("c_out", ctypes.c_int32), ("kw", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64)]


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("arch", ctypes.c_int32), ("size", ctypes.c_int32),
        ("kmer_len", ctypes.c_int32), ("num_out", ctypes.c_int32),
        ("n_sig_conv", ctypes.c_int32), ("n_seq_conv", ctypes.c_int32),
        ("n_merge_conv", ctypes.c_int32),
        ("sig_conv", ConvDesc * MAX_CONVS), ("seq_conv", ConvDesc * MAX_CONVS),
        ("merge_conv", ConvDesc * MAX_CONVS),
        ("n_lstm", ctypes.c_int32),
        ("lstm_w_ih_off", ctypes.c_int64 * 2), ("lstm_w_hh_off", ctypes.c_int64 * 2),
        ("lstm_b_off", ctypes.c_int64 * 2),
        ("fc_in", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("fc_w_off", ctypes.c_int64), ("fc_b_off", ctypes.c_int64),
    ]


_lib = None


def load_library():
    """Returns the loaded CDLL; raises RemoraError when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RemoraError(
            f"remora_b200 native library not found at {LIB_PATH}; run "
            "`python -m remora_b200.build_native` (there is no CPU fallback)")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise RemoraError(f"cannot load {LIB_PATH}: {e}")
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.rb200_version.restype = ctypes.c_int
    lib.rb200_last_error.restype = ctypes.c_char_p
    lib.rb200_create.argtypes = [ctypes.POINTER(ModelDesc), vp, i64, ctypes.c_int,
                                 ctypes.POINTER(vp)]
    lib.rb200_destroy.argtypes = [vp]
    lib.rb200_set_impl.argtypes = [vp, ctypes.c_int]
    lib.rb200_last_impl.argtypes = [vp]
    lib.rb200_launch_count.argtypes = [vp]
    lib.rb200_launch_count.restype = ctypes.c_uint64
    lib.rb200_set_debug.argtypes = [vp, ctypes.c_int]
    lib.rb200_debug_tensor.argtypes = [vp, ctypes.c_char_p, vp, i64, ctypes.POINTER(i64),
                                       ctypes.POINTER(i32), ctypes.POINTER(i32), vp]
    lib.rb200_encode_dense.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, vp]
    lib.rb200_forward_dense.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.rb200_forward_compact.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, vp]
    lib.rb200_forward_compact_gather.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, vp, i32, i64, vp, i64, vp]
    lib.rb200_forward_compact_ship.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, vp, i32, i32, vp, i64,
                                               i64, vp, i64, vp]
    lib.rb200_infer_host.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, i32, vp]
    lib.rb200_infer_host_async.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, vp]
    lib.rb200_softmax_ml.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.rb200_chunk_plan.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.rb200_chunk_fill.argtypes = [vp, i32, i32, ctypes.c_double, ctypes.c_double, vp, i32, vp, i32, vp, vp,
                                     vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.rb200_refine_normalize.argtypes = [vp, i32, vp, vp, vp, i32, i64, vp, vp]
    lib.rb200_refine_scratch_bytes.argtypes = [i32, i32, ctypes.POINTER(i64)]
    lib.rb200_refine_dp.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp,
                                    vp, vp, vp, vp, vp]
    lib.rb200_svb16_decode.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.rb200_svb16_scratch_bytes.argtypes = [i32, i32, ctypes.POINTER(i64)]
    lib.rb200_get_flags.argtypes = [vp, ctypes.POINTER(i32), ctypes.c_int]
    lib.rb200_set_profile.argtypes = [vp, ctypes.c_int]
    lib.rb200_get_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_float * 3),
                                      ctypes.POINTER(i32)]
    if lib.rb200_version() != 1:
        raise RemoraError("librb200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load_library().rb200_last_error().decode(errors="replace")
        raise RemoraError(f"{what} failed (code {rc}): {msg}")

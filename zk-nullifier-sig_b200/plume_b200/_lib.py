"""ctypes binding of include/plume_b200.h.  There is no fallback: if the CUDA library is missing
or no B200 is visible, importing works but every use raises."""
import ctypes
import os

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("PLUME_B200_LIB") or os.path.join(PKG, "libplume_b200.so")   # env: experiment builds

_u8p = ctypes.c_void_p
_lib = None

SYMBOLS = {
    "plume_version": (ctypes.c_int, []),
    "plume_ctx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]),
    "plume_ctx_create_multi": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]),
    "plume_ctx_device_count": (ctypes.c_int, [ctypes.c_void_p]),
    "plume_ctx_sub": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "plume_shard_range": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                         ctypes.POINTER(ctypes.c_size_t)]),
    "plume_ctx_destroy": (None, [ctypes.c_void_p]),
    "plume_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "plume_ctx_chunk_items": (ctypes.c_size_t, [ctypes.c_void_p]),
    "plume_sign_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                        _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_verify_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                          _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_hash_to_curve_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t, _u8p]),
    "plume_hash_to_curve_pk_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t, _u8p, _u8p]),
    "plume_hash_to_curve_witness_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                                         _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_fixed_base_mul_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p]),
    "plume_self_test": (ctypes.c_int, [ctypes.c_void_p]),
    "plume_debug_read_arena": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u8p, _u8p, ctypes.c_size_t,
                                              ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "plume_registers_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p]),
    "plume_ark_sign_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                            _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_ark_verify_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                              _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_ark_sign_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                                   _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_void_p]),
    "plume_ark_verify_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                                     _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_void_p]),
    "plume_points_compress_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p]),
    "plume_points_decompress_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, _u8p]),
    "plume_points_compress_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, ctypes.c_void_p]),
    "plume_points_decompress_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, _u8p, ctypes.c_void_p]),
    "plume_sign_batch_sec1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                             _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_verify_batch_sec1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                               _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p]),
    "plume_sign_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                               _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_void_p]),
    "plume_verify_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                                 _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_void_p]),
    "plume_hash_to_curve_batch_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _u8p, _u8p, ctypes.c_size_t,
                                                        _u8p, ctypes.c_void_p]),
    "plume_ctx_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "plume_ctx_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "plume_ctx_stage_ms": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64)]),
    "plume_debug_fe_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, _u8p, _u8p, _u8p]),
    "plume_measure_imad_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "plume_measure_imad_rates": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
}


def load():
    """Load libplume_b200.so (built by plume_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libplume_b200.so is not built (%s); run `python __graft_entry__.py build`. "
                               "There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

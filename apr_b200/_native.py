"""ctypes binding of libaprb200.so (C-ABI declared in include/aprb200.h).

There is NO CPU fallback: importing this module without the built library, or calling into it without a CUDA
device, raises. torch is used for device memory and streams only (pointers are passed as integers).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaprb200.so")

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/aprb200.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "aprb_version": (_i, []),
    "aprb_last_error": (C.c_char_p, []),
    "aprb_launch_count": (C.c_longlong, []),
    "aprb_prof_enable": (_i, [_i]),
    "aprb_prof_report": (_i, [C.c_char_p, _sz]),
    "aprb_set_option": (_i, [C.c_char_p, _i]),
    "aprb_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "aprb_grid_subsample_ws_bytes": (_sz, [_i, _i, _i]),
    "aprb_grid_subsample_batch": (_i, [_p, _p, _i, _i, _f, _i, _p, _i, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "aprb_grid_subsample_batch_labels": (_i, [_p, _p, _i, _i, _f, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "aprb_voxel_downsample_ws_bytes": (_sz, [_i, _i]),
    "aprb_voxel_downsample_raw": (_i, [_p, _i, _p, _i, _i, C.c_double, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "aprb_radius_neighbors_ws_bytes": (_sz, [_i, _i, _i]),
    "aprb_radius_neighbors_batch": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "aprb_cell_grid_bytes": (_sz, [_i, _i]),
    "aprb_cell_grid_build": (_i, [_p, _p, _i, _i, _f, _p, _sz, _p]),
    "aprb_cell_grid_query": (_i, [_p, _sz, _p, _p, _i, _i, _i, _f, _i, _p, _i, _p, _p, _p]),
    "aprb_kpconv_prepare_weights": (_i, [_p, _i, _i, _i, _p, _p]),
    "aprb_kpconv_prepare_weights_f16": (_i, [_p, _i, _i, _i, _p, _p]),
    "aprb_kpconv_ws_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "aprb_kpconv_forward": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "aprb_kpconv_weighted_ws_bytes": (_sz, [_i]),
    "aprb_kpconv_weighted": (_i, [_p, _p, _p, _i, _i, _p, _p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "aprb_kpconv_weighted_f16": (_i, [_p, _p, _p, _i, _p, _p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "aprb_kpconv_tc_supported": (_i, [_i, _i, _i, _i, _i]),
    "aprb_kpconv_prepare_weights_f16_ck": (_i, [_p, _i, _i, _i, _p, _p]),
    "aprb_kpconv_backward_data": (_i, [_p, _p, _p, _i, _i, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p]),
    "aprb_max_pool_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "aprb_max_pool": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "aprb_closest_pool": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "aprb_instnorm_ws_bytes": (_sz, [_i, _i]),
    "aprb_instnorm_lrelu": (_i, [_p, _i, _i, _f, _f, _p, _i, _i, _p, _p, _sz, _p]),
    "aprb_instnorm_backward_ws_bytes": (_sz, [_i]),
    "aprb_instnorm_lrelu_backward": (_i, [_p, _p, _p, _i, _i, _f, _p, _p, _sz, _p]),
    "aprb_instnorm_seg_ws_bytes": (_sz, [_i, _i, _i]),
    "aprb_instnorm_lrelu_seg": (_i, [_p, _i, _i, _p, _i, _f, _f, _p, _i, _i, _p, _p, _sz, _p]),
    "aprb_segment_offsets": (_i, [_p, _i, _i, _p, _p]),
    "aprb_linear_tf32_ws_bytes": (_sz, [_i, _i, _i]),
    "aprb_linear_tf32": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "aprb_linear_tf32_stats": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "aprb_group_stats_bytes": (_sz, [_i, _i]),
    "aprb_kpconv_forward_stats": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "aprb_instnorm_lrelu_seg_pre": (_i, [_p, _i, _i, _p, _i, _f, _f, _p, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "aprb_instnorm_lrelu_seg_f16": (_i, [_p, _i, _i, _p, _i, _f, _f, _p, _i, _i, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "aprb_max_pool_f16": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "aprb_linear_f16_stats": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "aprb_linear_f16_stats_ragged": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _i, _p]),
    "aprb_instnorm_seg_stats": (_i, [_p, _p, _i, _i, _p, _i, _f, _p, _p, _p, _p]),
    "aprb_linear_f16_norm_apply": (_i, [_p, _p, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _f, _p, _i, _p]),
    "aprb_f32_to_f16": (_i, [_p, _p, _sz, _p]),
    "aprb_round_tf32": (_i, [_p, _p, _sz, _p]),
    "aprb_kfe_create": (_i, [_p, _p, _i, _p]),
    "aprb_kfe_arena_bytes": (_sz, [_p, _i, _i]),
    "aprb_kfe_arena_bytes_est": (_sz, [_p, _i, _i, _f]),
    "aprb_kfe_forward": (_i, [_p, _p, _p, _p, _i, _i, _p, _sz, _p, _p, _p, _p]),
    "aprb_kfe_forward_host": (_i, [_p, _p, _p, _i, _i, _p, _sz, _p, _i, _p, _p, _p]),
    "aprb_kfe_forward_host_async": (_i, [_p, _p, _p, _i, _i, _p, _sz, _p, _i, _p, _p, _p, _p]),
    "aprb_kfe_wait_host": (_i, [_p, _i]),
    "aprb_kfe_set_host_output_f16": (_i, [_p, _i]),
    "aprb_kfe_get": (_i, [_p, _i, _i, _p, _p, _p]),
    "aprb_kfe_get_block_output": (_i, [_p, _i, _p, _p, _p, _p]),
    "aprb_kfe_set_tap": (_i, [_p, _p, _sz]),
    "aprb_kfe_tap_count": (_i, [_p]),
    "aprb_kfe_get_tap": (_i, [_p, _i, _p, _p, _p, _p, _p]),
    "aprb_max_pool_seg": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    "aprb_pool_seg_widths": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "aprb_cell_grid_query_nearest": (_i, [_p, _sz, _p, _p, _i, _i, _i, _f, _p, _i, _p]),
    "aprb_cell_grid_query_seg": (_i, [_p, _sz, _p, _p, _i, _i, _i, _f, _i, _p, _i, _i, _p, _p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libaprb200.so (once). Fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        l.aprb_kfe_destroy.restype = None
        l.aprb_kfe_destroy.argtypes = [_p]
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().aprb_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """Device (or host) address of a torch tensor, or None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("apr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib()


def launch_count():
    return int(lib().aprb_launch_count())


def prof_enable(on=True):
    check(lib().aprb_prof_enable(1 if on else 0), "aprb_prof_enable")


def prof_report():
    """{kernel_name: (launches, total_ms)} since the last report (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().aprb_prof_report(buf, len(buf)), "aprb_prof_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out

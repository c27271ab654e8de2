"""ctypes binding of libanimnerf_b200.so (the C ABI declared in include/animnerf_b200.h).

There is no fallback: if the library is missing or a call returns non-zero the caller gets
an exception.  torch is used only for device memory and the current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AN_LIB_PATH") or os.path.join(_HERE, "libanimnerf_b200.so")   # AN_LIB_PATH: A/B variant builds (dev only)

_c = ctypes
_vp, _i32, _i64, _f32, _u64 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_uint64

# name -> (restype, argtypes); mirrors include/animnerf_b200.h one to one
SIGNATURES = {
    "an_version": (_i32, []),
    "an_error_string": (_c.c_char_p, [_i32]),
    "an_raygen_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp]),
    "an_sample_training_rays_fwd": (_i32, [_vp] * 11 + [_i32] * 5 + [_f32, _f32, _i32, _i32, _vp, _u64] + [_vp] * 5),
    "an_rays_sample_fwd": (_i32, [_vp] * 6 + [_i32] * 5 + [_f32, _f32, _f32, _vp, _u64, _vp, _vp, _vp]),
    "an_rays_sample_bwd": (_i32, [_vp] * 9 + [_i32] * 5 + [_f32, _f32, _vp, _vp]),
    "an_ray_point_grad": (_i32, [_vp] * 6 + [_i64, _i32, _vp, _vp, _vp]),
    "an_sample_coarse_fwd": (_i32, [_vp, _i64, _i32, _f32, _vp, _u64, _vp, _vp]),
    "an_vertex_grid_bytes": (_i64, [_i32, _i32]),
    "an_vertex_grid_build": (_i32, [_vp, _i32, _i32, _f32, _f32, _vp, _vp]),
    "an_knn_query_ws_bytes": (_i64, [_i32, _i64]),
    "an_knn_unpose_fwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _i32,
                                 _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_knn_unpose_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_mlp_packed_bytes": (_i64, []),
    "an_mlp_pack": (_i32, [_vp, _vp, _vp, _vp]),
    "an_mlp_stash_bytes": (_i64, [_i64]),
    "an_mlp_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "an_mlp_grad_floats": (_i64, []),
    "an_mlp_bwd_scratch_bytes": (_i64, [_i64]),
    "an_mlp_wgrad_ws_bytes": (_i64, []),
    "an_mlp_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_mlp_bwd_dgrad": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "an_mlp_bwd_wgrad": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "an_mlp_fwd_tangent": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "an_mlp_bwd_wgrad_scaled": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "an_adam_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "an_composite_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "an_composite_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_searchsorted_right": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "an_body_tables_ws_bytes": (_i64, [_i32]),
    "an_body_tables_bwd_ws_bytes": (_i64, [_i32]),
    "an_body_tables_bwd": (_i32, [_vp] * 5 + [_i32] + [_vp] * 6 + [_i32, _i32, _i32] + [_vp] * 7),
    "an_body_tables_fwd": (_i32, [_vp] * 6 + [_i32, _i32] + [_vp] * 7 + [_i32, _i32, _i32] + [_vp] * 6),
    "an_sample_fine_merge_fwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "an_knn_unpose_lattice_fwd": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_compact_ws_bytes": (_i64, []),
    "an_compact_valid": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "an_render_loss_ws_bytes": (_i64, []),
    "an_render_loss": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "an_mc_count": (_i32, [_vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "an_mc_scan": (_i32, [_vp, _i64, _vp, _vp]),
    "an_mc_emit": (_i32, [_vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class AnimNerfB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AnimNerfB200Error(
                "libanimnerf_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU / PyTorch fallback for the rendering path" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(t):
    """device pointer of a tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "non-contiguous tensor passed across the C ABI"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(code, what):
    if code != 0:
        msg = load().an_error_string(code)
        raise AnimNerfB200Error("%s failed: %d (%s)" % (what, code, msg.decode() if msg else "?"))


# kernels launched per entry point (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {"an_mlp_bwd": 4, "an_mlp_bwd_wgrad": 3, "an_mlp_bwd_wgrad_scaled": 3, "an_mlp_pack": 2, "an_knn_unpose_fwd": 2, "an_knn_unpose_lattice_fwd": 2, "an_compact_valid": 2, "an_vertex_grid_build": 2,
                    "an_body_tables_fwd": 2, "an_body_tables_bwd": 2}      # (memsets are not counted)
launch_count = 0
_timing = None          # bench.py: dict name -> list of (start_event, stop_event) on the launching stream


def enable_timing(on=True):
    """Record a CUDA-event pair around every kernel-launching call (on the current stream)."""
    global _timing
    _timing = {} if on else None
    return _timing


def call(name, *args):
    global launch_count
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if _timing is None:
        check(getattr(load(), name)(*args), name)
        return
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    check(getattr(load(), name)(*args), name)
    e1.record(s)
    _timing.setdefault(name, []).append((e0, e1))

"""ctypes binding of libjmodt_b200.so (C ABI declared in include/jmodt_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, an exception is
raised.  PyTorch is used only for device memory and streams; every op below runs a
hand-written sm_100a kernel from jmodt_b200/csrc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libjmodt_b200.so")

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> argtypes (all return int unless listed in _RESTYPES); mirrors include/jmodt_b200.h
SIGNATURES = {
    "jmb_version": [],
    "jmb_last_error": [],
    "jmb_set_device": [_i],
    "jmb_sm_count": [],
    "jmb_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "jmb_ball_query_msg2": [_i, _i, _i, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_ball_query_grid_workspace_bytes": [_i, _i],
    "jmb_ball_query_msg2_grid": [_i, _i, _i, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "jmb_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_furthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_wait_indices": [_vp, _i, _i, _i, _i, _i, _vp],
    "jmb_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_roipool3d": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_roipool3d_canonical": [_i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_roipool3d_canonical_head": [_i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_rcnn_input_fused": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_longlong, _i, _vp, _vp, _i, _vp],
    "jmb_pair_corr": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_pack_point_features": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_boxes_dist": [_i, _vp, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp],
    "jmb_boxes_overlap_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "jmb_boxes_iou_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "jmb_boxes_iou3d": [_i, _vp, _i, _vp, _vp, _vp],
    "jmb_nms_workspace_bytes": [_i],
    "jmb_nms": [_i, _vp, _f, _vp, _vp, _i, _vp, _sz, _vp],
    "jmb_nms_normal": [_i, _vp, _f, _vp, _vp, _i, _vp, _sz, _vp],
    "jmb_pts_in_boxes3d_host": [_i, _i, _vp, _vp, _vp],
    "jmb_roipool3d_host": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_ia_attention": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp],
    "jmb_segmented_scatter_add": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_sa_first_layer": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "jmb_sa_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp],
    "jmb_proposal_workspace_bytes": [_i, _i, _i, _i],
    "jmb_proposal_layer": [_i, _i, _vp, _vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _sz, _vp],
    "jmb_feature_gather": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_tc_mlp_rows": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "jmb_tc_mlp_layer_dot": [_vp, _vp, _i, _i, _i, _i, _vp, C.c_longlong, _i, _i, _vp, _vp, _vp],
    "jmb_tc_dot_finish": [_i, _i, _i, _vp, _f, _i, _vp, _vp],
    "jmb_tc_conv3x3": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "jmb_feature_gather_nhwc": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "jmb_decode_workspace_bytes": [_i, _i],
    "jmb_decode_gather": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "jmb_tc_mlp_layer": [_vp, _vp, _i, _i, _i, _i, _i, _vp, C.c_longlong, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, C.c_longlong, _vp],
}
_RESTYPES = {"jmb_last_error": C.c_char_p, "jmb_nms_workspace_bytes": _sz, "jmb_proposal_workspace_bytes": _sz,
             "jmb_ball_query_grid_workspace_bytes": _sz, "jmb_decode_workspace_bytes": C.c_longlong}


class JmodtB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libjmodt_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise JmodtB200Error("building libjmodt_b200.so failed")
    return SO_PATH


_lib = None
_lock = threading.Lock()
_tls = threading.local()


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise JmodtB200Error(
                        f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C jmodt_b200/csrc`. jmodt_b200 has no CPU fallback.")
                l = C.CDLL(SO_PATH)
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(l, name)  # AttributeError if the library does not export it
                    fn.argtypes = argtypes
                    fn.restype = _RESTYPES.get(name, _i)
                _lib = l
    return _lib


launch_count = 0  # kernels enqueued through the C ABI (bench.py reports it as gpu_launches)


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += {"nms": 2, "proposal_layer": 3, "decode_gather": 5}.get(what, 1)   # nms = mask + sweep kernels, proposal = 3
    if rc != 0:
        msg = lib().jmb_last_error().decode(errors="replace")
        raise JmodtB200Error(f"{what} failed (code {rc}): {msg}")


def stream_and_device(t):
    """Current torch stream handle for tensor t's device; also binds the library's runtime to it."""
    import torch

    if not t.is_cuda:
        raise JmodtB200Error("jmodt_b200 ops need CUDA tensors (there is no CPU path)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if getattr(_tls, "device", None) != dev:
        check(lib().jmb_set_device(dev), "jmb_set_device")
        _tls.device = dev
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()

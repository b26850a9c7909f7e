"""ctypes binding of libg4c.so (include/g4c.h).  There is NO fallback: if the shared library is
missing or a call fails, a RuntimeError is raised — the product path never silently runs on
PyTorch eager or on the CPU oracle."""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G4C_LIB", os.path.join(_HERE, "libg4c.so"))      # G4C_LIB: alternative build (e.g. the -DG4C_PROFILE one)

ACT_NONE, ACT_SELU, ACT_TANH = 0, 1, 2
AGGR_MEAN, AGGR_SUM = 0, 1
PREC_FP32, PREC_FP16X3 = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "fp16x3": PREC_FP16X3}
EDGE_AUTO, EDGE_V3, EDGE_V5 = 0, 1, 2
ACTS = {None: ACT_NONE, "none": ACT_NONE, "selu": ACT_SELU, "tanh": ACT_TANH}
MAX_LAYERS, MAX_SEGS = 3, 3

_f32p = C.c_void_p
_i32p = C.c_void_p


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("in_width", C.c_int32), ("hidden", C.c_int32), ("out_width", C.c_int32),
                ("W_t", _f32p * MAX_LAYERS), ("b", _f32p * MAX_LAYERS), ("ln_gamma", _f32p), ("ln_beta", _f32p)]


class Seg(C.Structure):
    _fields_ = [("ptr", _f32p), ("gather", _i32p), ("width", C.c_int32), ("stride", C.c_int32),
                ("scale", C.c_float), ("_pad", C.c_int32)]


class RowMlpDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("n_segs", C.c_int32), ("act_out", C.c_int32), ("seg", Seg * MAX_SEGS),
                ("mlp", Mlp), ("out", _f32p), ("out_stride", C.c_int32), ("res_stride", C.c_int32),
                ("residual", _f32p)]


class MpDesc(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("aggr", C.c_int32), ("fixed_k", C.c_int32), ("act_e_out", C.c_int32),
                ("act_t_out", C.c_int32), ("precision", C.c_int32), ("n_targets", C.c_int64), ("n_edges", C.c_int64),
                ("rowptr", _i32p), ("src", _i32p), ("edge_perm", _i32p), ("tgt_perm", _i32p),
                ("e_in", _f32p), ("src_feat", _f32p), ("tgt_feat", _f32p), ("e_out", _f32p), ("t_out", _f32p),
                ("edge_mlp", Mlp), ("node_mlp", Mlp)]


class EdgeDesc(C.Structure):
    _fields_ = [("n_targets", C.c_int64), ("n_edges", C.c_int64), ("fixed_k", C.c_int32), ("n_layers", C.c_int32),
                ("act_e_out", C.c_int32), ("aggr", C.c_int32), ("rowptr", _i32p), ("src", _i32p), ("edge_perm", _i32p),
                ("tgt_perm", _i32p), ("e_in", _f32p), ("P_r", _f32p), ("P_c", _f32p), ("e_out", _f32p),
                ("agg_out", _f32p), ("W", C.c_void_p * 3), ("inv_scale", C.c_float * 3), ("p_scale", C.c_float),
                ("bias", _f32p * 3), ("gamma", _f32p), ("beta", _f32p), ("variant", C.c_int32), ("_pad", C.c_int32)]


class RowTcDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("n_segs", C.c_int32), ("n_layers", C.c_int32), ("act_out", C.c_int32),
                ("out_width", C.c_int32), ("out_stride", C.c_int32), ("res_stride", C.c_int32), ("seg", Seg * MAX_SEGS),
                ("W", C.c_void_p * 3), ("inv_scale", C.c_float * 3), ("_pad", C.c_int32), ("bias", _f32p * 3),
                ("gamma", _f32p), ("beta", _f32p), ("out", _f32p), ("residual", _f32p), ("out2", _f32p),
                ("dual", C.c_int32), ("_pad2", C.c_int32)]


class SegReduceDesc(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("width", C.c_int32), ("aggr", C.c_int32), ("act_out", C.c_int32),
                ("_pad", C.c_int32), ("ptr", _i32p), ("idx", _i32p), ("x", _f32p), ("out", _f32p)]


class ProjectDesc(C.Structure):
    _fields_ = [("n_edges", C.c_int64), ("n_feat", C.c_int32), ("n_extra", C.c_int32), ("col", _i32p),
                ("V", _f32p), ("U", _f32p), ("extra", _f32p * 2), ("out", _f32p)]


class EdgeToNodeDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("k", C.c_int32), ("n_feat", C.c_int32), ("Uinv", _f32p), ("e", _f32p),
                ("V", _f32p), ("out_stride", C.c_int32), ("res_stride", C.c_int32), ("residual", _f32p)]


class InterpDesc(C.Structure):
    _fields_ = [("n_out", C.c_int64), ("k", C.c_int32), ("width", C.c_int32), ("x_idx", _i32p), ("w", _f32p),
                ("y_row", _i32p), ("x", _f32p), ("y", _f32p)]


class StepUpdateDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("nf", C.c_int32), ("field_width", C.c_int32), ("in_stride", C.c_int32),
                ("out_stride", C.c_int32), ("t", C.c_int32), ("_pad", C.c_int32), ("pred", _f32p),
                ("node_in", _f32p), ("outputs", _f32p)]


MAX_PEERS = 8


class HaloPutDesc(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("width", C.c_int32), ("n_peers", C.c_int32), ("src", _f32p), ("send_idx", _i32p),
                ("seg_start", C.c_int32 * (MAX_PEERS + 1)), ("_pad", C.c_int32), ("dst", C.c_void_p * MAX_PEERS),
                ("peer_flag", C.c_void_p * MAX_PEERS), ("my_flag", C.c_void_p * MAX_PEERS), ("state", C.c_void_p),
                ("mail_stride", C.c_int64), ("mail", _f32p), ("ghost", _f32p), ("n_recv", C.c_int64)]


class KnnDesc(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_queries", C.c_int64), ("k", C.c_int32), ("exclude_self", C.c_int32),
                ("pos", _f32p), ("query", _f32p), ("cell_start", _i32p), ("sorted_idx", _i32p),
                ("x0", C.c_float), ("y0", C.c_float), ("cell", C.c_float), ("gx", C.c_int32), ("gy", C.c_int32),
                ("_pad", C.c_int32), ("nbr", _i32p)]


class HaloDesc(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("width", C.c_int32), ("_pad", C.c_int32), ("idx", _i32p),
                ("src", _f32p), ("dst", _f32p)]


EXPORTS = {
    "g4c_version": (C.c_int, []),
    "g4c_last_error": (C.c_char_p, []),
    "g4c_launch_count": (C.c_int64, []),
    "g4c_tc_launch_count": (C.c_int64, []),
    "g4c_rowmlp_fwd": (C.c_int, [C.POINTER(RowMlpDesc), C.c_void_p]),
    "g4c_mp_fwd": (C.c_int, [C.POINTER(MpDesc), C.c_void_p]),
    "g4c_rowmlp_tc_fwd": (C.c_int, [C.POINTER(RowTcDesc), C.c_void_p]),
    "g4c_edge_aggr_fwd": (C.c_int, [C.POINTER(EdgeDesc), C.c_void_p]),
    "g4c_seg_reduce_fwd": (C.c_int, [C.POINTER(SegReduceDesc), C.c_void_p]),
    "g4c_project_fwd": (C.c_int, [C.POINTER(ProjectDesc), C.c_void_p]),
    "g4c_edge_to_node_fwd": (C.c_int, [C.POINTER(EdgeToNodeDesc), C.c_void_p]),
    "g4c_interp_fwd": (C.c_int, [C.POINTER(InterpDesc), C.c_void_p]),
    "g4c_step_update": (C.c_int, [C.POINTER(StepUpdateDesc), C.c_void_p]),
    "g4c_halo_pack": (C.c_int, [C.POINTER(HaloDesc), C.c_void_p]),
    "g4c_halo_unpack": (C.c_int, [C.POINTER(HaloDesc), C.c_void_p]),
    "g4c_halo_put": (C.c_int, [C.POINTER(HaloPutDesc), C.c_void_p]),
    "g4c_plan_knn": (C.c_int, [C.POINTER(KnnDesc), C.c_void_p]),
    "g4c_host_guillard": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "g4c_debug_profile": (C.c_int, [C.c_int32, C.c_void_p]),
    "g4c_debug_tma": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "g4c_debug_tc2": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None


def lib():
    """Load libg4c.so once; raise loudly if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C graphs4cfd_b200/csrc` "
                               "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                               "graphs4cfd_b200 has no CPU or eager fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"libg4c error {rc}: {lib().g4c_last_error().decode()}")


def launch_count() -> int:
    return int(lib().g4c_launch_count())


def tc_launch_count() -> int:
    return int(lib().g4c_tc_launch_count())


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def cuda_device(device) -> torch.device:
    """``device`` as a torch.device with an explicit index ("cuda" -> the current device's ordinal)."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def graph_capture(graph: "torch.cuda.CUDAGraph", device: torch.device):
    """``torch.cuda.graph`` on a capture stream that belongs to ``device``.  (torch.cuda.graph's default capture stream is
    created once per process on whatever device was current at its first use; capturing another device's kernels on it
    records an EMPTY graph — seen when an engine on cuda:1 was built before one on cuda:0.)"""
    return torch.cuda.graph(graph, stream=torch.cuda.Stream(device=device))


def launch(name, desc, *tensors):
    """Call the descriptor entry point ``name`` for tensors that must all live on ONE CUDA device: the kernel goes to that
    device's current stream with that device made current for the call (kernel attributes, tensor maps and the launch
    itself belong to the current device; PyTorch's current device is often a different one than the model's)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"graphs4cfd_b200 kernels need CUDA tensors (got device={t.device}); there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"graphs4cfd_b200: tensors of one call live on different devices ({dev} and {t.device})")
    if dev is None:
        raise RuntimeError("graphs4cfd_b200: launch without a device tensor")
    fn = getattr(lib(), name)
    if dev.index == torch.cuda.current_device():
        rc = fn(C.byref(desc), stream_ptr(dev))
    else:
        with torch.cuda.device(dev):
            rc = fn(C.byref(desc), stream_ptr(dev))
    check(rc)


def require_cuda_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("graphs4cfd_b200 kernels need contiguous fp32 CUDA tensors "
                               f"(got device={t.device}, dtype={t.dtype}, contiguous={t.is_contiguous()}); "
                               "there is no CPU fallback")


def host_guillard(senders: np.ndarray, n: int) -> np.ndarray:
    senders = np.ascontiguousarray(senders, dtype=np.int64)
    mask = np.empty(n, dtype=np.uint8)
    check(lib().g4c_host_guillard(senders.ctypes.data_as(C.c_void_p), n, senders.shape[1],
                                  mask.ctypes.data_as(C.c_void_p)))
    return mask.astype(bool)

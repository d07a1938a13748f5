"""ctypes binding of the C-ABI in ``include/umpr_b200.h`` (``umpr_b200/libumpr_b200.so``).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libumpr_b200.so")

P, I, L, F = C.c_void_p, C.c_int, C.c_long, C.c_float

# name -> argtypes; mirrors include/umpr_b200.h line by line (tests/test_abi.py checks both against the header)
SIGNATURES = {
    "umpr_version": [],
    "umpr_sm_count": [I, P],
    "umpr_gather_pack": [P, P, P, P, I, I, I, I, I, P, P],
    "umpr_gru_inproj": [P, P, I, I, I, P, P],
    "umpr_gru_recurrence_fwd": [P, P, P, I, I, I, I, I, P, P, P, P],
    "umpr_gru_recurrence_bwd": [P, P, P, P, P, P, I, I, I, I, I, P, P],
    "umpr_gru_wgrad": [P, P, P, P, I, I, I, I, I, P, I, P],
    "umpr_gru_wgrad_tc": [P, P, P, P, I, I, I, I, I, P, I, P],
    "umpr_gather_pack_tc": [P, P, P, P, I, I, I, I, P, P],
    "umpr_gru_fwd_tc": [P, I, P, I, P, I, P],
    "umpr_gru_bwd_tc": [P, I, P, P, I, P, P, I, P],
    "umpr_sgemm": [P, L, L, P, L, L, P, L, I, I, I, I, I, P, I, P],
    "umpr_tc_gemm_nt": [P, L, P, L, P, L, I, I, I, I, P, I, I, P],
    "umpr_gru_inproj_tc": [P, P, I, I, I, P, I, P],
    "umpr_tc_gemm_tn": [P, L, P, L, P, L, I, I, L, I, P],
    "umpr_tc_gemm_ws": [P, L, P, L, P, L, I, I, I, I, P, I, I, P, I, I, I, P],
    "umpr_coattn_fwd": [P, P, P, I, I, P, P, P, P, P, P, P, P, P, P, P],
    "umpr_coattn_fwd_tc": [P, P, P, I, I, P, I, I, P, I, I, I, P, P, P, P, P, P, P, P, P, P],
    "umpr_coattn_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, P, I, I, P, I, I, P, P, P, P, P, P],
    "umpr_text_match_fwd": [P, P, P, P, P, P, I, P, P],
    "umpr_text_match_bwd": [P, P, P, I, P, P, P, P, P],
    "umpr_text_match_wgrad": [P, P, P, P, P, I, P, P, P],
    "umpr_bce_head_fwd": [P, P, P, P, P, I, P, P, P],
    "umpr_bce_head_bwd": [P, P, P, P, P, P, P, I, P, P, P, P, P],
    "umpr_workspace_bytes": [C.c_char_p, L, L, P],
    "umpr_comm_unique_id": [P],
    "umpr_comm_init": [I, I, P, P],
    "umpr_allreduce": [P, P, L, P],
    "umpr_comm_destroy": [P],
    "umpr_snet_fwd": [P, P, P, I, I, P, P, P, I, P],
    "umpr_snet_sentiment_fwd": [P, P, I, I, I, P, P, P],
    "umpr_snet_sentiment_bwd": [P, P, P, P, I, I, P, P, P],
    "umpr_snet_bwd": [P, P, P, P, P, P, I, I, P, P, P, I, P],
    "umpr_snet_fwd_tc": [P, P, I, P, P, I, I, P, I, P],
    "umpr_snet_bwd_tc": [P, P, I, P, P, P, I, I, P, P, P, I, P],
    "umpr_cnet_prep": [P, I, I, P, P],
    "umpr_cnet_conv_fwd": [P, P, P, I, I, I, P, P, I, P],
    "umpr_cnet_conv_fwd_tc": [P, P, P, I, I, I, I, P, I, P, I, P, P, I, P],
    "umpr_cnet_head_fwd": [P, P, P, F, I, I, I, I, P, P, P],
    "umpr_cnet_head_bwd": [P, P, P, P, P, P, I, I, I, I, P, P, P, P, P],
    "umpr_cnet_conv_bwd_dx": [P, P, P, I, I, I, P, P, P, I, P],
    "umpr_cnet_conv_bwd_dw": [P, P, P, I, I, I, P, P, I, P],
    "umpr_cnet_conv_bwd_dw_tc": [P, P, P, I, I, I, P, I, P, I, P],
    "umpr_cnet_conv_bwd_dx_tc": [P, P, P, I, I, I, P, I, P, P, I, P],
    "umpr_control_tail_fwd": [P, P, P, P, P, F, I, I, I, P, P, P, P, P],
    "umpr_control_tail_bwd": [P, P, P, P, P, P, P, P, F, I, I, I, P, P, P, P, P, P],
    "umpr_visual_fwd": [P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P, P, P],
    "umpr_visual_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P, P, P, P],
    "umpr_fusion_fwd": [P, P, P, P, P, I, I, P, P],
    "umpr_fusion_bwd": [P, P, P, P, P, P, I, I, P, P, P, P, P, P],
    "umpr_loss_fwd": [P, P, P, P, P, P, I, I, F, P, P],
    "umpr_loss_bwd": [P, P, P, P, P, P, P, I, I, F, P, P, P, P, P, P],
    "umpr_tanh_bwd": [P, P, L, P, P],
    "umpr_adam_step": [P, P, P, P, P, L, F, F, F, F, I, F, P, P],
    "umpr_plan_build": [P, P, L, I, P, L, P],
    "umpr_plan_table": [P, P, L, I, I, P, P],
    "umpr_plan_schedule": [P, P, I, I, P, L, P],
    "umpr_collate_ids": [P, P, L, I, L, P, P],
    "umpr_feature_gather": [P, P, L, L, I, L, P, P],
    "umpr_step_workspace_bytes": [P, P, I, P],
    "umpr_step_comm": [P, P, L, L],
    "umpr_step_profile_begin": [C.c_char_p],
    "umpr_step_profile_end": [I, P, P, P, P],
    "umpr_step_streams": [I],
    "umpr_step": [P, P, P, P, P, I, P, I, P, P, C.c_longlong, P, P, I, P],
    "umpr_ssnet_fwd": [P, P, P, L, P, P],
    "umpr_ssnet_bwd": [P, P, P, P, L, P, P, P, P],
}



class GruSeg(C.Structure):
    """``umpr_gru_seg`` of include/umpr_b200.h: one ImprovedRnn call inside a fused tensor-core GRU launch."""
    _fields_ = [("xq", P), ("plan", P), ("out", P), ("hn", P), ("hq", P), ("n_tiles", C.c_int32), ("n_slabs", C.c_int32),
                ("N", C.c_int32), ("L", C.c_int32)]


class GruBwdSeg(C.Structure):
    """``umpr_gru_bwd_seg`` of include/umpr_b200.h."""
    _fields_ = [("d_out", P), ("d_hn", P), ("xq", P), ("hq", P), ("plan", P), ("n_tiles", C.c_int32),
                ("n_slabs", C.c_int32), ("N", C.c_int32), ("L", C.c_int32)]


class StepSide(C.Structure):
    """``umpr_step_side`` of include/umpr_b200.h."""
    _fields_ = [("ids", P), ("plan", P), ("snet_table", P), ("cnet_table", P), ("snet_tiles", C.c_int32), ("cnet_tiles", C.c_int32),
                ("n_tiles", C.c_int32), ("n_slabs", C.c_int32), ("B", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("pv_max", C.c_int32)]


STEP_PARAMS = ["rnet_gru", "M", "snet_u_Ms", "snet_u_Ws", "snet_i_Ms", "snet_i_Ws", "lin_u", "lin_i", "fus_w", "fus_b",
               "cnet_gru", "conv_w", "conv_b", "clin_w", "clin_b", "csnet_Ms", "csnet_Ws", "ss_w", "ss_b", "pos_e", "neg_e", "vis_w", "vis_b"]


class StepModel(C.Structure):
    """``umpr_step_model`` of include/umpr_b200.h (parameter pointers, then gradient pointers in the same order)."""
    _fields_ = ([("review_net_only", C.c_int32), ("V", C.c_int32), ("Pc", C.c_int32), ("F", C.c_int32), ("KC", C.c_int32), ("ksize", C.c_int32),
                 ("E", C.c_int32), ("reserved", C.c_int32), ("threshold", F), ("loss_v_rate", F), ("eq18_eps", F), ("reserved_f", F), ("table", P)]
                + [(n, P * 8 if n.endswith("_gru") else P) for n in STEP_PARAMS]
                + [("g_" + n, P * 8 if n.endswith("_gru") else P) for n in STEP_PARAMS]
                + [("routing_coattn", P), ("routing_cnet", P * 3)])


_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"umpr_b200: CUDA library not built ({LIB_PATH} missing). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C umpr_b200/csrc`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = I
    lib.umpr_last_error.argtypes = []
    lib.umpr_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def last_error() -> str:
    return load().umpr_last_error().decode("utf-8", "replace")


def ptr(t):
    """Device pointer of a contiguous tensor, or NULL for None."""
    if t is None:
        return None
    assert t.is_contiguous(), "umpr_b200: non-contiguous tensor handed to the C-ABI"
    return t.data_ptr()


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
    return arr


def stream():
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {"umpr_coattn_fwd": 2, "umpr_coattn_fwd_tc": 3, "umpr_cnet_conv_fwd_tc": 3, "umpr_cnet_conv_bwd_dx": 2, "umpr_cnet_conv_bwd_dx_tc": 2, "umpr_visual_fwd": 2, "umpr_visual_bwd": 3}
launch_count = 0          # kernels launched by this process through the C-ABI
_timer = None             # optional {"only": set|None, "events": {name: [(start, end), ...]}}


def start_timing(only=None):
    """Record a CUDA-event pair around every call (or only the named entry points) on the launching stream."""
    global _timer
    _timer = {"only": set(only) if only else None, "events": {}, "work": {}}


def stop_timing():
    """→ {name: dict(calls, ms, flops, bytes)}; synchronises the device."""
    global _timer
    t, _timer = _timer, None
    torch.cuda.synchronize()
    out = {}
    for k, v in (t["events"] if t else {}).items():
        w = t["work"].get(k, [0.0, 0.0])
        out[k] = dict(calls=len(v), ms=sum(a.elapsed_time(b) for a, b in v), flops=w[0], bytes=w[1])
    return out


_FN = {}
_raw_stream = torch._C._cuda_getCurrentRawStream
_current_device = torch.cuda.current_device


def call(name, *args, work=None):
    """Invoke a C-ABI entry point on the current CUDA stream (appended as the last argument).
    ``work`` = (algorithmic FLOPs, algorithmic bytes) of this launch, recorded only while timing (bench.py roofline)."""
    global launch_count
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), name)
    timed = _timer is not None and (_timer["only"] is None or name in _timer["only"])
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*args, _raw_stream(_current_device()))
    if timed:
        e1.record()
        _timer["events"].setdefault(name, []).append((e0, e1))
        if work is not None:
            w = _timer["work"].setdefault(name, [0.0, 0.0])
            w[0] += work[0]
            w[1] += work[1]
    if rc != 0:
        kind = "argument error" if rc < 0 else f"CUDA error {rc}"
        raise RuntimeError(f"{name}: {kind}: {last_error()}")
    launch_count += KERNELS_PER_CALL.get(name, 1)


_SM = {}


def sm_count(device) -> int:
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _SM:
        out = C.c_int(0)
        rc = load().umpr_sm_count(idx, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"umpr_sm_count: {last_error()}")
        _SM[idx] = out.value
    return _SM[idx]

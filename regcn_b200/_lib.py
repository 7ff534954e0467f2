"""ctypes binding of libregcn_b200.so (the C ABI declared in include/regcn_b200.h).

There is deliberately no fallback: if the shared library is missing, or a kernel entry point is
called without a B200-class device, the call raises.  PyTorch is used only to own device memory
and the stream; every pointer handed to the library is `tensor.data_ptr()`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libregcn_b200.so")

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float
_d = ctypes.c_double
_sz = ctypes.c_size_t
_u32 = ctypes.c_uint32

# name -> (restype, argtypes); must list every symbol of include/regcn_b200.h (tests check this)
SIGNATURES = {
    "regcn_version": (_i, []),
    "regcn_last_error_string": (ctypes.c_char_p, []),
    "regcn_device_ok": (_i, []),
    "regcn_csr_build_workspace_bytes": (_sz, [_i, _i, _i]),
    "regcn_csr_build": (_i, [_p, _i, _i, _i] + [_p] * 17 + [_p, _sz, _p]),
    "regcn_csr_build_batch_workspace_bytes": (_sz, [_p, _i, _i, _i]),
    "regcn_csr_build_batch": (_i, [_p, _i, _i, _i, _p, _sz, _p]),
    "regcn_csr_concat": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "regcn_rel_mean_pool": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "regcn_union_aggregate": (_i, [_p] * 9 + [_i, _i, _p, _f, _i, _i, _p, _p, _p]),
    "regcn_block_aggregate": (_i, [_p] * 6 + [_i, _i, _i, _i, _p, _p]),
    "regcn_block_aggregate_radius": (_i, [_p, _p, _p, _f] + [_p] * 4 + [_i, _i, _i, _i, _p, _p]),
    "regcn_lorentz_aggregate": (_i, [_p] * 10 + [_i, _i, _i, _i, _i, _d, _p, _p, _p]),
    "regcn_gemm_f32_workspace_bytes": (_sz, [_i, _i, _i]),
    "regcn_gemm_f32": (_i, [_p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _p, _i, _i, _p, _sz, _p]),
    "regcn_split_tf32": (_i, [_p, _p, _p, _sz, _p]),
    "regcn_to_bf16": (_i, [_p, _p, _sz, _p]),
    "regcn_gemm_tf32_workspace_bytes": (_sz, [_i, _i, _i]),
    "regcn_gemm_tf32": (_i, [_p, _p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _sz, _p]),
    "regcn_regcn_evolve_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "regcn_regcn_evolve": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _p, _sz, _p]),
    "regcn_regcn_evolve_shared_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, ctypes.c_longlong, _i]),
    "regcn_regcn_evolve_shared": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _i, _p, _sz, _p]),
    "regcn_gemm_tf32_tune": (None, [_i, _i]),
    "regcn_gemm_tf32_grid_cap": (None, [_i]),
    "regcn_gemm_tf32_a32": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p, _i, _p, _i, _i, _i, _p, _i, _i, _i, _p, _sz, _p, _i, _p]),
    "regcn_gemm_tf32_layer_a32": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p,
                                       _p, _i, _p, _p, _i, _p]),
    "regcn_gemm_tf32_trace": (None, [_p]),
    "regcn_score_count_poly": (None, [_i]),
    "regcn_gemm_tf32_trace_slots": (_i, []),
    "regcn_gemm_tf32_trace_begin": (None, [_p, _sz]),
    "regcn_gemm_tf32_trace_count": (_i, []),
    "regcn_gemm_tf32_trace_read": (_i, [_i, _p, _p, _p, _p, _p, _p, _p]),
    "regcn_gemm_tf32_layer": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _i, _p, _p, _i, _p]),
    "regcn_pdl_enable": (None, [_i]),
    "regcn_two_stream_enable": (None, [_i]),
    "regcn_evolve_a32_mode": (None, [_i]),
    "regcn_kernel_launches": (ctypes.c_longlong, []),
    "regcn_aggregate_tune": (None, [_i]),
    "regcn_score_count_tf32": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _d, _p, _p, _i, _p]),
    "regcn_pair_scores_tf32": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _d, _p, _p, _p, _i, _p]),
    "regcn_score_lse_num_parts": (_i, [_i]),
    "regcn_score_lse_tf32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _d, _p, _p, _i, _p, _p, _p]),
    "regcn_ce_from_lse": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "regcn_ce_rows": (_i, [_p, _i64, _i, _i, _p, _i, _p, _p, _p]),
    "regcn_gather_rows2": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "regcn_gather_scalars": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "regcn_filter_correct": (_i, [_i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "regcn_prof_enable": (None, [_i]),
    "regcn_prof_read": (_i, [_i, _p, _p, _p]),
    "regcn_hyp_evolve_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "regcn_hyp_evolve": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _i, _p, _sz, _p]),
    "regcn_row_map": (_i, [_p, _p, _i, _i, _i, _d, _p, _p]),
    "regcn_row_map_split": (_i, [_p, _p, _p, _p, _i, _i, _i, _d, _p]),
    "regcn_gru_gate": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "regcn_union_combine": (_i, [_p] * 6 + [_i, _i, _i, _i, _d, _p, _p, _p, _p]),
    "regcn_time_gate": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "regcn_hyp_init": (_i, [_p, _p, _i, _i, _i, _i, _d, _f, _f, _p, _p]),
    "regcn_hyp_tangent": (_i, [_p, _i, _i, _d, _p, _p, _p, _p]),
    "regcn_hyp_time_gate": (_i, [_p] * 6 + [_f, _i, _i, _i, _i, _d, _f, _f, _f, _f, _p, _p]),
    "regcn_convtranse_features": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i] + [_p] * 9 + [_p]),
    "regcn_convtrans_fc_pack_weight": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "regcn_convtrans_fc_workspace_bytes": (_sz, [_i, _i]),
    "regcn_convtrans_fc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _i,
                                _p, _i, _p, _p, _p, _sz, _p]),
    "regcn_affine_relu": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "regcn_gather_log0": (_i, [_p, _p, _i, _i, _i, _i, _d, _p, _p]),
    "regcn_hyp_query": (_i, [_p] * 5 + [_i, _i, _i, _d, _p, _p, _p]),
    "regcn_hyp_score_epilogue": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _d, _p, _p, _p]),
    "regcn_rel_curvature": (_i, [_p, _p, _i, _i, _d, _d, _p, _p]),
    "regcn_gather_target_score": (_i, [_p, _i64, _i, _i, _p, _i, _i, _p, _p]),
    "regcn_rank_count": (_i, [_p, _i64, _i, _i, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p]),
    "regcn_counts_to_ranks": (_i, [_p, _p, _i, _p, _p, _p]),
    "regcn_apply_filter": (_i, [_p, _i64, _i, _i, _p, _i, _p, _p, _i, _p, _p]),
    "regcn_queries_prepare": (_i, [_p, _i, _i, _p, _p, _p, _p, _p]),
    "regcn_queries_prepare_batch": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "regcn_filter_count": (_i, [_p, _i, _i, _p, _p]),
    "regcn_filter_fill": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "regcn_filter_fill2": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    # ---- training (csrc/backward.cu)
    "regcn_csr_gather_sum": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _i, _i, _p, _f, _p, _p]),
    "regcn_group_by_key_workspace_bytes": (_sz, [_i]),
    "regcn_group_by_key": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "regcn_expand_rowptr": (_i, [_p, _i, _i, _p, _p, _p]),
    "regcn_normalize_bwd": (_i, [_p, _p, _p, _i, _i, _p]),
    "regcn_gru_gate_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "regcn_union_combine_bwd": (_i, [_p, _p, _p, _i, _i, _f, _p, _p, _p]),
    "regcn_time_gate_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "regcn_tanh_bwd": (_i, [_p, _p, _p, _sz, _p]),
    "regcn_dropout": (_i, [_p, _sz, _f, _u32, _p]),
    "regcn_col_reduce_workspace_bytes": (_sz, [_i, _i]),
    "regcn_bn_stats": (_i, [_p, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _sz, _p]),
    "regcn_bn_bwd_stats": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _sz, _p]),
    "regcn_col_sum": (_i, [_p, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "regcn_bn_act_drop": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _i, _f, _u32, _p, _p]),
    "regcn_bn_bwd_apply": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _i, _f, _p, _p]),
    "regcn_dec_gather_stack": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "regcn_dec_conv_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _f, _u32, _p, _p, _p, _p, _p]),
    "regcn_dec_conv_bwd_input": (_i, [_p, _p, _i, _i, _i, _i, _p, _f, _p, _p]),
    "regcn_dec_conv_bwd_weight_workspace_bytes": (_sz, [_i, _i]),
    "regcn_dec_conv_bwd_weight": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "regcn_ce_lse_rows": (_i, [_p, _i64, _i, _i, _p, _i, _p, _p, _p, _p]),
    "regcn_softmax_grad_rows": (_i, [_p, _i64, _i, _i, _p, _i, _p, _p, _p]),
    "regcn_transpose_split": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, _p]),
    "regcn_adam_workspace_bytes": (_sz, []),
    "regcn_grad_norm": (_i, [_p, _sz, _p, _p, _sz, _p]),
    "regcn_adam_step": (_i, [_p, _p, _p, _p, _sz, _f, _f, _f, _f, _f, _i, _f, _p, _p]),
    "regcn_block_aggregate_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "regcn_block_aggregate_bwd": (_i, [_p] * 10 + [_i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "regcn_static_angle_fwd": (_i, [_p, _p, _i, _i, _f, _f, _i, _p, _p]),
    "regcn_topk_construct_snap": (_i, [_p, _i64, _i, _i, _i, _p, _i, _i, _p, _p, _p]),
    "regcn_atth_query": (_i, [_p] * 8 + [_i, _i, _i, _d, _p, _p, _p]),
    "regcn_gemm_tf32_mn": (_i, [_p, _p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _p, _sz, _p]),
    "regcn_convtrans_decode_rank_workspace_bytes": (_sz, [_i] * 6),
    "regcn_convtrans_decode_rank": (_i, [_p] * 10 + [_i, _p, _p, _p] + [_i] * 7 + [_p, _p, _sz, _p]),
    "regcn_radial_bwd": (_i, [_p, _p, _p, _i, _i, _i, _d, _p]),
    "regcn_row_radius": (_i, [_p, _i, _i, _p, _p]),
    "regcn_row_radius_bwd": (_i, [_p, _p, _i, _i, _p, _p]),
    "regcn_apply_radius": (_i, [_p, _p, _i, _i, _d, _p, _p]),
    "regcn_apply_radius_bwd": (_i, [_p, _p, _p, _i, _i, _d, _p, _p, _p]),
    "regcn_eltwise_fwd": (_i, [_p, _p, _sz, _i, _f, _p]),
    "regcn_eltwise_bwd": (_i, [_p, _p, _p, _sz, _i, _f, _p]),
    "regcn_radius_combine": (_i, [_p, _p, _p, _i, _f, _f, _d, _f, _f, _p, _p]),
    "regcn_radius_combine_bwd": (_i, [_p, _p, _p, _i, _f, _f, _d, _f, _f, _p, _p, _p, _p]),
    "regcn_row_dot": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "regcn_row_dot_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "regcn_edge_radius_grad": (_i, [_p] * 8 + [_f, _i, _i, _p, _p, _p]),
    "regcn_edge_scalar_gather": (_i, [_p, _p, _p, _i, _p, _i, _p]),
    "regcn_radius_mse": (_i, [_p, _p, _p, _i, _f, _f, _d, _f, _p, _p]),
    "regcn_radius_mse_bwd": (_i, [_p, _p, _p, _i, _f, _f, _d, _f, _p, _p, _p]),
    "regcn_mobius_fwd": (_i, [_p, _p, _i, _i, _d, _p, _p]),
    "regcn_mobius_bwd": (_i, [_p, _p, _p, _i, _i, _d, _p, _p, _p]),
    "regcn_eltwise_mul": (_i, [_p, _p, _p, _sz, _p]),
    "regcn_row_axpy": (_i, [_p, _p, _f, _i, _i, _p, _p]),
    "regcn_hyp_dist_grad": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _d, _p, _p, _p, _p, _p]),
    "regcn_givens_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "regcn_givens_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "regcn_attn_mix_fwd": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _p, _p]),
    "regcn_attn_mix_bwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "regcn_lorentz_aggregate_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "regcn_lorentz_bwd_splits": (_i, []),
    "regcn_lorentz_aggregate_bwd": (_i, [_p] * 11 + [_i, _i, _i, _i, _d, _p, _p, _p, _p, _sz, _p]),
    "regcn_hyp_truedist_grad": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "regcn_rel_curvature_bwd": (_i, [_p, _p, _i, _i, _d, _d, _p, _p, _p, _p]),
    "regcn_static_angle_bwd": (_i, [_p, _p, _i, _i, _f, _f, _i, _p, _p, _i, _p, _p]),
}



class CsrArrays(ctypes.Structure):
    """struct regcn_csr_arrays (include/regcn_b200.h): the device pointers of one snapshot's index."""
    _fields_ = ([("triples", _p), ("T", ctypes.c_int32)] +
                [(n, _p) for n in ("src", "dst", "etype", "indeg", "norm", "rowptr", "src_sorted", "etype_sorted",
                                   "eperm", "vptr", "sptr", "vrow_row", "active_pos", "active_rows", "rel_rowptr",
                                   "rel_ents",
                                   "counts")])


_lib = None
launch_count = 0  # kernels-launching C-ABI calls made by this process (bench.py reports it)


def load():
    """Load the shared library (raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"regcn_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C regcn_b200/csrc`; there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().regcn_last_error_string().decode("utf-8", "replace")


def ptr(t):
    """Device pointer of a tensor (None -> NULL); refuses CPU tensors and non-contiguous layouts."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("regcn_b200: tensor must be contiguous")
    return t.data_ptr()


def stream():
    """Raw cudaStream_t of torch's current stream on the current device (the fast private getter when available:
    torch.cuda.current_stream() costs ~15 us of Python per call, more than a kernel launch)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # pragma: no cover
        return torch.cuda.current_stream().cuda_stream


_fn_cache = {}


def call(name, *args):
    """Invoke a kernel entry point on torch's current stream; raise on a non-zero status."""
    global launch_count
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args, torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}: {last_error()}")


def require_device():
    if not torch.cuda.is_available() or not load().regcn_device_ok():
        raise RuntimeError("regcn_b200: no sm_100 (B200) CUDA device visible -- the product path has no CPU fallback")

"""ctypes binding of the C-ABI CUDA library (include/omni_avsr.h).

There is no CPU fallback: if the library cannot be loaded the import raises, and every wrapper
refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libomni_avsr.so"

OMNI_ERR_UNSUPPORTED = -4
ERR = {-1: "bad argument", -2: "CUDA error", -3: "CUDA driver entry point unavailable", -4: "unsupported",
       -5: "workspace too small"}


class OmniKernelError(RuntimeError):
    pass


def _load() -> C.CDLL:
    # (re)build when the sources changed since the last build (no-op when the stamp matches).  A failed rebuild is fatal
    # even when an older library is lying around: running kernels that do not correspond to the sources would make every
    # measurement and parity claim about this tree meaningless (OMNI_ALLOW_STALE_LIB=1 overrides, for bisecting only).
    import os
    try:
        from .build import build
        build()
    except Exception:
        if not (LIB_PATH.exists() and os.environ.get("OMNI_ALLOW_STALE_LIB") == "1"):
            raise
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m omni_avsr_b200.build` (needs nvcc)")
    return C.CDLL(str(LIB_PATH))


lib = _load()


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("A2", C.c_void_p), ("B2", C.c_void_p),
        ("out", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("tile_group", C.c_void_p), ("b_row_table", C.c_void_p), ("ext_table", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("lda2", C.c_int64), ("ldb2", C.c_int64),
        ("ldo", C.c_int64), ("ldr", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("b_rows", C.c_int32),
        ("a2_cols", C.c_int32), ("b2_rows", C.c_int32), ("b2_cols", C.c_int32),
        ("n_ext", C.c_int32), ("block_n", C.c_int32), ("act", C.c_int32), ("out_fp32", C.c_int32),
        ("alpha", C.c_float), ("pair_aligned", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("slope", C.c_void_p), ("res_bias", C.c_void_p),
        ("ring_h", C.c_int32), ("ring_w", C.c_int32), ("ring_group", C.c_int32), ("ring_c", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("norm_weight", C.c_void_p), ("norm_out", C.c_void_p), ("norm_ld", C.c_int64), ("norm_eps", C.c_float),
    ]


class BeamSelectArgs(C.Structure):
    """omni_beam_select_args (include/omni_avsr.h)."""
    _fields_ = [
        ("cand_score", C.c_void_p), ("cand_tok", C.c_void_p), ("beam_scores", C.c_void_p),
        ("step_idx", C.c_void_p), ("eos", C.c_void_p), ("pad", C.c_void_p),
        ("seqs", C.c_void_p), ("ind", C.c_void_p),
        ("hyp_seq", C.c_void_p), ("hyp_len", C.c_void_p), ("hyp_score", C.c_void_p), ("hyp_order", C.c_void_p),
        ("hyp_count", C.c_void_p), ("hyp_worst", C.c_void_p),
        ("done", C.c_void_p), ("n_done", C.c_void_p), ("status", C.c_void_p),
        ("embed", C.c_void_p), ("x_next", C.c_void_p),
        ("ld_embed", C.c_int64), ("ld_x", C.c_int64),
        ("B", C.c_int32), ("K", C.c_int32), ("n_cand", C.c_int32), ("V", C.c_int32), ("max_new", C.c_int32),
        ("ind_ld", C.c_int32), ("H", C.c_int32),
    ]


class SpliceArgs(C.Structure):
    _fields_ = [
        ("tokens", C.c_void_p), ("labels", C.c_void_p), ("embed", C.c_void_p),
        ("audio_tok", C.c_void_p), ("video_tok", C.c_void_p),
        ("prompt", C.c_void_p * 3), ("out", C.c_void_p * 3), ("out_labels", C.c_void_p * 3),
        ("prompt_len", C.c_int32 * 3),
        ("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("n_a", C.c_int32), ("n_v", C.c_int32),
        ("id_audio_sos", C.c_int32), ("id_audio_eos", C.c_int32),
        ("id_video_sos", C.c_int32), ("id_video_eos", C.c_int32),
        ("has_bos", C.c_int32), ("task_mask", C.c_int32),
        ("vocab", C.c_int64), ("status", C.c_void_p),
    ]


class PpsModality(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_bs", C.c_int64),
        ("n_tok", C.c_int32), ("rate", C.c_int32), ("D", C.c_int32),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("pooled", C.c_void_p), ("hidden", C.c_void_p), ("tok", C.c_void_p),
    ]


class PpsArgs(C.Structure):
    _fields_ = [
        ("audio", PpsModality), ("video", PpsModality), ("splice", SpliceArgs),
        ("I", C.c_int32), ("mode", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


WGRAD_MAX_RANGES = 8


class WgradArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("out", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldo", C.c_int64), ("out_zstride", C.c_int64),
        ("K", C.c_int32), ("a_cols", C.c_int32), ("b_cols", C.c_int32),
        ("Mo", C.c_int32), ("No", C.c_int32), ("a_col0", C.c_int32), ("b_col0", C.c_int32),
        ("n_ranges", C.c_int32),
        ("k0", C.c_int32 * WGRAD_MAX_RANGES), ("k1", C.c_int32 * WGRAD_MAX_RANGES),
        ("out_fp32", C.c_int32), ("accumulate", C.c_int32),
        ("alpha", C.c_float),
    ]


def _sig(name, argtypes, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


_sig("omni_abi_version", [])
_sig("omni_device_cc", [])
_sig("omni_gemm_bf16", [C.POINTER(GemmArgs), C.c_void_p])
_sig("omni_gemm_skinny_bf16", [C.POINTER(GemmArgs), C.c_void_p])
_sig("omni_gemm_skinny_workspace_bytes", [], C.c_int64)
_sig("omni_matryoshka_compress",
     [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p])
_sig("omni_matryoshka_compress_bwd",
     [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p])
_sig("omni_splice_seq_len", [C.POINTER(SpliceArgs), C.c_int32], C.c_int32)
_sig("omni_splice_prompt", [C.POINTER(SpliceArgs), C.c_void_p])
_sig("omni_splice_prompt_bwd", [C.POINTER(SpliceArgs), C.c_void_p * 3, C.c_void_p, C.c_void_p, C.c_void_p])
_sig("omni_pps_workspace_bytes", [C.POINTER(PpsArgs)], C.c_int64)
_sig("omni_pool_project_splice", [C.POINTER(PpsArgs), C.c_void_p])


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_sig("omni_rmsnorm_fwd", [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _F, _P])
_sig("omni_rmsnorm_bwd", [_P, _P, _P, _P, _P, _P, _I64, _I32, _P])
_sig("omni_layernorm_fwd", [_P, _P, _P, _P, _P, _P, _I64, _I32, _I64, _I64, _F, _P])
_sig("omni_layernorm_bwd", [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P])
_sig("omni_rope", [_P, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P])
_sig("omni_swiglu_fwd", [_P, _P, _I64, _I32, _P])
_sig("omni_swiglu_bwd", [_P, _P, _P, _I64, _I32, _P])
_sig("omni_swiglu_bwd_blocked", [_P, _P, _P, _I64, _I32, _I32, _P])
_sig("omni_gelu_fwd", [_P, _P, _I64, _P])
_sig("omni_gelu_bwd", [_P, _P, _P, _I64, _P])
_sig("omni_gather_rows", [_P, _P, _P, _I64, _I32, _I64, _I64, _P, _P])
_sig("omni_scatter_rows", [_P, _P, _P, _I64, _I32, _I64, _P])
_sig("omni_transpose_bf16", [_P, _P, _I32, _I32, _I32, _P])
_sig("omni_ce_fwd", [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _P])
_sig("omni_ce_bwd", [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _P])
_sig("omni_argmax", [_P, _P, _I64, _I32, _I64, _P])
_sig("omni_sumsq", [_P, _I64, _P, _P])
_sig("omni_adamw", [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I32, _F, _F, _P, _P])
_sig("omni_gemm_wgrad_bf16", [C.POINTER(WgradArgs), _P])
_sig("omni_colsum_bf16", [_P, _P, _I64, _I32, _I64, _P])
_sig("omni_prelu_res", [_P, _P, _P, _P, _P, _I64, _I32, _P])
_sig("omni_prelu_maxpool3x3s2", [_P, _P, _P, _I64, _I32, _I32, _I32, _P])
_sig("omni_im2col_front3d", [_P, _P, _I32, _I32, _I32, _I32, _P])
_sig("omni_im2col_front2d", [_P, _P, _I32, _I32, _I32, _I32, _P])
_sig("omni_prelu_maxpool_front", [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P])
_sig("omni_prelu_maxpool_front_ring", [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P])
_sig("omni_prelu_res_ring", [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P])
_sig("omni_gather_s2_ring", [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P])
_sig("omni_avgpool_ring", [_P, _P, _I64, _I32, _I32, _I32, _P])
_sig("omni_avgpool_frames", [_P, _P, _I64, _I32, _I32, _I32, _P])
_sig("omni_attention_fwd", [_P, _I64, _I64, _P, _I64, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _F, _P])
_sig("omni_attention_bwd", [_P, _I64, _I64, _P, _I64, _P, _I64, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32,
                            _I32, _F, _P])
_sig("omni_attention_bwd_scratch_floats", [_I32, _I32, _I32], C.c_int64)
_sig("omni_decode_pick", [_P, _I32, _I64, _P, _P, _P, _P, _P, _I32, _P, _P, _I64, _P, _I64, _I32, _P])
_sig("omni_decode_advance", [_P, _P, _P, _I32, _P])
_sig("omni_decode_attention", [_P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _F, _P])
_sig("omni_decode_attention_rope", [_P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _F, _P, _P, _I32, _P])
_sig("omni_decode_attention_beam", [_P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _F, _P, _P, _I32, _P, _I32,
                                    _P, _I32, _P])
_sig("omni_beam_topk_rows", [_P, _I64, _I32, _I32, _P, _I32, _P, _P, _P])
_sig("omni_beam_select", [C.POINTER(BeamSelectArgs), _P])
_sig("omni_logmel_workspace_bytes", [_I32], C.c_int64)
_sig("omni_logmel", [_P, _I32, _I64, _I32, _I32, _P, _P, _P, _I64, _P])
_sig("omni_video_transform", [_P, _I32, _I32, _I32, _I32, _I32, _I32, C.POINTER(C.c_int32), _I32, _P, _I32, _P])
_sig("omni_audio_transform_workspace_bytes", [], C.c_int64)
_sig("omni_audio_transform", [_P, _P, _I64, _F, C.POINTER(C.c_int32), _I32, _P, _P, _I64, _P])

# every symbol include/omni_avsr.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = [
    "omni_abi_version", "omni_device_cc", "omni_gemm_bf16", "omni_gemm_skinny_bf16", "omni_gemm_skinny_workspace_bytes", "omni_matryoshka_compress",
    "omni_matryoshka_compress_bwd", "omni_splice_seq_len", "omni_splice_prompt", "omni_splice_prompt_bwd",
    "omni_rmsnorm_fwd", "omni_rmsnorm_bwd", "omni_layernorm_fwd", "omni_layernorm_bwd", "omni_rope",
    "omni_swiglu_fwd", "omni_swiglu_bwd", "omni_swiglu_bwd_blocked", "omni_gelu_fwd", "omni_gelu_bwd", "omni_gather_rows", "omni_scatter_rows",
    "omni_ce_fwd", "omni_ce_bwd", "omni_argmax", "omni_decode_pick", "omni_decode_advance", "omni_sumsq", "omni_adamw", "omni_gemm_wgrad_bf16",
    "omni_colsum_bf16", "omni_logmel_workspace_bytes", "omni_logmel", "omni_prelu_res", "omni_prelu_maxpool3x3s2",
    "omni_im2col_front3d", "omni_im2col_front2d", "omni_prelu_maxpool_front", "omni_prelu_maxpool_front_ring", "omni_prelu_res_ring", "omni_gather_s2_ring", "omni_avgpool_ring", "omni_avgpool_frames", "omni_attention_fwd", "omni_attention_bwd", "omni_attention_bwd_scratch_floats", "omni_decode_attention", "omni_decode_attention_rope",
    "omni_decode_attention_beam", "omni_beam_topk_rows", "omni_beam_select",
    "omni_transpose_bf16", "omni_pps_workspace_bytes", "omni_pool_project_splice", "omni_video_transform", "omni_audio_transform_workspace_bytes", "omni_audio_transform",
]


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise OmniKernelError(f"{what} failed: {ERR.get(rc, rc)}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OmniKernelError("omni_avsr_b200 kernels are CUDA-only (sm_100a); got a CPU tensor — "
                                  "there is no CPU fallback (the CPU restatement lives in oracle/ for tests only)")

"""ctypes binding of `include/arxiv_rag_b200.h`.

The CUDA library is the product: there is no CPU fallback. Importing this module never touches
the GPU; `lib()` fails loudly if `lib/libarxiv_rag_b200.so` has not been built
(`python -m arxiv_rag_b200.build`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# ARB_LIB_PATH selects another build of the same library (e.g. the -DARB_HANG_GUARD debug build).
LIB_PATH = Path(os.environ.get("ARB_LIB_PATH") or
                Path(__file__).resolve().parent / "lib" / "libarxiv_rag_b200.so")

ARB_DTYPE_F32 = 0
ARB_DTYPE_BF16 = 1
ARB_DTYPE_F16 = 2
ARB_DTYPE_BF16_PURE = 3  # encoder handles: bf16 without the short-batch fp16 path (A/B baseline)
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2


class ArbError(RuntimeError):
    """A C-ABI call returned a negative code; carries the library's message."""

    def __init__(self, code: int, message: str):
        super().__init__(f"arxiv_rag_b200 error {code}: {message}")
        self.code = code


class MpnetConfig(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32),
        ("max_position_embeddings", C.c_int32),
        ("hidden_size", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_heads", C.c_int32),
        ("intermediate_size", C.c_int32),
        ("relative_attention_num_buckets", C.c_int32),
        ("pad_token_id", C.c_int32),
        ("layer_norm_eps", C.c_float),
        ("compute_dtype", C.c_int32),
        ("position_mode", C.c_int32),
    ]


_FP = C.POINTER(C.c_float)


class MpnetLayerWeights(C.Structure):
    _fields_ = [(n, _FP) for n in (
        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b", "attn_ln_g", "attn_ln_b",
        "ffn_in_w", "ffn_in_b", "ffn_out_w", "ffn_out_b", "out_ln_g", "out_ln_b")]


class MpnetWeights(C.Structure):
    _fields_ = [
        ("word_embeddings", _FP),
        ("position_embeddings", _FP),
        ("emb_ln_g", _FP),
        ("emb_ln_b", _FP),
        ("relative_attention_bias", _FP),
        ("layers", C.POINTER(MpnetLayerWeights)),
    ]


# name -> (restype, argtypes); every symbol include/arxiv_rag_b200.h declares.
_VP, _I32, _I64, _SZ, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
SIGNATURES = {
    "arb_last_error": (C.c_char_p, []),
    "arb_abi_version": (C.c_int, []),
    "arb_mpnet_create": (C.c_int, [C.POINTER(MpnetConfig), C.POINTER(MpnetWeights), _I64, _I32, _I32,
                                   C.POINTER(_VP)]),
    "arb_mpnet_destroy": (C.c_int, [_VP]),
    "arb_mpnet_device_bytes": (_I64, [_VP]),
    "arb_mpnet_encode": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _VP, _VP]),
    "arb_mpnet_launches_per_encode": (C.c_int, [_VP]),
    "arb_mpnet_status": (C.c_int, [_VP]),
    "arb_mpnet_short_seq": (C.c_int, [_VP]),
    "arb_mpnet_relative_bucket": (C.c_int, [_I32, _I32, _I32]),
    "arb_topk_search_workspace_bytes": (_SZ, [_I32, _I64, _I64, _I32, _I32]),
    "arb_topk_search": (C.c_int, [_VP, _VP, _I32, _I64, _I64, _I32, _I32, _VP, _VP, _I64, _VP, _SZ, _VP]),
    "arb_topk_search_f32_workspace_bytes": (_SZ, [_I64, _I64, _I32, _I32, _I32]),
    "arb_topk_search_f32": (C.c_int, [_VP, _VP, _I64, _I64, _I32, _I32, _F, _VP, _VP, _I64, _VP, _I32, _VP, _SZ, _VP]),
    "arb_topk_merge": (C.c_int, [_VP, _VP, _I32, _I64, _I32, _VP, _VP, _VP]),
    "arb_topk_search_launches": (C.c_int, [_I32]),
    "arb_set_gemm_mode": (C.c_int, [_I32]),
    "arb_set_pdl_mode": (C.c_int, [_I32]),
    "arb_set_search_mode": (C.c_int, [_I32]),
    "arb_set_search_pace": (C.c_int, [_I32]),
    "arb_gemm16_lnfold": (C.c_int, [_VP, _I64, _VP, _I64, _VP, _I64, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _I32, _I32, _VP,
                                    C.c_float, _I64, _I32, _I32, _I32, _I32, _VP]),
    "arb_topk_exchange_bytes": (_SZ, [_I32, _SZ]),
    "arb_exchange_alloc": (C.c_int, [_SZ, C.POINTER(C.c_void_p)]),
    "arb_exchange_free": (C.c_int, [_VP]),
    "arb_ipc_export": (C.c_int, [_VP, _VP]),
    "arb_ipc_import": (C.c_int, [_VP, C.POINTER(C.c_void_p)]),
    "arb_ipc_close": (C.c_int, [_VP]),
    "arb_topk_exchange_merge": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _I32, _SZ, _VP, _VP, _VP]),
    "arb_topk_exchange_status": (C.c_int, [_VP]),
    "arb_topk_record_bytes": (_SZ, [_I64, _I32]),
    "arb_topk_record_ids_offset": (_SZ, [_I64, _I32]),
    "arb_topk_merge_records": (C.c_int, [_VP, _I32, _I64, _I32, _VP, _VP, _VP]),
    "arb_adjacent_cosine": (C.c_int, [_VP, _I64, _I32, _VP, _VP]),
    "arb_tokenizer_create": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, C.POINTER(C.c_void_p)]),
    "arb_tokenizer_destroy": (C.c_int, [_VP]),
    "arb_tokenizer_encode": (C.c_int, [_VP, _VP, _VP, _I64, _I32, _I32, _VP, _I64, _VP, _VP]),
    "arb_gemm16": (C.c_int, [_VP, _I64, _VP, _I64, _VP, _I64, _VP, _VP, _I64, _I64, _I32, _I32, _I32, _I32, _VP]),
    "arb_gemm16_f32out": (C.c_int, [_VP, _I64, _VP, _I64, _VP, _I64, _I64, _I32, _I32, _I32, _VP]),
    "arb_embed_layernorm": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _F, _I32, _VP]),
    "arb_layernorm16": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _I32, _F, _I32, _VP]),
    "arb_attention16": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    "arb_pool_normalize": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _VP]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the CUDA library; raises if it was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m arxiv_rag_b200.build` "
                "(nvcc, sm_100a). arxiv_rag_b200 has no CPU fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().arb_last_error()
        raise ArbError(code, msg.decode("utf-8", "replace") if msg else "")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return 0
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream

"""MPNet weights: config, HF-state-dict -> C-ABI struct packing, and a seeded synthetic generator.

There are no all-mpnet-base-v2 weights offline, so "all-mpnet-base-v2" here means
architecture- and shape-identical with seeded synthetic weights; `from_state_dict` accepts a real
`transformers.MPNetModel.state_dict()` unchanged when one is available.
Parameter names follow transformers' MPNetModel (modeling_mpnet.py:57-96, 116-273, 284-291).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class MPNetArch:
    """Published all-mpnet-base-v2 architecture (config.json + sentence_bert_config.json)."""

    vocab_size: int = 30527
    max_position_embeddings: int = 514
    hidden_size: int = 768
    num_layers: int = 12
    num_heads: int = 12
    intermediate_size: int = 3072
    relative_attention_num_buckets: int = 32
    pad_token_id: int = 1
    layer_norm_eps: float = 1e-5
    max_seq_length: int = 384
    kind: str = "mpnet"  # "mpnet" (relative-position bias, pad-aware positions) or "bert"

    def c_struct(self, compute_dtype: int = _lib.ARB_DTYPE_BF16) -> _lib.MpnetConfig:
        bert = self.kind == "bert"
        return _lib.MpnetConfig(self.vocab_size, self.max_position_embeddings, self.hidden_size,
                                self.num_layers, self.num_heads, self.intermediate_size,
                                0 if bert else self.relative_attention_num_buckets, self.pad_token_id,
                                self.layer_norm_eps, compute_dtype, 1 if bert else 0)


ALL_MPNET_BASE_V2 = MPNetArch()
# all-MiniLM-L6-v2 (BERT-style; the reference's semantic-chunking encoder, text_processor.py:853-885,
# and the second `--model` choice of generate_embeddings_parallel.py:473-475): published config.json
# values; sentence_bert_config max_seq_length 256.
ALL_MINILM_L6_V2 = MPNetArch(vocab_size=30522, max_position_embeddings=512, hidden_size=384, num_layers=6,
                             num_heads=12, intermediate_size=1536, relative_attention_num_buckets=0,
                             pad_token_id=0, layer_norm_eps=1e-12, max_seq_length=256, kind="bert")
ARCH_BY_MODEL_NAME = {"all-mpnet-base-v2": ALL_MPNET_BASE_V2, "all-MiniLM-L6-v2": ALL_MINILM_L6_V2}

_BERT_LAYER_FIELDS = {
    "q_w": "attention.self.query.weight", "q_b": "attention.self.query.bias",
    "k_w": "attention.self.key.weight", "k_b": "attention.self.key.bias",
    "v_w": "attention.self.value.weight", "v_b": "attention.self.value.bias",
    "o_w": "attention.output.dense.weight", "o_b": "attention.output.dense.bias",
    "attn_ln_g": "attention.output.LayerNorm.weight", "attn_ln_b": "attention.output.LayerNorm.bias",
    "ffn_in_w": "intermediate.dense.weight", "ffn_in_b": "intermediate.dense.bias",
    "ffn_out_w": "output.dense.weight", "ffn_out_b": "output.dense.bias",
    "out_ln_g": "output.LayerNorm.weight", "out_ln_b": "output.LayerNorm.bias",
}

_LAYER_FIELDS = {
    "q_w": "attention.attn.q.weight", "q_b": "attention.attn.q.bias",
    "k_w": "attention.attn.k.weight", "k_b": "attention.attn.k.bias",
    "v_w": "attention.attn.v.weight", "v_b": "attention.attn.v.bias",
    "o_w": "attention.attn.o.weight", "o_b": "attention.attn.o.bias",
    "attn_ln_g": "attention.LayerNorm.weight", "attn_ln_b": "attention.LayerNorm.bias",
    "ffn_in_w": "intermediate.dense.weight", "ffn_in_b": "intermediate.dense.bias",
    "ffn_out_w": "output.dense.weight", "ffn_out_b": "output.dense.bias",
    "out_ln_g": "output.LayerNorm.weight", "out_ln_b": "output.LayerNorm.bias",
}
_TOP_FIELDS = {
    "word_embeddings": "embeddings.word_embeddings.weight",
    "position_embeddings": "embeddings.position_embeddings.weight",
    "emb_ln_g": "embeddings.LayerNorm.weight",
    "emb_ln_b": "embeddings.LayerNorm.bias",
    "relative_attention_bias": "encoder.relative_attention_bias.weight",
}


def _as_f32(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().float().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


class PackedWeights:
    """Owns the fp32 host arrays and the ctypes structs that point into them."""

    def __init__(self, arch: MPNetArch, state_dict: dict):
        self.arch = arch
        self._keep = []  # numpy arrays must outlive the structs

        def fp(name: str, shape: tuple):
            key = name if name in state_dict else "0.auto_model." + name
            if key not in state_dict:
                raise KeyError(f"state dict lacks '{name}'")
            a = _as_f32(state_dict[key])
            if tuple(a.shape) != shape:
                raise ValueError(f"{name}: shape {a.shape}, expected {shape}")
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_float))

        H, I = arch.hidden_size, arch.intermediate_size
        shapes = {
            "q_w": (H, H), "k_w": (H, H), "v_w": (H, H), "o_w": (H, H),
            "q_b": (H,), "k_b": (H,), "v_b": (H,), "o_b": (H,),
            "attn_ln_g": (H,), "attn_ln_b": (H,), "out_ln_g": (H,), "out_ln_b": (H,),
            "ffn_in_w": (I, H), "ffn_in_b": (I,), "ffn_out_w": (H, I), "ffn_out_b": (H,),
        }
        bert = arch.kind == "bert"
        self.layers = (_lib.MpnetLayerWeights * arch.num_layers)()
        for l in range(arch.num_layers):
            for field, suffix in (_BERT_LAYER_FIELDS if bert else _LAYER_FIELDS).items():
                setattr(self.layers[l], field, fp(f"encoder.layer.{l}.{suffix}", shapes[field]))
        self.struct = _lib.MpnetWeights()
        top_shapes = {
            "word_embeddings": (arch.vocab_size, H),
            "position_embeddings": (arch.max_position_embeddings, H),
            "emb_ln_g": (H,), "emb_ln_b": (H,),
            "relative_attention_bias": (arch.relative_attention_num_buckets, arch.num_heads),
        }
        for field, name in _TOP_FIELDS.items():
            if bert and field == "relative_attention_bias":
                continue  # BERT has no relative-position bias (NULL pointer, zero buckets)
            if bert and field == "position_embeddings":
                # BertEmbeddings adds token_type_embeddings[token_type_ids]; sentence-transformers
                # feeds token type 0 everywhere, so row 0 is folded into the position table
                key = lambda n: n if n in state_dict else "0.auto_model." + n  # noqa: E731
                pos = _as_f32(state_dict[key("embeddings.position_embeddings.weight")])
                tt = _as_f32(state_dict[key("embeddings.token_type_embeddings.weight")])
                if tuple(pos.shape) != top_shapes[field]:
                    raise ValueError(f"position_embeddings: shape {pos.shape}, expected {top_shapes[field]}")
                folded = np.ascontiguousarray(pos + tt[0][None, :], dtype=np.float32)
                self._keep.append(folded)
                setattr(self.struct, field, folded.ctypes.data_as(C.POINTER(C.c_float)))
                continue
            setattr(self.struct, field, fp(name, top_shapes[field]))
        self.struct.layers = C.cast(self.layers, C.POINTER(_lib.MpnetLayerWeights))


def heavy_tail_state_dict(arch: MPNetArch = ALL_MPNET_BASE_V2, seed: int = 0) -> dict:
    """`synthetic_state_dict` reshaped towards what trained encoders look like where it hurts
    16-bit arithmetic: two outlier hidden channels (x20 rows of every FFN down-projection, the
    "massive activation" channels of BERT-family models), LayerNorm gains up to 5 on a few
    channels, and a relative-position table spread over +-8 (23 log2 units between the most and
    the least favoured offset, which exercises the softmax shift). No checkpoint exists offline;
    this is the stand-in the parity tests run both 16-bit modes on."""
    sd = {k: np.array(v, copy=True) for k, v in synthetic_state_dict(arch, seed).items()}
    rng = np.random.default_rng(seed + 7919)
    H = arch.hidden_size
    outliers = rng.choice(H, 2, replace=False)
    key = "encoder.relative_attention_bias.weight"
    if key in sd:
        sd[key] = rng.uniform(-8.0, 8.0, sd[key].shape).astype(np.float32)
    bert = arch.kind == "bert"
    for l in range(arch.num_layers):
        p = f"encoder.layer.{l}."
        for nm in ("attention.output.LayerNorm" if bert else "attention.LayerNorm", "output.LayerNorm"):
            sd[p + nm + ".weight"][rng.choice(H, 8, replace=False)] = rng.uniform(2.0, 5.0, 8).astype(np.float32)
        sd[p + "output.dense.weight"][outliers, :] *= 20.0
        sd[p + "output.dense.bias"][outliers] *= 20.0
    return sd


def synthetic_state_dict(arch: MPNetArch = ALL_MPNET_BASE_V2, seed: int = 0) -> dict:
    """Seeded random weights with the statistics of a trained encoder's parameter classes:
    N(0, 0.02) matrices/embeddings (HF init), and NON-trivial biases, LayerNorm gamma/beta and
    relative-position table so that bias/affine bugs cannot hide (SURVEY.md §7 'hard parts').
    Returned as {HF parameter name: float32 numpy array}."""
    rng = np.random.default_rng(seed)
    H, I = arch.hidden_size, arch.intermediate_size

    def mat(*shape, std=0.02):
        return (rng.standard_normal(shape, dtype=np.float32) * std).astype(np.float32)

    if arch.kind == "bert":  # transformers.BertModel parameter names
        sd = {
            "embeddings.word_embeddings.weight": mat(arch.vocab_size, H),
            "embeddings.position_embeddings.weight": mat(arch.max_position_embeddings, H),
            "embeddings.token_type_embeddings.weight": mat(2, H),
            "embeddings.LayerNorm.weight": 1.0 + mat(H, std=0.1),
            "embeddings.LayerNorm.bias": mat(H, std=0.1),
        }
        sd["embeddings.word_embeddings.weight"][arch.pad_token_id] = 0.0
        for l in range(arch.num_layers):
            p = f"encoder.layer.{l}."
            for nm in ("self.query", "self.key", "self.value", "output.dense"):
                sd[p + f"attention.{nm}.weight"] = mat(H, H, std=0.06)
                sd[p + f"attention.{nm}.bias"] = mat(H, std=0.05)
            sd[p + "attention.output.LayerNorm.weight"] = 1.0 + mat(H, std=0.1)
            sd[p + "attention.output.LayerNorm.bias"] = mat(H, std=0.1)
            sd[p + "intermediate.dense.weight"] = mat(I, H, std=0.05)
            sd[p + "intermediate.dense.bias"] = mat(I, std=0.05)
            sd[p + "output.dense.weight"] = mat(H, I, std=0.05)
            sd[p + "output.dense.bias"] = mat(H, std=0.05)
            sd[p + "output.LayerNorm.weight"] = 1.0 + mat(H, std=0.1)
            sd[p + "output.LayerNorm.bias"] = mat(H, std=0.1)
        return sd
    sd = {
        "embeddings.word_embeddings.weight": mat(arch.vocab_size, H),
        "embeddings.position_embeddings.weight": mat(arch.max_position_embeddings, H),
        "embeddings.LayerNorm.weight": 1.0 + mat(H, std=0.1),
        "embeddings.LayerNorm.bias": mat(H, std=0.1),
        "encoder.relative_attention_bias.weight": mat(arch.relative_attention_num_buckets, arch.num_heads, std=0.5),
    }
    # HF zeroes the padding rows of both embedding tables
    sd["embeddings.word_embeddings.weight"][arch.pad_token_id] = 0.0
    sd["embeddings.position_embeddings.weight"][arch.pad_token_id] = 0.0
    for l in range(arch.num_layers):
        p = f"encoder.layer.{l}."
        for nm in ("q", "k", "v", "o"):
            # trained attention projections are O(1/sqrt(H))-scaled: use that so softmax is not flat
            sd[p + f"attention.attn.{nm}.weight"] = mat(H, H, std=0.04)
            sd[p + f"attention.attn.{nm}.bias"] = mat(H, std=0.05)
        sd[p + "attention.LayerNorm.weight"] = 1.0 + mat(H, std=0.1)
        sd[p + "attention.LayerNorm.bias"] = mat(H, std=0.1)
        sd[p + "intermediate.dense.weight"] = mat(I, H, std=0.04)
        sd[p + "intermediate.dense.bias"] = mat(I, std=0.05)
        sd[p + "output.dense.weight"] = mat(H, I, std=0.04)
        sd[p + "output.dense.bias"] = mat(H, std=0.05)
        sd[p + "output.LayerNorm.weight"] = 1.0 + mat(H, std=0.1)
        sd[p + "output.LayerNorm.bias"] = mat(H, std=0.1)
    return sd

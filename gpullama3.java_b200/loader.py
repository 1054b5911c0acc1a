"""Host-side mirror of the reference model loaders (metadata -> Configuration, tensor names ->
weight slots).  Mirrors ``model/loader/ModelLoader.java:47-108`` (type detection on
``general.name``), ``LlamaModelLoader.java:47-63`` / ``Qwen3ModelLoader.java:48-74``
(configuration keys), ``AbstractModelLoader.java:40-50`` (file_type -> quantisation) and
``AbstractModelLoader.java:186-195`` (tied output falls back to ``token_embd.weight``).
Weights stay in GGUF block layout; the native library repacks at upload.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .gguf import GGMLType, GGUFFile

ARCH_LLAMA = 0
ARCH_QWEN3 = 1
ARCH_PHI3 = 2


@dataclass
class Configuration:
    arch: int
    quantization: str  # "FP16" | "Q8_0"
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    head_size: int
    vocab_size: int
    context_length: int
    rms_norm_eps: float
    rope_theta: float

    @property
    def q_dim(self):
        return self.n_heads * self.head_size

    @property
    def kv_dim(self):
        return self.n_kv_heads * self.head_size


class UnsupportedModel(Exception):
    """Maps to the reference's UnsupportedOperationException (ForwardPlanFactory.java:84-87)."""


def detect_model_type(metadata: dict) -> str:
    name = metadata.get("general.name")
    if name is not None:
        low = name.lower()
        for key, typ in (("granite", "GRANITE"), ("devstral", "DEVSTRAL_2"), ("mistral", "MISTRAL"),
                         ("llama", "LLAMA_3"), ("qwen2", "QWEN_2"), ("qwen3", "QWEN_3"),
                         ("deepseek r1 distill", "DEEPSEEK_R1_DISTILL_QWEN"), ("phi3", "PHI_3"), ("phi-3", "PHI_3")):
            if key in low:
                return typ
    return "UNKNOWN"


def _quantization(metadata: dict) -> str:
    ft = int(metadata["general.file_type"])
    if ft == 1:
        return "FP16"
    if ft == 7:
        return "Q8_0"
    if ft in (14, 15, 16, 17, 18):  # Q4_K_S/M, Q5_K_S/M, Q6_K: the accelerator path computes in Q8_0 (AbstractModelLoader.java:45-47)
        return "Q8_0"
    raise UnsupportedModel(f"Unsupported quantization format: {ft} (as int).")


class Model:
    """Configuration + raw tensor views.  ``tensors[name] = (ggml_type, dims, uint8 ndarray)``."""

    def __init__(self, gguf: GGUFFile | None, config: Configuration, model_type: str, tensors: dict | None = None):
        self.gguf = gguf
        self.configuration = config
        self.model_type = model_type
        self.tensors = dict(tensors) if tensors is not None else {}
        if gguf is not None:
            for name, ti in gguf.tensor_infos.items():
                if name == "rope_freqs.weight":  # GGUF.java:121-124
                    continue
                self.tensors[name] = (ti.ggml_type, ti.dims, gguf.tensor_bytes(name))
        self.plan = None  # Model.setTornadoVMPlan
        self.latest_token = None


def model_from_tensors(shape, quant: int, tensors: dict, context_length: int) -> Model:
    """In-memory model (bench: synthetic weights never touch the disk)."""
    arch = {"llama": ARCH_LLAMA, "qwen3": ARCH_QWEN3, "phi3": ARCH_PHI3}[shape.arch]
    cfg = Configuration(arch, "Q8_0" if quant == GGMLType.Q8_0 else "FP16",
                        shape.dim, shape.hidden, shape.n_layers, shape.n_heads, shape.n_kv_heads, shape.head_size,
                        shape.vocab, context_length, float(shape.eps), float(shape.rope_theta))
    return Model(None, cfg, {"llama": "LLAMA_3", "qwen3": "QWEN_3", "phi3": "PHI_3"}[shape.arch], tensors)


def load_model(path: str, context_length: int = -1) -> Model:
    """``ModelLoader.loadModel(Path,int,boolean,boolean)`` (ModelLoader.java:113-120)."""
    g = GGUFFile(path)
    md = g.metadata
    typ = detect_model_type(md)
    q = _quantization(md)
    if typ == "LLAMA_3" or typ == "MISTRAL":
        # Mistral runs the same forward as Llama (Mistral.java -> InferenceCore.forwardJava; weights in the same
        # LlamaStandardWeights slots, MistralModelLoader.java:92-113).  Differences are host-side only: the context is
        # clamped to the model's (MistralModelLoader.java:45-46) and the vocabulary size may come from the token list.
        vocab = md.get("llama.vocab_size")
        if vocab is None:
            vocab = len(md["tokenizer.ggml.tokens"])
        model_ctx = int(md["llama.context_length"])
        n_heads = int(md["llama.attention.head_count"])
        dim = int(md["llama.embedding_length"])
        if typ == "MISTRAL":
            ctx = model_ctx if (context_length < 0 or model_ctx < context_length) else context_length
        else:  # withContextLength(contextLength): LlamaConfiguration keeps the requested length when >= 0
            ctx = model_ctx if context_length < 0 else context_length
        cfg = Configuration(
            ARCH_LLAMA, q, dim, int(md["llama.feed_forward_length"]), int(md["llama.block_count"]), n_heads,
            int(md.get("llama.attention.head_count_kv", n_heads)), dim // n_heads, int(vocab), ctx,
            float(md.get("llama.attention.layer_norm_rms_epsilon", 1e-5)), float(md.get("llama.rope.freq_base", 10000.0)))
    elif typ == "QWEN_3":
        model_ctx = int(md["qwen3.context_length"])
        ctx = model_ctx if (context_length < 0 or model_ctx < context_length) else context_length
        vocab = md.get("qwen3.vocab_size")
        if vocab is None:
            vocab = len(md["tokenizer.ggml.tokens"])
        n_heads = int(md["qwen3.attention.head_count"])
        if int(md["qwen3.attention.key_length"]) != int(md["qwen3.attention.value_length"]):
            raise UnsupportedModel("key_length != value_length")
        cfg = Configuration(
            ARCH_QWEN3, q, int(md["qwen3.embedding_length"]), int(md["qwen3.feed_forward_length"]),
            int(md["qwen3.block_count"]), n_heads, int(md.get("qwen3.attention.head_count_kv", n_heads)),
            int(md["qwen3.attention.key_length"]), int(vocab), ctx,
            float(md["qwen3.attention.layer_norm_rms_epsilon"]), float(md["qwen3.rope.freq_base"]))
    elif typ == "PHI_3":
        # Phi3ModelLoader.createConfiguration (Phi3ModelLoader.java:51-71): head size = dim / heads, the context is the requested one
        # (the RoPE table is precomputed for the model's), the vocabulary size is the token list's.
        n_heads = int(md["phi3.attention.head_count"])
        dim = int(md["phi3.embedding_length"])
        model_ctx = int(md["phi3.context_length"])
        vocab = len(md["tokenizer.ggml.tokens"]) if "tokenizer.ggml.tokens" in md else int(md["phi3.vocab_size"])
        cfg = Configuration(
            ARCH_PHI3, q, dim, int(md["phi3.feed_forward_length"]), int(md["phi3.block_count"]), n_heads,
            int(md.get("phi3.attention.head_count_kv", n_heads)), dim // n_heads, int(vocab), model_ctx if context_length < 0 else context_length,
            float(md.get("phi3.attention.layer_norm_rms_epsilon", 1e-5)), float(md.get("phi3.rope.freq_base", 10000.0)))
    else:
        raise UnsupportedModel(f"model type {typ} is outside the B200 hot-path scope (Llama / Mistral / Qwen3 / Phi-3 forward passes only)")
    return Model(g, cfg, typ)


def tensor_as_f32(model: Model, name: str) -> np.ndarray:
    """Dequantise a whole tensor the way ``FloatTensor.getFloat`` does (tests / debugging)."""
    tt, dims, raw = model.tensors[name]
    if tt == GGMLType.F32:
        return raw.view("<f4").copy()
    if tt == GGMLType.F16:
        return raw.view("<f2").astype(np.float32)
    blocks = raw.reshape(-1, 34)
    d = blocks[:, :2].copy().view("<f2").astype(np.float32)
    q = blocks[:, 2:].view(np.int8).astype(np.float32)
    return (q * d).reshape(-1)

"""Host-side mirror of the reference's plan interface.

``B200MasterPlan`` stands where ``TornadoVMMasterPlan`` does
(``tornadovm/TornadoVMMasterPlan.java:30-85``): same factory, same three forward entry
points, same ``free`` -- but each call is one C-ABI call into libb200llama.so instead of
N+2 TornadoVM TaskGraph executions (``TornadoVMMasterPlanSingleToken.java:68-95``).
The flag names follow the reference's system properties (``TornadoVMMasterPlan.java:32-41``).
"""
from __future__ import annotations

import os

import numpy as np

from . import native
from .loader import Model


def _flag(name: str, default: str) -> str:
    # -Dllama.xxx system properties become LLAMA_XXX environment variables here
    return os.environ.get(name.upper().replace(".", "_"), default)


WITH_PREFILL_DECODE = _flag("llama.withPrefillDecode", "false").lower() == "true"
PREFILL_BATCH_SIZE = int(_flag("llama.prefillBatchSize", "1"))
FP16_LANES = int(_flag("llama.VectorBitSize", "512")) // 32  # FloatTensor.java:21 (species width / 32-bit lanes)


def tp_shard_plan(c, n: int) -> list[dict]:
    """Row ranges every rank owns under n-way tensor parallelism (mirrors csrc/plan.cu).  Every matrix is
    split by OUTPUT rows -- query/KV heads, FFN hidden units, residual rows, vocabulary rows -- so each dot
    product keeps its full column range and therefore the reference's summation order; what the usual
    column split turns into an all-reduce is an all-gather of the output slices here."""
    if n < 1 or n > 8 or c.n_heads % n or c.n_kv_heads % n or c.dim % (4 * n) or c.hidden_dim % (32 * n) or c.vocab_size % (4 * n):
        raise native.UnsupportedOperation(-2, f"shape does not split {n} ways")
    hs = c.head_size
    out = []
    for r in range(n):
        nh, nkv = c.n_heads // n, c.n_kv_heads // n
        out.append({
            "q_rows": (r * nh * hs, (r + 1) * nh * hs), "kv_rows": (r * nkv * hs, (r + 1) * nkv * hs),
            "heads": (r * nh, (r + 1) * nh), "kv_heads": (r * nkv, (r + 1) * nkv),
            "residual_rows": (r * c.dim // n, (r + 1) * c.dim // n),          # rows of Wo and W2
            "hidden_units": (r * c.hidden_dim // n, (r + 1) * c.hidden_dim // n),  # rows of gate/up
            "vocab_rows": (r * c.vocab_size // n, (r + 1) * c.vocab_size // n),
        })
    return out


def exchange_handles(handle: bytes, group=None) -> list[bytes]:
    """All-gather the 64-byte CUDA IPC handles in rank order (any torch.distributed backend)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, handle, group=group)
    return out


def make_config(model: Model, fp16_lanes: int | None = None) -> native.Config:
    c = model.configuration
    cfg = native.Config()
    cfg.arch = c.arch
    cfg.dim, cfg.hidden_dim, cfg.n_layers = c.dim, c.hidden_dim, c.n_layers
    cfg.n_heads, cfg.n_kv_heads, cfg.head_size = c.n_heads, c.n_kv_heads, c.head_size
    cfg.vocab_size, cfg.context_length = c.vocab_size, c.context_length
    cfg.rms_norm_eps, cfg.rope_theta = c.rms_norm_eps, c.rope_theta
    cfg.fp16_lanes = FP16_LANES if fp16_lanes is None else fp16_lanes
    cfg.tp_rank, cfg.tp_size = 0, 1
    return cfg


class B200MasterPlan:
    """One plan per model, used from one thread at a time (InferenceService.java:31,58)."""

    def __init__(self, model: Model, prefill_batch_size: int | None = None, device: int = 0, fp16_lanes: int | None = None,
                 tp_rank: int = 0, tp_size: int = 1, tp_group=None):
        self.model = model
        self.prefill_batch_size = PREFILL_BATCH_SIZE if prefill_batch_size is None else prefill_batch_size
        cfg = make_config(model, fp16_lanes)
        cfg.tp_rank, cfg.tp_size = tp_rank, tp_size
        if tp_size > 1:
            tp_shard_plan(model.configuration, tp_size)  # raises early on shapes that do not split
        self._native = native.NativePlan(cfg, model.tensors, self.prefill_batch_size, device)
        self.tp_rank, self.tp_size = tp_rank, tp_size
        if tp_size > 1:
            # one process per GPU: swap IPC handles of the communication buffers, then wire the peers
            import torch.distributed as dist

            handles = exchange_handles(self._native.tp_handle(), tp_group)
            self._native.tp_attach(handles)
            dist.barrier(group=tp_group)

    # TornadoVMMasterPlan.initializeTornadoVMPlan(state, model)  (TornadoVMMasterPlan.java:55-70)
    @staticmethod
    def initialize_plan(model: Model, **kw) -> "B200MasterPlan":
        plan = B200MasterPlan(model, **kw)
        model.plan = plan  # model.setTornadoVMPlan(plan)
        return plan

    # FloatArray tornadoVMForwardDecode(int position) + the embedding gather of
    # InferenceCore.forwardTornadoVM (InferenceCore.java:956-980)
    def forward_decode(self, token: int, position: int, logits: bool = True):
        lg, am = self._native.forward_decode(token, position, want_logits=logits, want_argmax=True)
        return lg, am

    # forward + Sampler.sampleToken on the device (Sampler.java:74-122): 4 bytes in (the uniform number), 4 bytes out
    def forward_decode_sample(self, token: int, position: int, temperature: float, topp: float, uniform01: float, want_info: bool = False):
        return self._native.forward_decode_sample(token, position, temperature, topp, uniform01, want_info)

    # void tornadoVMForwardPrefill(int position)  (TornadoVMMasterPlanPrefillDecode.java:116)
    def forward_prefill(self, token: int, position: int):
        self._native.forward_prefill(token, position)

    # void tornadoVMForwardBatchPrefill()  (TornadoVMMasterPlanBatchPrefillDecode.java:107-123)
    def forward_batch_prefill(self, tokens, start_pos: int):
        self._native.forward_batch_prefill(np.asarray(tokens, dtype=np.int32), start_pos)

    # TensorCoreSupport.java's switch between the MMA and the plain batch-prefill layer families
    PREFILL_EXACT, PREFILL_TENSOR_CORE = 0, 1

    def set_prefill_mode(self, mode):
        """"exact" (token-by-token graph, bit-identical KV cache) or "tensor_core" (tcgen05 GEMMs, FP16 tolerance)."""
        if isinstance(mode, str):
            mode = {"exact": 0, "tensor_core": 1}[mode]
        self._native.set_prefill_mode(int(mode))

    def prefill_info(self):
        """(active mode, kernels launched, device ms) of the last tensor-core prefill chunk."""
        return self._native.prefill_info()

    # -Dllama.cudaGraphs-style switch of the decode implementation (both bit-identical): "graph" = one CUDA graph of
    # ~7 kernels per layer, "persistent" = one persistent kernel per token (csrc/decode_persistent.cuh)
    DECODE_GRAPH, DECODE_PERSISTENT = 0, 1

    def set_decode_mode(self, mode):
        if isinstance(mode, str):
            mode = {"graph": 0, "persistent": 1}[mode]
        self._native.set_decode_mode(int(mode))

    def decode_info(self):
        return self._native.decode_info()

    def trace_persistent(self, token: int, position: int):
        return self._native.trace_persistent(token, position)

    def decode_sequence(self, tokens, n: int, start_pos: int, feedback: bool = False):
        return self._native.decode_sequence(tokens, n, start_pos, feedback)

    def time_kernel(self, which: int, reps: int = 3):
        return self._native.time_kernel(which, reps)

    def kv_reset(self):
        self._native.kv_reset()

    def read_buffer(self, *a, **kw):
        return self._native.read_buffer(*a, **kw)

    def upload_info(self):
        """Seconds / bytes of the weight upload pipeline of plan creation (load-time metric, ModelLoader.java:102-106)."""
        return self._native.upload_info()

    @property
    def launches_per_decode(self):
        return self._native.launches_per_decode

    @property
    def device_bytes(self):
        return self._native.device_bytes

    # void freeTornadoExecutionPlan()
    def free(self):
        self._native.free()

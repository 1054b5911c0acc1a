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

    def __init__(self, model: Model, prefill_batch_size: int | None = None, device: int = 0, fp16_lanes: int | None = None):
        self.model = model
        self.prefill_batch_size = PREFILL_BATCH_SIZE if prefill_batch_size is None else prefill_batch_size
        self._native = native.NativePlan(make_config(model, fp16_lanes), model.tensors, self.prefill_batch_size, device)

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

    # void tornadoVMForwardPrefill(int position)  (TornadoVMMasterPlanPrefillDecode.java:116)
    def forward_prefill(self, token: int, position: int):
        self._native.forward_prefill(token, position)

    # void tornadoVMForwardBatchPrefill()  (TornadoVMMasterPlanBatchPrefillDecode.java:107-123)
    def forward_batch_prefill(self, tokens, start_pos: int):
        self._native.forward_batch_prefill(np.asarray(tokens, dtype=np.int32), start_pos)

    def decode_sequence(self, tokens, n: int, start_pos: int, feedback: bool = False):
        return self._native.decode_sequence(tokens, n, start_pos, feedback)

    def time_kernel(self, which: int, reps: int = 3):
        return self._native.time_kernel(which, reps)

    def kv_reset(self):
        self._native.kv_reset()

    def read_buffer(self, *a, **kw):
        return self._native.read_buffer(*a, **kw)

    @property
    def launches_per_decode(self):
        return self._native.launches_per_decode

    @property
    def device_bytes(self):
        return self._native.device_bytes

    # void freeTornadoExecutionPlan()
    def free(self):
        self._native.free()

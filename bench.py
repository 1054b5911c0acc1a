#!/usr/bin/env python
"""bench.py -- Llama-3-8B Q8_0 single-stream decode (tg) on B200, the BASELINE.json headline.

A "step" is one single-token decode forward (all layers + lm_head + on-device argmax) at a
growing KV position, on the LlamaBench synthetic token stream (`new Random(42).nextInt(vocab)`,
LlamaBench.java:188-193) over a seeded synthetic GGUF-layout model of the real Llama-3-8B shape.

  value        tok/s with tokens already resident in HBM (device loop b200_decode_sequence,
               CUDA-event time on the plan's stream, max over ranks)
  e2e          tok/s through the reference-facing call b200_forward_decode with HOST buffers:
               every step copies the token/position H2D and the argmax D2H inside the timed region
  roofline     dominant kernel (fused gate/up dequant-matvec) timed stand-alone with CUDA events;
               achieved = algorithmic bytes / avg launch time, peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the oracle (CPU restatement of the reference's onGPU=false path) on this box's
               host cores, bounded sample
`--impl reference` times only that CPU restatement (the reference itself needs a JDK + TornadoVM,
neither is installable here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

WORKLOAD = "llama-3-8b"
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return FALLBACK_HBM_GBS, "fallback"


FALLBACK_TENSOR_TFLOPS = 2250.0  # nominal dense bf16/fp16


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (cuBLAS bf16 8192^3; fp16 shares the rate)"
    return FALLBACK_TENSOR_TFLOPS, FALLBACK_TENSOR_TFLOPS, "fallback (nominal)"


def prefill_leg(pkg, lb, local: int, n: int, reps: int):
    """BASELINE config 3: Llama-3-8B FP16, --batch-prefill-size 512, pp512 from depth 0 (LlamaBench `pp`
    semantics: forward only, no logits).  Tensor-core path: TMA + tcgen05 GEMMs, csrc/prefill*.cuh."""
    shape = pkg.synth.SHAPES[WORKLOAD]
    F16 = pkg.gguf.GGMLType.F16
    model = pkg.loader.model_from_tensors(shape, F16, pkg.synth.build_tensors_fast(shape, F16, seed=1234, device=f"cuda:{local}"), n + 8)
    plan = pkg.B200MasterPlan.initialize_plan(model, prefill_batch_size=n, device=local)
    toks = np.asarray(lb.synthetic_tokens(shape.vocab, n), dtype=np.int32)
    mode = plan.prefill_info()[0]
    dev, wall = [], []
    for r in range(3 + reps):  # 3 untimed warm-up chunks
        t0 = time.perf_counter()
        plan.forward_batch_prefill(toks, 0)  # host tokens in, synchronous: the e2e call
        t1 = time.perf_counter()
        if r >= 3:
            dev.append(plan.prefill_info()[2])
            wall.append((t1 - t0) * 1e3)
    launches = plan.prefill_info()[1]
    plan.free()
    d, w = float(np.mean(dev)), float(np.mean(wall))
    gemm = 2.0 * shape.matmul_elements_no_head() * n
    att = 4.0 * shape.q_dim * shape.n_layers * (n * (n + 1) / 2.0)
    burst, sustained, src = tensor_peak()
    tf = (gemm + att) / (d * 1e-3) / 1e12
    return {"metric": "prefill_tokens_per_s", "value": n / d * 1e3, "unit": "tok/s", "ms_per_chunk": d, "reps": reps, "dtype": "f16 operands, f32 accumulate (TMEM)",
            "e2e": {"value": n / w * 1e3, "unit": "tok/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 0},
            "gpu_launches": launches, "mode": "tensor_core" if mode == 1 else "exact",
            "config": {"workload": f"Llama-3-8B-shaped synthetic GGUF, FP16, pp{n} in one chunk (--batch-prefill-size {n}) from depth 0, KV cache only (no logits)",
                       "l2": "inputs larger than L2 (15.0 GB of FP16 weights per chunk)"},
            "roofline": {"kernel": "whole prefill chunk (k_gemm_f16_tcgen05 = 85 % of it, profiles/)", "bound": "tensor", "achieved": tf, "peak": burst, "peak_sustained": sustained,
                         "peak_source": src, "unit": "TFLOP/s", "frac": tf / burst, "flop_per_chunk": {"gemm": gemm, "attention": att}, "traffic": None}}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes_per_token(shape, quant_q8: bool, ctx_avg_pos: float) -> dict:
    """matmul weights read once + F32 norm weights + FP32 KV reads."""
    elems = shape.matmul_elements()
    w = elems // 32 * 34 if quant_q8 else elems * 2
    norms = (2 * shape.n_layers + 1) * shape.dim * 4
    kv = 2 * shape.n_layers * (ctx_avg_pos + 1) * shape.kv_dim * 4
    return {"weights": w, "norms": norms, "kv": kv, "total": w + norms + kv}


def build_model(pkg, ctx: int, device: str | None):
    shape = pkg.synth.SHAPES[WORKLOAD]
    quant = pkg.gguf.GGMLType.Q8_0
    t0 = time.time()
    tensors = pkg.synth.build_tensors_fast(shape, quant, seed=1234, device=device)
    model = pkg.loader.model_from_tensors(shape, quant, tensors, ctx)
    return shape, model, time.time() - t0


def cpu_leg(orc, model, tokens, budget_s: float, max_tokens: int):
    """Times the CPU restatement (per-row activation quantisation exactly as the reference does,
    Q8_0FloatTensor.java:100-117; rows over all host cores like Parallel.parallelFor)."""
    om = orc.OracleModel(model, per_row_quant=True)
    cores = int(orc.lib().oracle_omp_threads())
    om.forward(int(tokens[0]), 0)  # untimed warm-up (page-in)
    n, t0 = 0, time.perf_counter()
    while n < max_tokens:
        om.forward(int(tokens[1 + n]), 1 + n)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    om.close()
    return n / dt, n, dt, cores


def main():
    global WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-baseline work")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-pp", action="store_true", help="skip the pp512 tensor-core prefill leg")
    ap.add_argument("--workload", default=WORKLOAD, choices=["llama-3-8b", "llama-3-70b", "llama-3.2-1b", "qwen3-4b"],
                    help="shape of the synthetic model (default: the BASELINE headline, Llama-3-8B; 70B is BASELINE config 5, meant for --gpus 2/4/8)")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 3)
    WORKLOAD = args.workload
    pretty = {"llama-3-8b": "Llama-3-8B", "llama-3-70b": "Llama-3-70B", "llama-3.2-1b": "Llama-3.2-1B", "qwen3-4b": "Qwen3-4B"}[WORKLOAD]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = ge.import_package()
    lb = pkg.llama_bench
    shape = pkg.synth.SHAPES[WORKLOAD]
    ctx = W + K + 8  # LlamaBench: depth + tokens + 8 (LlamaBench.java:174)
    tokens = np.asarray(lb.synthetic_tokens(shape.vocab, W + K), dtype=np.int32)
    config = {"workload": f"{pretty}-shaped synthetic GGUF, Q8_0, tg{K} single-stream decode from depth {W}",
              "weights": "seeded N(0,1/sqrt(fan_in)) quantised with the ggml Q8_0 rule; tokens java.util.Random(42)",
              "context": ctx, "l2": f"inputs larger than L2 ({shape.matmul_elements() // 32 * 34 / 1e9:.2f} GB of weights stream per step vs 126 MB L2)"}

    if args.impl == "reference":
        if world > 1 and rank != 0:
            return 0
        orc = ge.import_oracle()
        _, model, _ = build_model(pkg, ctx, None)
        budget = max(10.0, min(120.0, 8.0 * (K + W)))  # bounded sample: the whole run ends within minutes
        tps, n, dt, cores = cpu_leg(orc, model, tokens, budget, max(1, min(K, W + K - 1)))
        line = {"metric": "decode_tokens_per_s", "value": tps, "unit": "tok/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 / tps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "q8_0xq8_0->int32, f32 accumulate",
                "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port",
                                 "sample": f"{n} decode steps of the same workload in {dt:.1f} s (C restatement of InferenceCore.forwardJava, -O2, OpenMP rows)"},
                "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    shape, model, gen_s = build_model(pkg, ctx, f"cuda:{local}")
    t0 = time.time()
    plan = pkg.B200MasterPlan.initialize_plan(model, device=local, tp_rank=rank, tp_size=world)
    load_s = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    # ---- value: device loop, tokens resident in HBM ------------------------------------------
    plan.decode_sequence(tokens[:W], W, 0)  # W untimed warm-up steps (positions 0..W-1)
    barrier()
    if rank == 0:
        sampler.start()
    ids, ms = plan.decode_sequence(tokens[W:W + K], K, W)
    barrier()
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = K / (ms / 1e3)  # one stream; under --gpus N the model is tensor-parallel over N ranks (strong scaling)

    # ---- e2e: reference-facing call, host token in / host argmax out every step ----------------
    plan.kv_reset()
    for i in range(W):
        plan.forward_decode(int(tokens[i]), i, logits=False)
    barrier()
    t0 = time.perf_counter()
    e2e_ids = []
    for i in range(K):
        _, am = plan.forward_decode(int(tokens[W + i]), W + i, logits=False)
        e2e_ids.append(am)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    assert list(ids) == e2e_ids, "device loop and per-step API disagree"

    if rank != 0:
        plan.free()
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peak, peak_src = peaks()
    ab = algorithmic_bytes_per_token(shape, True, W + (K - 1) / 2.0)
    step_gbs = ab["total"] / world * value / 1e9  # per-GPU bytes per token x tok/s
    per_kernel = {}
    if world == 1:
        k_ms, k_bytes = plan.time_kernel(0, reps=3)
        achieved = k_bytes / (k_ms / 1e3) / 1e9
        for which, name in ((1, "down_proj"), (2, "qkv"), (3, "attn_out"), (4, "lm_head")):
            m, b = plan.time_kernel(which, reps=3 if which != 4 else 1)
            per_kernel[name] = {"ms": m, "GB/s": b / (m / 1e3) / 1e9}
    else:  # the stand-alone kernel timer is single-GPU; report the whole-step figure per GPU
        k_ms, k_bytes, achieved = None, None, step_gbs
    line = {
        "metric": "decode_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
        "dtype": "q8_0xq8_0->int32, f32 accumulate", "data": "synthetic", "config": config,
        "parallelism": "single GPU" if world == 1 else f"tp{world}: row-sharded weights, in-kernel all-gather over NVLink peer memory (bit-exact with tp1)",
        "e2e": {"value": K / e2e_s, "unit": "tok/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 4},
        "gpu_launches": plan.launches_per_decode * K,
        "clocks": clocks,
        "roofline": {"kernel": "k_stream_matvec_q8<GATEUP> (TMA-ring fused gate/up dequant-matvec + SwiGLU + Q8_0 requantise), timed stand-alone without PDL prefetch" if world == 1 else "whole decode step per GPU",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "bytes_per_launch": k_bytes, "ms_per_launch": k_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, ncu --set full, profiles/r1_stream_matvec_full.ncu-rep
                     "traffic": 124823296 + 3408640 if (world == 1 and WORKLOAD == "llama-3-8b") else None,
                     "whole_step": {"algorithmic_bytes_per_token": ab, "achieved": step_gbs, "frac": step_gbs / peak},
                     "other_kernels": per_kernel},
        "load": {"synthesise_s": gen_s, "upload_repack_s": load_s, "device_bytes": plan.device_bytes},
    }
    if not args.no_cpu and world == 1:
        orc = ge.import_oracle()
        tps, n, dt, cores = cpu_leg(orc, model, tokens, args.cpu_budget, 16)
        line["cpu_baseline"] = {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port",
                                "sample": f"{n} decode steps of the same workload in {dt:.1f} s (C restatement of InferenceCore.forwardJava, -O2, OpenMP rows)"}
    plan.free()
    if not args.no_pp and world == 1 and WORKLOAD == "llama-3-8b":  # BASELINE config 3 is the 8B FP16 model
        del model, plan
        line["pp512"] = prefill_leg(pkg, lb, local, 512, 5)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

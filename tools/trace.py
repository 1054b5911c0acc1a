"""In-graph timeline of one decode step (PDL on) for the 8B-shaped model: per-launch
entry / dependency-wait return / exit times from %globaltimer.  Usage: python tools/trace.py [shape] [pos] [q8_0|f16]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.import_package()
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
pos = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sh = pkg.synth.SHAPES[shape]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
qt = 1 if (len(sys.argv) > 3 and sys.argv[3] == "f16") else 8
m = pkg.loader.model_from_tensors(sh, qt, pkg.synth.build_tensors_fast(sh, qt, seed=3, device=f"cuda:{local}"), pos + 16)
plan = pkg.B200MasterPlan.initialize_plan(m, device=local, tp_rank=rank, tp_size=world)
toks = pkg.llama_bench.synthetic_tokens(sh.vocab, pos + 2)
plan.decode_sequence(toks[:pos], pos, 0)
names = {1: "rmsnorm", 2: "qkv", 3: "rope_kv", 4: "attention", 5: "attn_out", 6: "gate_up", 7: "down", 8: "lm_head", 9: "argmax"}
for rep in range(2):
    rec = plan._native.trace_decode(toks[pos], pos).astype(np.int64)
if rank != 0:
    plan.free(); dist.barrier(); dist.destroy_process_group(); sys.exit(0)
t0 = rec[0, 1]
print(f"{'#':>4} {'kernel':10} {'entry':>9} {'wait_ret':>9} {'exit':>9} {'run(us)':>8} {'resident_before_dep(us)':>22} {'gap_prev_exit->wait_ret':>24}")
prev_exit = None
agg = {}
for i, (kid, te, tw, tx) in enumerate(rec):
    run = (tx - tw) / 1e3
    early = (tw - te) / 1e3
    gap = (tw - prev_exit) / 1e3 if prev_exit is not None else 0.0
    if 7 * 8 <= i < 7 * 10 or i >= len(rec) - 4 or i < 8:
        print(f"{i:4d} {names.get(int(kid), '?'):10} {(te - t0) / 1e3:9.2f} {(tw - t0) / 1e3:9.2f} {(tx - t0) / 1e3:9.2f} {run:8.2f} {early:22.2f} {gap:24.2f}")
    a = agg.setdefault(int(kid), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += run; a[2] += early; a[3] += gap
    prev_exit = tx
print("\nper kernel type: launches, mean run (wait_ret->exit), mean residency before dependency, mean gap prev_exit->wait_ret [us]")
for k, a in sorted(agg.items()):
    print(f"  {names.get(k,'?'):10} n={a[0]:4d} run={a[1]/a[0]:8.2f} early={a[2]/a[0]:8.2f} gap={a[3]/a[0]:6.2f}  total_run={a[1]:9.1f}")
print(f"step total: {(rec[-1,3]-rec[0,1])/1e3:.1f} us; sum(run)={sum(a[1] for a in agg.values()):.1f} us; sum(gap)={sum(a[3] for a in agg.values()):.1f} us")
plan.free()
if world > 1:
    dist.barrier(); dist.destroy_process_group()

"""Tensor-parallel parity check, one process per GPU:
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py [shape] [steps]
Every rank builds the same seeded model, uploads its share, decodes a teacher-forced stream and a greedy
stream; rank 0 compares every argmax with the CPU oracle (token-exact expected: row sharding keeps the
reference's summation order) and the residual stream bit for bit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import __graft_entry__ as ge
pkg = ge.import_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
shape = sys.argv[1] if len(sys.argv) > 1 else "mid-llama"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sh = pkg.synth.SHAPES[shape]
m = pkg.loader.model_from_tensors(sh, 8, pkg.synth.build_tensors_fast(sh, 8, seed=5, device=f"cuda:{local}"), steps + 8)
plan = pkg.B200MasterPlan(m, device=local, tp_rank=rank, tp_size=world)
toks = pkg.llama_bench.synthetic_tokens(sh.vocab, steps)
ids, ms = plan.decode_sequence(toks, steps, 0)
x = plan.read_buffer("x", sh.dim)
ok = True
if rank == 0:
    orc = ge.import_oracle()
    om = orc.OracleModel(m)
    ref = [orc.argmax(om.forward(int(toks[p]), p)) for p in range(steps)]
    ok &= list(ids) == ref
    print(f"[tp{world}] teacher-forced argmax == oracle: {list(ids) == ref}  ({steps} steps, {ms / steps * 1e3:.1f} us/token)", flush=True)
    om.reset()
    for p in range(steps):
        om.forward(int(toks[p]), p, want_logits=False)
    same = np.array_equal(om.x().view(np.uint32), x.view(np.uint32))
    # (the oracle's x after a no-logits forward is the residual stream after the last layer, as is ours)
    print(f"[tp{world}] residual stream bit-exact vs oracle: {same}", flush=True)
    ok &= same
plan.kv_reset()
g, _ = plan.decode_sequence(toks[:1], steps, 0, feedback=True)
all_g = [None] * world
dist.all_gather_object(all_g, [int(t) for t in g])
if rank == 0:
    agree = all(a == all_g[0] for a in all_g)
    om.reset(); tok, refg = int(toks[0]), []
    for p in range(steps):
        tok = orc.argmax(om.forward(tok, p)); refg.append(tok)
    print(f"[tp{world}] greedy loop: ranks agree {agree}, == oracle {all_g[0] == refg}", flush=True)
    ok &= agree and all_g[0] == refg
    print("RESULT", "OK" if ok else "MISMATCH", flush=True)
plan.free()
dist.barrier()
dist.destroy_process_group()

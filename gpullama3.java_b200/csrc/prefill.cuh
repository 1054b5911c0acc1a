// prefill.cuh -- batched prefill (tensor-core) path.  Placeholder until the tcgen05 GEMM
// pipeline lands: b200_forward_batch_prefill currently runs the exact single-token graph
// per token (bit-identical KV cache), see plan.cu.
#pragma once
#include "../../include/b200llama.h"

struct PrefillCtx {
    int batch = 0;
};
inline int prefill_init(PrefillCtx &c, const b200_config &, int batch) { c.batch = batch; return B200_OK; }
inline void prefill_free(PrefillCtx &) {}

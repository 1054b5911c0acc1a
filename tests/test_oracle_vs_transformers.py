"""STRUCTURAL pin of the oracle against an independent implementation (CPU).

No JDK exists in this image, so the oracle cannot be compared with the reference's own output ("parity unpinned").  What can
be checked here is everything except the float summation order: the oracle's forward pass (oracle/oracle.c, a restatement of
InferenceCore.forwardJava / forwardJavaQwen3, inference/InferenceCore.java:39-172,565-697) is run on a seeded synthetic GGUF model and
compared with Hugging Face transformers' LlamaForCausalLM / Qwen3ForCausalLM / Phi3ForCausalLM evaluating THE SAME weights in float64 -- an implementation
that shares no code with the reference or with this repository.  Any structural mistake (RoPE pairing or frequency, GQA head mapping, norm
placement, q/k-norm, SwiGLU operand order, tied classifier, KV-cache indexing across positions) produces O(1) errors; agreement is at
rounding level: <= 2e-4 of max|logit| for FP16 weights (fp32 arithmetic vs float64), <= 5e-2 for Q8_0 weights (the CPU path additionally
rounds every activation block to int8, Q8_0FloatTensor.vectorDot -- that is the algorithm, not an error).

Weight layout: GGUF Llama files hold Wq / Wk with the rows of each head permuted for interleaved-pair RoPE (llama.cpp's convert script),
which is what the reference's forwardJava applies (InferenceCore.java:75-87); transformers uses rotate-half, so the rows are permuted back
here.  Qwen3 GGUF files keep the rotate-half (NeoX) layout, as does forwardJavaQwen3 (:604-619)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def _unpermute(w, n_head):
    """Inverse of llama.cpp's permute(): GGUF (interleaved pairs 2i, 2i+1) -> HF (i, i + head/2)."""
    rows, cols = w.shape
    hs = rows // n_head
    return w.reshape(n_head, hs // 2, 2, cols).swapaxes(1, 2).reshape(rows, cols)


def _hf_model(pkg, m):
    c = m.configuration
    qwen = m.model_type == "QWEN_3"
    common = dict(hidden_size=c.dim, intermediate_size=c.hidden_dim, num_hidden_layers=c.n_layers, num_attention_heads=c.n_heads,
                  num_key_value_heads=c.n_kv_heads, vocab_size=c.vocab_size, rms_norm_eps=c.rms_norm_eps, max_position_embeddings=c.context_length,
                  tie_word_embeddings=False, rope_theta=c.rope_theta, attention_bias=False, head_dim=c.head_size)
    phi3 = m.model_type == "PHI_3"
    if qwen:
        hf = transformers.Qwen3ForCausalLM(transformers.Qwen3Config(**common))
    elif phi3:
        common.pop("head_dim"), common.pop("attention_bias")
        hf = transformers.Phi3ForCausalLM(transformers.Phi3Config(**common, original_max_position_embeddings=c.context_length, pad_token_id=0))
    else:
        hf = transformers.LlamaForCausalLM(transformers.LlamaConfig(**common, mlp_bias=False))

    def W(name, rows, cols):
        return pkg.loader.tensor_as_f32(m, name).reshape(rows, cols).astype(np.float64)

    def V(name):
        return pkg.loader.tensor_as_f32(m, name).astype(np.float64)
    sd = {"model.embed_tokens.weight": W("token_embd.weight", c.vocab_size, c.dim), "model.norm.weight": V("output_norm.weight")}
    sd["lm_head.weight"] = W("output.weight", c.vocab_size, c.dim) if "output.weight" in m.tensors else sd["model.embed_tokens.weight"]
    qd, kvd = c.n_heads * c.head_size, c.n_kv_heads * c.head_size
    for l in range(c.n_layers):
        g, h = f"blk.{l}.", f"model.layers.{l}."
        if phi3:  # transformers keeps Phi-3's fused layout too: qkv_proj = [q; k; v] rows, gate_up_proj = [gate; up] rows, rotate-half RoPE
            sd[h + "self_attn.qkv_proj.weight"] = W(g + "attn_qkv.weight", qd + 2 * kvd, c.dim)
            sd[h + "self_attn.o_proj.weight"] = W(g + "attn_output.weight", c.dim, qd)
            sd[h + "mlp.gate_up_proj.weight"] = W(g + "ffn_up.weight", 2 * c.hidden_dim, c.dim)
            sd[h + "mlp.down_proj.weight"] = W(g + "ffn_down.weight", c.dim, c.hidden_dim)
            sd[h + "input_layernorm.weight"] = V(g + "attn_norm.weight")
            sd[h + "post_attention_layernorm.weight"] = V(g + "ffn_norm.weight")
            continue
        wq, wk = W(g + "attn_q.weight", qd, c.dim), W(g + "attn_k.weight", kvd, c.dim)
        if not qwen:
            wq, wk = _unpermute(wq, c.n_heads), _unpermute(wk, c.n_kv_heads)
        sd[h + "self_attn.q_proj.weight"], sd[h + "self_attn.k_proj.weight"] = wq, wk
        sd[h + "self_attn.v_proj.weight"] = W(g + "attn_v.weight", kvd, c.dim)
        sd[h + "self_attn.o_proj.weight"] = W(g + "attn_output.weight", c.dim, qd)
        sd[h + "mlp.gate_proj.weight"] = W(g + "ffn_gate.weight", c.hidden_dim, c.dim)
        sd[h + "mlp.up_proj.weight"] = W(g + "ffn_up.weight", c.hidden_dim, c.dim)
        sd[h + "mlp.down_proj.weight"] = W(g + "ffn_down.weight", c.dim, c.hidden_dim)
        sd[h + "input_layernorm.weight"] = V(g + "attn_norm.weight")
        sd[h + "post_attention_layernorm.weight"] = V(g + "ffn_norm.weight")
        if qwen:
            sd[h + "self_attn.q_norm.weight"] = V(g + "attn_q_norm.weight")
            sd[h + "self_attn.k_norm.weight"] = V(g + "attn_k_norm.weight")
    hf = hf.to(torch.float64)
    missing, unexpected = hf.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return hf.eval()


@pytest.mark.parametrize("shape,quant,tol", [("tiny-llama", "F16", 2e-4), ("tiny-llama-tied", "F16", 2e-4), ("tiny-qwen3", "F16", 2e-4),
                                             ("tiny-phi3", "F16", 2e-4), ("tiny-phi3-gqa", "F16", 2e-4),
                                             ("tiny-llama", "Q8_0", 5e-2), ("tiny-qwen3", "Q8_0", 5e-2), ("tiny-phi3", "Q8_0", 5e-2)])
def test_oracle_forward_agrees_with_transformers(pkg, orc, make_model, shape, quant, tol):
    n_tok = 20
    m = make_model(shape, getattr(pkg.gguf.GGMLType, quant), 32)
    c = m.configuration
    toks = orc.bench_tokens(c.vocab_size, n_tok)
    om = orc.OracleModel(m, lanes=16)
    try:
        ours = np.stack([om.forward(int(toks[p]), p).copy() for p in range(n_tok)])
    finally:
        om.close()
    hf = _hf_model(pkg, m)
    with torch.no_grad():
        theirs = hf(torch.tensor(toks[None, :].astype(np.int64))).logits[0].numpy()
    scale = np.abs(theirs).max()
    err = np.abs(ours - theirs).max() / scale
    assert err <= tol, f"{shape} {quant}: oracle vs transformers max|d| / max|logit| = {err:.3e}"
    if quant == "F16":  # the greedy continuation is the same wherever the top-2 margin exceeds the rounding noise
        margin = np.sort(theirs, axis=1)
        clear = (margin[:, -1] - margin[:, -2]) > 10 * tol * scale
        assert np.array_equal(ours.argmax(axis=1)[clear], theirs.argmax(axis=1)[clear]) and clear.sum() >= n_tok // 2


def test_structural_mistakes_would_be_caught(pkg, orc, make_model):
    """Sensitivity of the check above: feeding transformers the GGUF-ordered (un-restored) Wq / Wk -- i.e. pairing RoPE the wrong way --
    moves the logits by far more than the tolerance."""
    m = make_model("tiny-llama", pkg.gguf.GGMLType.F16, 32)
    c = m.configuration
    toks = orc.bench_tokens(c.vocab_size, 12)
    hf = _hf_model(pkg, m)
    with torch.no_grad():
        good = hf(torch.tensor(toks[None, :].astype(np.int64))).logits[0].numpy()
        sd = hf.state_dict()
        for l in range(c.n_layers):
            k = f"model.layers.{l}.self_attn.q_proj.weight"
            sd[k] = torch.from_numpy(pkg.loader.tensor_as_f32(m, f"blk.{l}.attn_q.weight").reshape(c.n_heads * c.head_size, c.dim).astype(np.float64))
        hf.load_state_dict(sd)
        bad = hf(torch.tensor(toks[None, :].astype(np.int64))).logits[0].numpy()
    assert np.abs(good - bad).max() / np.abs(good).max() > 1e-2

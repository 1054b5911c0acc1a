"""K-quant load path and third-party pins of the data formats -- CPU.

The reference's accelerator path re-quantises Q4_K / Q5_K / Q6_K tensors to Q8_0 while loading (ModelLoader.java:163,173-224) and then
computes as for a Q8_0 file.  The oracle restates the element reads and the re-quantiser (oracle/oracle.c); here:
  * the element reads are PINNED against gguf-py (llama.cpp's own Python implementation of the formats, pinned in this image as
    gguf 0.19) -- bit-equal floats for all three formats, and the same for Q8_0 / F16 through the product's own dequantiser;
  * the re-quantiser is checked against a second, vectorised numpy restatement, and its rounding rule against hand-computed ties;
  * the product's GGUF reader and writer are pinned against gguf-py's GGUFReader / GGUFWriter;
  * the loader maps K-quant file types to a Q8_0 configuration (AbstractModelLoader.java:45-47).
The device side (csrc/kquant.cuh) is compared with the oracle in tests/test_gpu_kquants.py."""
import os

import numpy as np
import pytest

gguf_py = pytest.importorskip("gguf")


def _qt(pkg, tt):
    Q = gguf_py.GGMLQuantizationType
    G = pkg.gguf.GGMLType
    return {G.Q4_K: Q.Q4_K, G.Q5_K: Q.Q5_K, G.Q6_K: Q.Q6_K, G.Q8_0: Q.Q8_0, G.F16: Q.F16, G.F32: Q.F32}[tt]


@pytest.mark.parametrize("name", ["Q4_K", "Q5_K", "Q6_K"])
def test_kquant_reads_match_gguf_py_bit_exactly(pkg, orc, name):
    G = pkg.gguf.GGMLType
    tt = getattr(G, name)
    rng = np.random.Generator(np.random.PCG64(tt))
    n = 256 * 777
    raw = pkg.synth.random_kquant(tt, n, rng, zero_blocks=3)
    ours = orc.kquant_dequantize(tt, raw, n)
    theirs = gguf_py.quants.dequantize(raw.reshape(-1, G.SIZES[tt][0]), _qt(pkg, tt)).reshape(-1).astype(np.float32)
    assert np.array_equal(ours.view(np.uint32), theirs.view(np.uint32))
    assert np.abs(ours).max() > 0.05  # the synthetic blocks are not degenerate


def test_kquant_reads_with_arbitrary_bytes(pkg, orc):
    """Fully random bytes (every scale / min / high-bit pattern, any finite FP16 block scale) still agree with gguf-py."""
    G = pkg.gguf.GGMLType
    rng = np.random.Generator(np.random.PCG64(99))
    for tt in G.K_QUANTS:
        ts = G.SIZES[tt][0]
        raw = rng.integers(0, 256, size=(400, ts), dtype=np.uint8)
        for col in ((0, 2) if tt != G.Q6_K else (208,)):  # keep the FP16 scales finite (exponent 31 = inf/nan is not a weight)
            hi = raw[:, col + 1]
            raw[:, col + 1] = np.where((hi & 0x7C) == 0x7C, hi & 0xBF, hi)
        ours = orc.kquant_dequantize(tt, raw.reshape(-1), 400 * 256)
        theirs = gguf_py.quants.dequantize(raw, _qt(pkg, tt)).reshape(-1).astype(np.float32)
        assert np.array_equal(ours.view(np.uint32), theirs.view(np.uint32)), G.NAMES[tt]


@pytest.mark.parametrize("name", ["Q4_K", "Q5_K", "Q6_K"])
def test_requantiser_two_restatements_agree(pkg, orc, name):
    G = pkg.gguf.GGMLType
    tt = getattr(G, name)
    rng = np.random.Generator(np.random.PCG64(100 + tt))
    n = 256 * 300
    raw = pkg.synth.random_kquant(tt, n, rng, zero_blocks=2)
    q8 = orc.kquant_to_q8_0(tt, raw, n)
    assert np.array_equal(q8, orc.np_requant_q8_0(orc.kquant_dequantize(tt, raw, n)))
    blocks = q8.reshape(-1, 34)
    assert not blocks[:16].any()  # the two all-zero super-blocks: scale 0, quants 0 (the 1/scale guard)
    mx = np.abs(blocks[16:, 2:].view(np.int8).astype(np.int16)).max(axis=1)
    assert np.all((mx == 0) | (mx >= 126)) and (mx >= 126).mean() > 0.9  # a block is all zero or uses the full int8 range


def test_requantiser_rounding_is_java_math_round(orc):
    """Math.round(float) = floor(x + 1/2): ties go towards +infinity (-2.5 -> -2, 2.5 -> 3), unlike C's roundf / numpy's rint."""
    x = np.zeros(32, dtype=np.float32)
    x[0] = 127.0          # scale = 1, inv = 1
    x[1:7] = [2.5, -2.5, 0.5, -0.5, -126.5, 1.4999999]
    q = orc.np_requant_q8_0(x).reshape(-1, 34)[0, 2:].view(np.int8)
    assert list(q[:7]) == [127, 3, -2, 1, 0, -126, 1]
    assert orc.np_requant_q8_0(x)[:2].view(np.float16)[0] == np.float16(1.0)


def test_loader_maps_kquant_files_to_q8_0(pkg, tmp_path):
    """A K-quant GGUF written by the product's writer loads as a Q8_0 configuration with its K-quant tensors intact, and gguf-py's
    reader sees the same tensors at the same offsets."""
    G = pkg.gguf.GGMLType
    sh = pkg.synth.SHAPES["tiny-llama"]
    tensors = pkg.synth.build_tensors_kquant(sh, seed=3)
    md = pkg.synth.metadata_for(sh, G.Q8_0, "Llama synthetic tiny-llama")
    md["general.file_type"] = 15  # Q4_K_M
    order = [(name, tensors[name][0], dims, tensors[name][2]) for name, _, dims, _ in pkg.synth.tensor_plan(sh, G.Q8_0)]
    path = os.path.join(tmp_path, "tiny-q4_k_m.gguf")
    pkg.gguf.write_gguf(path, md, order)
    m = pkg.load_model(path, 32)
    assert m.configuration.quantization == "Q8_0"
    kinds = {m.tensors[n][0] for n in m.tensors}
    assert {G.Q4_K, G.Q5_K, G.Q6_K, G.F32} <= kinds
    for name, tt, dims, raw in order:
        assert m.tensors[name][0] == tt and np.array_equal(np.asarray(m.tensors[name][2]), raw), name
    rd = gguf_py.GGUFReader(path)
    seen = {t.name: t for t in rd.tensors}
    for name, tt, dims, raw in order:
        t = seen[name]
        assert int(t.tensor_type) == tt and tuple(int(d) for d in t.shape) == tuple(dims), name
        assert np.array_equal(np.asarray(t.data).reshape(-1).view(np.uint8), raw), name
    md["general.file_type"] = 10  # Q2_K: not supported by the reference either
    pkg.gguf.write_gguf(path, md, order)
    with pytest.raises(pkg.loader.UnsupportedModel):
        pkg.load_model(path, 32)


@pytest.mark.parametrize("quant", ["Q8_0", "F16"])
def test_reader_and_dequantiser_pinned_against_gguf_py(pkg, make_model, model_dir, quant):
    """The product's GGUF writer/reader and its getFloat restatement (loader.tensor_as_f32) against gguf-py on a synthetic model file:
    same metadata scalars, same tensor table, same bytes, bit-equal dequantised floats."""
    G = pkg.gguf.GGMLType
    tt = getattr(G, quant)
    m = make_model("tiny-qwen3", tt, 32)
    path = m.gguf.path if hasattr(m.gguf, "path") else os.path.join(model_dir, f"tiny-qwen3-{tt}-1234.gguf")
    rd = gguf_py.GGUFReader(path)
    assert int(rd.fields["general.file_type"].parts[-1][0]) == (7 if tt == G.Q8_0 else 1)
    assert int(rd.fields["qwen3.embedding_length"].parts[-1][0]) == m.configuration.dim
    assert len(rd.tensors) == len(m.tensors)
    for t in rd.tensors:
        ours_type, ours_dims, ours_raw = m.tensors[t.name]
        assert int(t.tensor_type) == ours_type and tuple(int(d) for d in t.shape) == tuple(ours_dims), t.name
        assert np.array_equal(np.asarray(t.data).reshape(-1).view(np.uint8), np.asarray(ours_raw)), t.name
        if ours_type in (G.Q8_0, G.F16):
            theirs = gguf_py.quants.dequantize(np.asarray(t.data), t.tensor_type).reshape(-1).astype(np.float32)
            ours = pkg.loader.tensor_as_f32(m, t.name)
            assert np.array_equal(ours.view(np.uint32), theirs.view(np.uint32)), t.name


def test_reader_reads_a_gguf_py_written_file(pkg, tmp_path):
    """The other direction: a file produced by gguf-py's GGUFWriter (llama.cpp's writer) parses with the product's reader."""
    G = pkg.gguf.GGMLType
    path = os.path.join(tmp_path, "third_party.gguf")
    rng = np.random.Generator(np.random.PCG64(5))
    w = gguf_py.GGUFWriter(path, "llama")
    w.add_name("Llama third-party writer")
    w.add_uint32("llama.block_count", 3)
    w.add_float32("llama.rope.freq_base", 500000.0)
    w.add_array("tokenizer.ggml.tokens", ["a", "bc", "def"])
    f32 = rng.standard_normal((4, 64)).astype(np.float32)
    f16 = rng.standard_normal((8, 32)).astype(np.float16)
    q8 = gguf_py.quants.quantize(rng.standard_normal((16, 64)).astype(np.float32), gguf_py.GGMLQuantizationType.Q8_0)
    w.add_tensor("a.f32", f32)
    w.add_tensor("b.f16", f16)
    w.add_tensor("c.q8", q8, raw_dtype=gguf_py.GGMLQuantizationType.Q8_0)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    g = pkg.gguf.GGUFFile(path)
    assert g.metadata["general.name"] == "Llama third-party writer"
    assert int(g.metadata["llama.block_count"]) == 3 and float(g.metadata["llama.rope.freq_base"]) == 500000.0
    assert list(g.metadata["tokenizer.ggml.tokens"]) == ["a", "bc", "def"]
    ti = g.tensor_infos
    assert ti["a.f32"].ggml_type == G.F32 and tuple(ti["a.f32"].dims) == (64, 4)
    assert ti["b.f16"].ggml_type == G.F16 and tuple(ti["b.f16"].dims) == (32, 8)
    assert ti["c.q8"].ggml_type == G.Q8_0 and tuple(ti["c.q8"].dims) == (64, 16)
    assert np.array_equal(np.asarray(g.tensor_bytes("a.f32")), f32.view(np.uint8).reshape(-1))
    assert np.array_equal(np.asarray(g.tensor_bytes("b.f16")), f16.view(np.uint8).reshape(-1))
    assert np.array_equal(np.asarray(g.tensor_bytes("c.q8")), np.asarray(q8).reshape(-1).view(np.uint8))

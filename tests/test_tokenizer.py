"""CPU tests of the native tokenizer (libb200tok.so through its C ABI) against the oracle restatement of the reference's
LlamaTokenizer / Qwen3Tokenizer / LlamaChatFormat (oracle/tokenizer_oracle.py).  Bar: token ids identical, bytes identical.
No real vocabularies exist offline: the vocabularies are synthetic but well-formed byte-level BPE vocabularies
(synth.build_vocab), plus an adversarial one whose merge ids are NOT in merge order."""
import ctypes
import importlib.util
import os
import random
import re
import time
import unicodedata

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tor():
    spec = importlib.util.spec_from_file_location("b200_tok_oracle", os.path.join(ROOT, "oracle", "tokenizer_oracle.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def vocabs(pkg):
    return {arch: pkg.synth.build_vocab(700, arch) for arch in ("llama", "qwen3")}


def make_pair(pkg, tor, vocabs, arch):
    tokens, merges, types, base = vocabs[arch]
    if arch == "llama":
        return pkg.tokenizer.LlamaTokenizer(tokens, merges, base_tokens=base), tor.OracleTokenizer(tokens, merges, "llama", base)
    return pkg.tokenizer.Qwen3Tokenizer(tokens, merges, types), tor.OracleTokenizer(tokens, merges, "qwen3", token_types=types)


WORDS = "the quick brown fox isn't I'll they've we're I'm he'd it's DON'T 'Tis 12345 3.14159 2024-09-24 $100% a1b2 x²+y³ ½ café naïve München ª º µ × ÷ 東京 こんにちは 🙂🚀 \t\n\r\n  !!! ?! ... -- 'quoted' \"double\" {json: [1,2,3]} <|eot_id|> end".split(" ")


def random_texts(n, seed):
    rng = random.Random(seed)
    out = ["", " ", "a", "'", "''s", "'s", "x's", "123", "1234567", "  leading", "trailing  ", "\n\n", "a b", " "]
    for _ in range(n):
        kind = rng.randrange(4)
        if kind == 0:
            out.append(" ".join(rng.choice(WORDS) for _ in range(rng.randrange(1, 30))))
        elif kind == 1:
            out.append("".join(rng.choice(WORDS) for _ in range(rng.randrange(1, 20))))
        elif kind == 2:
            out.append("".join(chr(rng.choice([rng.randrange(32, 127), rng.randrange(0xA0, 0x180), rng.randrange(0x400, 0x500), rng.randrange(0x4E00, 0x4F00),
                                                rng.randrange(0x1F600, 0x1F650), 9, 10, 13, 32, 39, 39])) for _ in range(rng.randrange(1, 80))))
        else:
            out.append(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 60))).decode("utf-8", errors="ignore"))
    return out


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "b200tok.h")).read()
    declared = set(re.findall(r"\b(b200_tok_\w+)\s*\(", hdr))
    assert declared == set(pkg.tokenizer.EXPORTS)
    L = pkg.tokenizer.lib()
    for name in declared:
        assert isinstance(getattr(L, name), ctypes._CFuncPtr)


def test_byte_table_properties(tor):
    enc = tor.BYTE_ENCODER
    assert sorted(enc) == list(range(256)) and len(set(enc.values())) == 256
    assert all(enc[b] == b for b in range(0x21, 0x7F)) and all(enc[b] == b for b in list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100)))
    assert enc[0x20] == 0x120 and enc[0x0A] == 0x10A and enc[0x00] == 0x100  # the famous 'Ġ' and 'Ċ'
    # premise of the native matcher: no whitespace / line break survives the mapping, and every symbol is L, N or neither
    assert not any(chr(c).isspace() or chr(c) in "\r\n" for c in enc.values())


@pytest.mark.parametrize("kind", [0, 1])
def test_pretokenisation_matches_the_regex_on_every_symbol_pair(pkg, tor, kind):
    """The hand-written matcher == java.util.regex semantics of the reference pattern on byte-mapped text: all 65536
    ordered byte pairs between letters (classification + the optional-prefix rule), then random byte strings."""
    import regex
    pat = regex.compile(tor.LLAMA_3_PATTERN if kind == 0 else tor.QWEN3_PATTERN)
    enc = tor.BYTE_ENCODER

    def expect(data):
        mapped = "".join(chr(enc[b]) for b in data)
        return [len(m) for m in pat.findall(mapped)]

    for a in range(256):
        data = b"".join(bytes([ord("x"), a, b, ord("y")]) for b in range(256))
        assert pkg.tokenizer.split_lengths(data, kind) == expect(data), f"first byte {a}"
    rng = random.Random(5)
    alphabet = bytes(range(256)) + b"'''sStTrReEvVmMlLdD0123456789  \n"
    for _ in range(2000):
        data = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
        assert pkg.tokenizer.split_lengths(data, kind) == expect(data), data
    # the class table itself, against Unicode
    for b in range(256):
        cat = unicodedata.category(chr(enc[b]))[0]
        got = pkg.tokenizer.split_lengths(bytes([b, b, b, b]), kind)
        want = [4] if cat == "L" else ([3, 1] if kind == 0 else [1, 1, 1, 1]) if cat == "N" else [4]
        assert got == want, (b, cat)


@pytest.mark.parametrize("arch", ["llama", "qwen3"])
def test_encode_decode_parity_with_oracle(pkg, tor, vocabs, arch):
    nat, orc = make_pair(pkg, tor, vocabs, arch)
    for text in random_texts(1500, 11):
        ids = nat.encode_as_list(text)
        assert ids == orc.encode(text), repr(text)
        assert nat.decode(ids) == text == orc.decode(ids)  # byte-level BPE is lossless on valid UTF-8
        assert nat.decode_bytes(ids) == text.encode("utf-8")
    assert nat.encode_as_list("") == []
    nat.close()


@pytest.mark.parametrize("arch", ["llama", "qwen3"])
def test_committed_tokenizer_fixture(pkg, tor, vocabs, arch):
    """Regression pin (tests/golden/tokenizer_golden.json, self-generated by make_tokenizer_golden.py): the synthetic
    vocabulary, the oracle and the native encoder all still produce the committed ids."""
    import hashlib
    import json
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json")) as f:
        gold = json.load(f)[arch]
    tokens, merges, _, _ = vocabs[arch]
    assert hashlib.sha256("\n".join(tokens + merges).encode()).hexdigest() == gold["vocab_sha"]
    nat, orc = make_pair(pkg, tor, vocabs, arch)
    for text, ids in gold["cases"]:
        assert nat.encode_as_list(text) == ids == orc.encode(text), repr(text)


def test_merge_priority_is_the_merged_token_id_not_the_merge_order(pkg, tor):
    """Adversarial vocabulary: the merge list says (a,b) first, but the merged token of (b,c) has the lower id, and a
    later merge consumes the product of an earlier one with a lower id -- the reference's loop semantics decide."""
    sym = pkg.synth.gpt2_byte_symbols()
    tokens = list(sym) + ["bc", "ab", "abc", "aba", "abab"] + ["<|begin_of_text|>"]
    merges = ["a b", "b c", "ab c", "a bc", "ab a", "ab ab"]
    nat = pkg.tokenizer.LlamaTokenizer(tokens, merges, base_tokens=len(tokens) - 1)
    orc = tor.OracleTokenizer(tokens, merges, "llama", len(tokens) - 1)
    for text in ["abc", "abab", "ababab", "abcabc", "aabcb", "ababa", "bcab", "abababc" * 3]:
        assert nat.encode_as_list(text) == orc.encode(text), text
    assert nat.encode_as_list("abc") == [tokens.index("abc")]
    two = list(sym) + ["bc", "ab", "<|begin_of_text|>"]
    nat2 = pkg.tokenizer.LlamaTokenizer(two, ["a b", "b c"], base_tokens=len(two) - 1)
    assert nat2.encode_as_list("abc") == [ord("a"), two.index("bc")]  # 'bc' (id 256) beats 'ab' (id 257) although "a b" is listed first


def test_special_tokens_and_chat_format(pkg, tor, vocabs):
    nat, orc = make_pair(pkg, tor, vocabs, "llama")
    assert nat.get_special_tokens() == orc.special_tokens and len(nat.get_special_tokens()) == len(pkg.synth.LLAMA_SPECIALS)
    eot = nat.get_special_tokens()["<|eot_id|>"]
    assert nat.is_special_token(eot) and not nat.should_display_token(eot) and nat.should_display_token(65)
    mapped = tor.map_bytes("hi <|eot_id|> there<|eot_id|>")
    assert nat.encode_with_special(mapped, {"<|eot_id|>"}) == orc.encode_with_special(mapped, {"<|eot_id|>"})
    # Java's String.split drops the delimiters: the reference's encode(text, allowedSpecial) loses the special tokens it splits at
    # (LlamaTokenizer.java:164-176); the text between them is encoded chunk by chunk
    got = nat.encode_with_special(mapped, {"<|eot_id|>"})
    assert eot not in got and got == nat.encode_with_special(tor.map_bytes("hi "), set()) + nat.encode_with_special(tor.map_bytes(" there"), set())
    with pytest.raises(pkg.tokenizer.TokenizerError):
        nat.encode_with_special(mapped, {"<|nope|>"})
    fmt, ofmt = pkg.chat_format.LlamaChatFormat(nat), tor.OracleLlamaChatFormat(orc)
    M = pkg.chat_format.Message
    dialog = [("system", "  You are terse. \n"), ("user", "What's 2+2?\n"), ("assistant", "4"), ("user", " and 3+3? ")]
    got = fmt.encode_dialog_prompt(True, [M(r, c) for r, c in dialog])
    assert got == ofmt.encode_dialog_prompt(True, dialog)
    assert got[0] == fmt.get_begin_of_text() and got.count(fmt.end_of_turn) == 4 and fmt.get_stop_tokens() == {fmt.end_of_text, fmt.end_of_turn}
    assert nat.decode([t for t in got if not nat.is_special_token(t)]).startswith("system\nYou are terse.user\nWhat's 2+2?")
    q, oq = make_pair(pkg, tor, vocabs, "qwen3")
    assert "<think>" not in q.get_special_tokens() and q.think_start_token == q.tokens.index("<think>")
    assert q.get_special_tokens() == oq.special_tokens
    assert q.should_display_token(q.think_start_token) and not q.should_display_token(q.get_special_tokens()["<|im_end|>"])


def test_qwen3_chat_format(pkg, tor, vocabs):
    q, oq = make_pair(pkg, tor, vocabs, "qwen3")
    fmt = pkg.chat_format.Qwen3ChatFormat(q)
    ofmt = tor.OracleQwen3ChatFormat(oq, q.tokens.index("<think>"), q.tokens.index("</think>"))
    M = pkg.chat_format.Message
    for role, content in [("system", " You are helpful.\n"), ("user", "What's 7*6?"), ("assistant", "42")]:
        assert fmt.encode_message(M(role, content)) == ofmt.encode_message(role, content)
    assert fmt.encode_header(M("assistant", "")) == ofmt.encode_header("assistant")
    assert fmt.get_begin_of_text() == q.get_special_tokens()["<|im_start|>"]
    assert fmt.get_stop_tokens() == ofmt.stop_tokens() == {q.get_special_tokens()["<|im_end|>"], q.get_special_tokens()["<|endoftext|>"]}
    assert fmt.encode_thinking_control(True) == [] and fmt.encode_thinking_control(False) == ofmt.thinking_control(False)
    assert fmt.encode_thinking_control(False)[0] == q.think_start_token and q.decode(fmt.encode_thinking_control(False)) == "<think>\n\n</think>\n\n"


def test_error_conventions(pkg):
    sym = pkg.synth.gpt2_byte_symbols()
    with pytest.raises(pkg.tokenizer.TokenizerError, match="missing from the vocabulary"):
        pkg.tokenizer.LlamaTokenizer(list(sym), ["a b"], base_tokens=256)  # merged token 'ab' absent: orElseThrow()
    t = pkg.tokenizer.LlamaTokenizer([s for s in sym if s != "z"] + ["<|x|>"], [], base_tokens=255)
    assert t.encode_as_list("abc") == [t.index("a"), t.index("b"), t.index("c")]
    with pytest.raises(pkg.tokenizer.TokenizerError):
        t.encode_as_list("zebra")  # vocabulary.getIndex("z").orElseThrow()


def test_tokenizer_metadata_round_trips_through_gguf(pkg, tmp_path):
    path = str(tmp_path / "tiny.gguf")
    pkg.synth.write_model(path, "tiny-llama", pkg.gguf.GGMLType.Q8_0)
    model = pkg.load_model(path, 16)
    tok = pkg.tokenizer.from_metadata(model.gguf.metadata, model.model_type)
    assert len(tok.tokens) == model.configuration.vocab_size
    ids = tok.encode_as_list("the quick brown fox")
    assert tok.decode(ids) == "the quick brown fox" and max(ids) < model.configuration.vocab_size
    assert pkg.chat_format.LlamaChatFormat(tok).encode_dialog_prompt(True, [pkg.chat_format.Message("user", "hi")])[0] == tok.index("<|begin_of_text|>")


def test_native_is_faster_than_the_restatement(pkg, tor, vocabs):
    """Measurement beside parity: tokens/s of the native encoder vs the Python restatement of the reference algorithm
    (the reference's own Java loop rebuilds a HashMap and an ArrayList per merge step, as the restatement does)."""
    nat, orc = make_pair(pkg, tor, vocabs, "llama")
    text = " ".join(random.Random(3).choice(WORDS) for _ in range(4000))
    t0 = time.perf_counter(); a = nat.encode_as_list(text); t1 = time.perf_counter(); b = orc.encode(text); t2 = time.perf_counter()
    assert a == b
    print(f"\nnative {len(a) / (t1 - t0):,.0f} tok/s, restatement {len(b) / (t2 - t1):,.0f} tok/s on {len(text)} chars")
    assert (t1 - t0) < (t2 - t1)


def test_encode_agrees_with_huggingface_tokenizers(pkg, tor, vocabs):
    """THIRD-PARTY pin (SURVEY 8(f) N2 was checked only against this repository's own restatement): Hugging Face `tokenizers` (the Rust
    BPE) configured with the reference's pipeline ORDER -- byte-level symbol mapping first, THEN the pre-tokenisation regex on the mapped
    text (LlamaTokenizer.java:164-200 does exactly this, which is why words are not split at spaces there), then BPE over the chunk --
    produces the same ids as the native tokenizer and as the oracle on the synthetic vocabularies (whose token ids are in merge order, so
    'lowest merged-token id first' and 'lowest merge rank first' coincide).  Independent implementations of the regex (Oniguruma vs the
    native matcher vs Python `regex`) and of the merge loop."""
    hf = pytest.importorskip("tokenizers")
    for arch in ("llama", "qwen3"):
        tokens, merges, types, base = vocabs[arch]
        ours, oracle = make_pair(pkg, tor, vocabs, arch)
        pattern = tor.LLAMA_3_PATTERN if arch == "llama" else tor.QWEN3_PATTERN
        tk = hf.Tokenizer(hf.models.BPE(vocab={t: i for i, t in enumerate(tokens)}, merges=[tuple(m.split(" ")) for m in merges]))
        tk.pre_tokenizer = hf.pre_tokenizers.Sequence([hf.pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False),
                                                       hf.pre_tokenizers.Split(hf.Regex(pattern), behavior="isolated")])
        n_tok = 0
        for text in random_texts(300, 77):
            want = tk.encode(text, add_special_tokens=False).ids
            assert ours.encode(text) == want, (arch, text)
            assert oracle.encode(text) == want, (arch, text)
            n_tok += len(want)
        assert n_tok > 5000

"""Regenerates tests/golden/tokenizer_golden.json from the tokenizer ORACLE (python tests/golden/make_tokenizer_golden.py).
Self-generated: the reference holds no tokenizer vectors and no real vocabulary is available offline; the fixture pins the
restatement (and the synthetic vocabulary builder) against drift."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.import_package()
spec = importlib.util.spec_from_file_location("tok_oracle", os.path.join(ROOT, "oracle", "tokenizer_oracle.py"))
tor = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tor)

TEXTS = ["", "hello world", "the quick brown fox isn't 12345 done!!", "I'll they've we're I'M DON'T", "  double  spaces\tand\nnewlines\r\n",
         "3.14159 2024-09-24 $100% x²+y³ ½", "Grüße aus München, naïve café", "東京 こんにちは 🙂🚀", "'quoted' \"double\" {json: [1,2,3]}", "a" * 40 + " " + "ab" * 20]
out = {}
for arch in ("llama", "qwen3"):
    tokens, merges, types, base = pkg.synth.build_vocab(700, arch)
    o = tor.OracleTokenizer(tokens, merges, arch, base, token_types=types)
    out[arch] = {"vocab_sha": __import__("hashlib").sha256("\n".join(tokens + merges).encode()).hexdigest(), "cases": [[t, o.encode(t)] for t in TEXTS]}
with open(os.path.join(os.path.dirname(__file__), "tokenizer_golden.json"), "w") as f:
    json.dump(out, f, indent=1, ensure_ascii=False)
print({k: len(v["cases"]) for k, v in out.items()})

"""tokenizer_oracle.py -- TEST INFRASTRUCTURE: line-by-line Python restatement of the reference's byte-level BPE
tokenizers (tokenizer/LlamaTokenizer.java:30-269, tokenizer/Qwen3Tokenizer.java:20-352) and of LlamaChatFormat's
prompt assembly (model/format/LlamaChatFormat.java:24-77).  Only tests/ may import it; the product is
gpullama3.java_b200/csrc/tokenizer.cpp behind include/b200tok.h.

PARITY UNPINNED: the reference holds no tokenizer vectors and cannot run here (no JDK); there are no real
vocabularies offline either, so the tests use synthetic byte-level vocabularies (tests build them with a tiny BPE
trainer).  What pins this restatement: the GPT-2 byte<->unicode table is checked against its published properties
(bijection on 0..255, printable ASCII / Latin-1 fixed points), and decode(encode(s)) == s on arbitrary UTF-8.

Faithfully kept quirks of the reference:
  * the pre-tokenisation regex runs on the BYTE-MAPPED text (LlamaTokenizer.java:246-252 maps first, then
    encodeImpl -> encodeOrdinary -> findAll), where space/newline/tab are the letters U+0120/U+010A/U+0109: the
    whitespace alternatives of the pattern can never match and words are NOT split at spaces;
  * the merge priority is the MERGED TOKEN'S ID, not the position in the merges list (`this.merges.put(pair,
    mergeIndex)` :57-62, `min(comparingInt(merges.getOrDefault(key, MAX)))` :213);
  * every occurrence of the chosen pair is merged in one left-to-right pass (:77-91) before the next pair is chosen.
One documented deviation: when two DIFFERENT pairs present in a chunk produce the same merged id (impossible for a
consistent BPE vocabulary) Java's choice depends on HashMap iteration order; here (and in the native code) the pair
whose first occurrence is leftmost wins.
"""
from __future__ import annotations

import regex

LLAMA_3_PATTERN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
QWEN3_PATTERN = r"(?:'[sS]|'[tT]|'[rR][eE]|'[vV][eE]|'[mM]|'[lL][lL]|'[dD])|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
INT_MAX = 2 ** 31 - 1


def bytes_to_unicode() -> dict[int, int]:
    """LlamaTokenizer.bytesToUnicode (:98-116)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = list(bs)
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, cs))


BYTE_ENCODER = bytes_to_unicode()
BYTE_DECODER = {v: k for k, v in BYTE_ENCODER.items()}


def map_bytes(text: str) -> str:
    return "".join(chr(BYTE_ENCODER[b]) for b in text.encode("utf-8"))


class OracleTokenizer:
    def __init__(self, tokens: list[str], merge_lines: list[str], kind: str = "llama", base_tokens: int = 128000,
                 token_types: list[int] | None = None):
        self.tokens = tokens
        self.index = {t: i for i, t in enumerate(tokens)}
        self.pattern = regex.compile(LLAMA_3_PATTERN if kind == "llama" else QWEN3_PATTERN)
        self.kind = kind
        self.token_types = token_types
        if kind == "qwen3":  # Qwen3Tokenizer.java:58-60: everything from <|endoftext|> on is special
            base_tokens = self.index["<|endoftext|>"]
        self.special_tokens = {tokens[i]: i for i in range(base_tokens, len(tokens))}
        if kind == "qwen3":  # :76-79
            self.special_tokens.pop("<think>", None)
            self.special_tokens.pop("</think>", None)
        self.merges = {}
        for line in merge_lines:
            a, b = line.split(" ")
            ia, ib = self.index[a], self.index[b]
            self.merges[(ia, ib)] = self.index[tokens[ia] + tokens[ib]]

    # -- encode ---------------------------------------------------------------------------------------------
    def encode_chunk(self, chunk: str) -> list[int]:
        ids = [self.index[c] for c in chunk]
        while len(ids) >= 2:
            stats = {}
            for i in range(len(ids) - 1):
                stats[(ids[i], ids[i + 1])] = stats.get((ids[i], ids[i + 1]), 0) + 1
            pair = min(stats, key=lambda k: self.merges.get(k, INT_MAX))  # first minimum in first-occurrence order
            if pair not in self.merges:
                break
            idx = self.merges[pair]
            new, i = [], 0
            while i < len(ids):
                if ids[i] == pair[0] and i < len(ids) - 1 and ids[i + 1] == pair[1]:
                    new.append(idx)
                    i += 2
                else:
                    new.append(ids[i])
                    i += 1
            ids = new
        return ids

    def encode_ordinary(self, mapped_text: str) -> list[int]:
        out = []
        for chunk in self.pattern.findall(mapped_text):
            out.extend(self.encode_chunk(chunk))
        return out

    def encode(self, text: str) -> list[int]:
        """Tokenizer.encodeAsList / encode(String): UTF-8 bytes -> byte-level unicode -> regex split -> BPE."""
        return self.encode_ordinary(map_bytes(text))

    def encode_with_special(self, mapped_text: str, allowed_special: set[str]) -> list[int]:
        """encode(String, Set) (:140-171): split at exact occurrences of the allowed special tokens."""
        if not allowed_special:
            return self.encode_ordinary(mapped_text)
        assert all(s in self.special_tokens for s in allowed_special)
        # String.split(regex) never returns the delimiters, even for a capturing group: the specials are dropped (:164)
        pat = "(?:" + "|".join(regex.escape(s) for s in allowed_special) + ")"
        ids = []
        parts = regex.split(pat, mapped_text)
        while parts and parts[-1] == "":  # String.split drops trailing empty strings
            parts.pop()
        for part in parts:
            if part in allowed_special:
                ids.append(self.special_tokens[part])
            else:
                ids.extend(self.encode_ordinary(part))
        return ids

    # -- decode ---------------------------------------------------------------------------------------------
    def decode(self, ids: list[int]) -> str:
        s = "".join(self.tokens[i] for i in ids)
        if self.kind == "qwen3":  # :311-312: code points above 512 pass through (DeepSeek's full-width bars)
            raw = bytes((BYTE_DECODER[ord(c)] if ord(c) <= 512 else ord(c)) & 0xFF for c in s)
        else:
            raw = bytes(BYTE_DECODER[ord(c)] for c in s)
        return raw.decode("utf-8", errors="replace")

    def is_special_token(self, i: int) -> bool:
        return i in self.special_tokens.values()

    def should_display_token(self, i: int) -> bool:
        if self.kind == "qwen3":  # :174-178
            return self.token_types[i] in (1, 4, 6)
        return not self.is_special_token(i)


class OracleLlamaChatFormat:
    """LlamaChatFormat.java:24-77 (no tool calling)."""

    def __init__(self, tok: OracleTokenizer):
        st = tok.special_tokens
        self.tok = tok
        self.begin_of_text = st["<|begin_of_text|>"]
        self.start_header = st["<|start_header_id|>"]
        self.end_header = st["<|end_header_id|>"]
        self.end_of_turn = st["<|eot_id|>"]
        self.end_of_text = st["<|end_of_text|>"]
        self.stop_tokens = {self.end_of_text, self.end_of_turn}

    def encode_header(self, role: str, content: str = "") -> list[int]:
        return [self.start_header] + self.tok.encode(role) + [self.end_header] + self.tok.encode("\n")

    def encode_message(self, role: str, content: str) -> list[int]:
        return self.encode_header(role) + self.tok.encode(content.strip()) + [self.end_of_turn]

    def encode_dialog_prompt(self, append_assistant_turn: bool, dialog: list[tuple[str, str]]) -> list[int]:
        out = [self.begin_of_text]
        for role, content in dialog:
            out += self.encode_message(role, content)
        if append_assistant_turn:
            out += self.encode_header("assistant")  # ChatFormat.Role.ASSISTANT = new Role("assistant") (ChatFormat.java:246)
        return out


class OracleQwen3ChatFormat:
    """Qwen3ChatFormat.java:26-183, ChatML branch (chat tokens of Qwen3ModelLoader.java:89)."""

    def __init__(self, tok: OracleTokenizer, think_start: int, think_end: int):
        st = tok.special_tokens
        self.tok = tok
        self.im_start = st.get("<|im_start|>", -1)
        self.im_end = st.get("<|im_end|>", -1)
        self.end_of_text = st.get("<|end_of_text|>", -1)
        self.end_of_text_fim = st.get("<|endoftext|>", -1)
        self.think_start, self.think_end = think_start, think_end

    def encode_header(self, role: str) -> list[int]:
        return [self.im_start] + self.tok.encode(role) + self.tok.encode("\n")

    def encode_message(self, role: str, content: str) -> list[int]:
        return self.encode_header(role) + self.tok.encode(content.strip()) + [self.im_end] + self.tok.encode("\n")

    def stop_tokens(self) -> set[int]:
        return {t for t in (self.im_end, self.end_of_text, self.end_of_text_fim) if t != -1}

    def thinking_control(self, enable: bool) -> list[int]:
        if enable:
            return []
        return [self.think_start] + self.tok.encode("\n\n") + [self.think_end] + self.tok.encode("\n\n")

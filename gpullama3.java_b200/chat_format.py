"""Prompt assembly mirrors: ``LlamaChatFormat`` <- model/format/LlamaChatFormat.java:24-77 and the plain
(non-tool, thinking left to the template) part of ``Qwen3ChatFormat`` <- model/format/Qwen3ChatFormat.java:60-100.
Host string code; the token work goes through tokenizer.py (native)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Message:
    role: str      # ChatFormat.Role: "system" | "user" | "assistant" (ChatFormat.java:243-249)
    content: str


class LlamaChatFormat:
    def __init__(self, tokenizer):
        st = tokenizer.get_special_tokens()
        self.tokenizer = tokenizer
        self.begin_of_text = st["<|begin_of_text|>"]
        self.start_header = st["<|start_header_id|>"]
        self.end_header = st["<|end_header_id|>"]
        self.end_of_turn = st["<|eot_id|>"]
        self.end_of_text = st["<|end_of_text|>"]
        self.end_of_message = st.get("<|eom_id|>", -1)   # only in 3.1
        self.python_tag = st.get("<|python_tag|>", -1)   # only in 3.1
        self.stop_tokens = {self.end_of_text, self.end_of_turn}

    def get_begin_of_text(self) -> int:
        return self.begin_of_text

    def get_stop_tokens(self) -> set[int]:
        return self.stop_tokens

    def encode_header(self, message: Message) -> list[int]:
        t = self.tokenizer
        return [self.start_header] + t.encode_as_list(message.role) + [self.end_header] + t.encode_as_list("\n")

    def encode_message(self, message: Message) -> list[int]:
        return self.encode_header(message) + self.tokenizer.encode_as_list(message.content.strip()) + [self.end_of_turn]

    def encode_dialog_prompt(self, append_assistant_turn: bool, dialog: list[Message]) -> list[int]:
        tokens = [self.begin_of_text]
        for m in dialog:
            tokens += self.encode_message(m)
        if append_assistant_turn:
            tokens += self.encode_header(Message("assistant", ""))
        return tokens

    default_temperature = 0.3
    default_top_p = 0.95

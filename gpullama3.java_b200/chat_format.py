"""Prompt assembly mirrors: ``LlamaChatFormat`` <- model/format/LlamaChatFormat.java:24-77 and ``Qwen3ChatFormat`` <-
model/format/Qwen3ChatFormat.java:26-183 (ChatML header / message / stop tokens / thinking control; the DeepSeek-R1 and
tool-calling branches are not mirrored).  Host string code; the token work goes through tokenizer.py (native)."""
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


@dataclass(frozen=True)
class ChatTokens:
    """ChatFormat.ChatTokens (ChatFormat.java:214); Qwen3ModelLoader.java:88-89 builds the Qwen3 instance below."""
    t_start_header: str = "<|im_start|>"
    t_end_header: str = "<|im_end|>"
    t_end_of_turn: str = ""
    t_end_of_text: str = "<|end_of_text|>"
    t_end_of_text_fim: str = "<|endoftext|>"


class Qwen3ChatFormat:
    def __init__(self, tokenizer, chat_tokens: ChatTokens = ChatTokens()):
        st = tokenizer.get_special_tokens()
        self.tokenizer = tokenizer
        self.chat_tokens = chat_tokens
        self.begin_of_text = -1  # Qwen3 has no BOS token
        self.start_header = st.get(chat_tokens.t_start_header, -1)
        self.end_header = st.get(chat_tokens.t_end_header, -1)
        self.end_of_turn = st.get(chat_tokens.t_end_of_turn, -1)
        self.end_of_text = st.get(chat_tokens.t_end_of_text, -1)
        self.end_of_text_fim = st.get(chat_tokens.t_end_of_text_fim, -1)
        self.im_start, self.im_end = self.start_header, self.end_header
        self.fim = {r: st.get(f"<|{r}|>", -1) for r in ("fim_prefix", "fim_suffix", "fim_middle")}
        if self.end_header == -1:
            raise NotImplementedError("DeepSeek-R1 distill header tokens are not mirrored")

    def encode_header(self, message: Message) -> list[int]:
        if message.role in self.fim:
            return [self.fim[message.role]]
        t = self.tokenizer
        return [self.im_start] + t.encode_as_list(message.role) + t.encode_as_list("\n")  # encodeOrdinaryAsList == encodeAsList here

    def encode_message(self, message: Message) -> list[int]:
        tokens = self.encode_header(message) + self.tokenizer.encode_as_list(message.content.strip())
        if self.im_end != -1 and message.role not in self.fim:
            tokens += [self.im_end] + self.tokenizer.encode_as_list("\n")  # ChatML: a newline follows <|im_end|>
        return tokens

    def get_begin_of_text(self) -> int:
        return self.start_header if self.begin_of_text == -1 else self.begin_of_text

    def get_stop_tokens(self) -> set[int]:
        if self.im_end == -1 and self.end_of_text == -1:
            raise RuntimeError("No stop token is defined.")
        return {t for t in (self.im_end, self.end_of_text, self.end_of_text_fim) if t != -1}

    def supports_thinking(self) -> bool:
        return self.im_end != -1

    def encode_thinking_control(self, enable_thinking: bool) -> list[int]:
        """Qwen3ChatFormat.encodeThinkingControl (:168-183): a pre-closed <think> block when thinking is disabled."""
        if enable_thinking or not self.supports_thinking():
            return []
        t = self.tokenizer
        if t.think_start_token == -1 or t.think_end_token == -1:
            return t.encode_as_list("<think>\n\n</think>\n\n")
        return [t.think_start_token] + t.encode_as_list("\n\n") + [t.think_end_token] + t.encode_as_list("\n\n")

    default_temperature = 0.8
    default_top_p = 0.9

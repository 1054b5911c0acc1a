/* b200tok.h -- C ABI of the native byte-level BPE tokenizer (gpullama3.java_b200/csrc/tokenizer.cpp -> libb200tok.so).
 *
 * the host-side tokenizer of the reference, natively.  Replaces, for the Llama-3 and Qwen3 GGUF
 * vocabularies, tokenizer/LlamaTokenizer.java:30-269 and tokenizer/Qwen3Tokenizer.java:20-352 (encode / encodeOrdinary /
 * encodeChunk / decode), keeping their observable behaviour token for token -- including that the pre-tokenisation
 * pattern is applied to the byte-mapped text and that merge priority is the merged token's id (see
 * oracle/tokenizer_oracle.py for the list of quirks).  CPU only; no CUDA, no torch types. */
#ifndef B200TOK_H
#define B200TOK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_tok b200_tok;

#define B200_TOK_LLAMA3 0 /* LLAMA_3_PATTERN (LlamaTokenizer.java:33): digits in groups of up to 3 */
#define B200_TOK_QWEN3 1  /* QWEN3_PATTERN  (Qwen3Tokenizer.java:23): one digit per chunk */

#define B200_TOK_OK 0
#define B200_TOK_ERR_BAD_ARG (-1)
#define B200_TOK_ERR_VOCAB (-2)   /* a byte token, merge part or merged token is missing from the vocabulary */
#define B200_TOK_ERR_SPACE (-3)   /* output buffer too small: *n_out holds the required size */

/* LlamaTokenizer(metadata, vocabulary): `tokens` = tokenizer.ggml.tokens (UTF-8 strings of byte-mapped text),
 * `merges` = tokenizer.ggml.merges ("left right" lines).  Strings are copied. */
int b200_tok_create(const char *const *tokens, int32_t n_tokens, const char *const *merges, int32_t n_merges, int32_t kind,
                    b200_tok **out, char *err, size_t err_len);

/* Tokenizer.encodeAsList(text) / encode(String): UTF-8 in, token ids out (no special-token handling, exactly
 * like the reference's encode(String)).  Returns B200_TOK_ERR_SPACE with *n_out = needed when cap is too small. */
int b200_tok_encode(const b200_tok *tok, const char *utf8, size_t len, int32_t *ids, size_t cap, size_t *n_out);

/* encodeOrdinary(String) on text that is ALREADY byte-mapped (what encode(String, Set) passes down). */
int b200_tok_encode_mapped(const b200_tok *tok, const char *mapped_utf8, size_t len, int32_t *ids, size_t cap, size_t *n_out);

/* Tokenizer.decode(tokens): concatenated token strings mapped back to bytes (raw bytes out, not validated as UTF-8;
 * Qwen3: code points above 512 are truncated to a byte as Qwen3Tokenizer.java:311-317 does). */
int b200_tok_decode(const b200_tok *tok, const int32_t *ids, size_t n, char *out, size_t cap, size_t *n_out);

/* Test hook: the pre-tokenisation alone -- byte lengths of the chunks java.util.regex's findAll would return for the
 * byte-mapped text (LlamaTokenizer.findAll, :66-73). */
int b200_tok_split(int32_t kind, const char *utf8, size_t len, int32_t *chunk_lens, size_t cap, size_t *n_out);

/* vocabulary.getIndex(token): id of an exact token string, or -1. */
int32_t b200_tok_index(const b200_tok *tok, const char *token_utf8);

int32_t b200_tok_vocab_size(const b200_tok *tok);
void b200_tok_free(b200_tok *tok);

#ifdef __cplusplus
}
#endif
#endif

// tokenizer.cpp -- native byte-level BPE tokenizer behind include/b200tok.h (libb200tok.so, CPU only).
//
// What the reference does per call (LlamaTokenizer.java:173-252, Qwen3Tokenizer.java:190-262) and how it is done here:
//   1. UTF-8 bytes -> GPT-2 "byte to unicode" code points                         -> table lookup, the text stays an int array
//   2. java.util.regex findAll with LLAMA_3_PATTERN / QWEN3_PATTERN on THAT text   -> a hand-written matcher: over the 256-symbol
//      mapped alphabet there is no whitespace, '\r' or '\n' (they are the letters U+0120, U+010D, U+010A), so of the seven
//      alternatives only four can ever match: contraction | [other]?letters+ | digits{1,3} (Qwen3: one digit) | others+
//   3. per chunk: repeatedly pick the adjacent pair whose MERGED token id is lowest and merge all its occurrences left
//      to right (HashMap + stream.min + ArrayList rebuild in the reference)         -> the same loop on a flat int array
// Behaviour is checked token for token against oracle/tokenizer_oracle.py (tests/test_tokenizer.py).
#include "../../include/b200tok.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

enum Cls : uint8_t { OTHER = 0, LETTER = 1, NUMBER = 2 };

struct ByteTable {
    int cp_of_byte[256];   // BYTE_ENCODER (LlamaTokenizer.bytesToUnicode, :98-116)
    int byte_of_cp[0x144]; // BYTE_DECODER for code points < 0x144, -1 where undefined
    Cls cls_of_byte[256];  // Unicode general category of the MAPPED code point: L* -> LETTER, N* -> NUMBER
    ByteTable() {
        bool keep[256] = {false};
        for (int b = '!'; b <= '~'; b++) keep[b] = true;
        for (int b = 0xA1; b <= 0xAC; b++) keep[b] = true;
        for (int b = 0xAE; b <= 0xFF; b++) keep[b] = true;
        for (int i = 0; i < 0x144; i++) byte_of_cp[i] = -1;
        int n = 0;
        for (int b = 0; b < 256; b++) {
            cp_of_byte[b] = keep[b] ? b : 256 + n++;
            byte_of_cp[cp_of_byte[b]] = b;
        }
        for (int b = 0; b < 256; b++) {
            const int cp = cp_of_byte[b];
            Cls c = OTHER;
            if (cp >= 0x100) c = LETTER; // U+0100..U+0143: Latin Extended-A, all letters
            else if ((cp >= 'A' && cp <= 'Z') || (cp >= 'a' && cp <= 'z')) c = LETTER;
            else if (cp >= '0' && cp <= '9') c = NUMBER;
            else if (cp == 0xAA || cp == 0xB5 || cp == 0xBA) c = LETTER;                 // feminine/masculine ordinal, micro sign
            else if (cp == 0xB2 || cp == 0xB3 || cp == 0xB9 || (cp >= 0xBC && cp <= 0xBE)) c = NUMBER; // superscripts, fractions (No)
            else if (cp >= 0xC0 && cp != 0xD7 && cp != 0xF7) c = LETTER;                 // Latin-1 letters except the two operators
            cls_of_byte[b] = c;
        }
    }
};
const ByteTable &table() {
    static const ByteTable t;
    return t;
}

// UTF-8 -> code points (lenient: invalid sequences become one code point per byte)
void utf8_to_cps(const char *s, size_t len, std::vector<int> &out) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(s);
    size_t i = 0;
    while (i < len) {
        unsigned c = p[i];
        int extra = c < 0x80 ? 0 : (c >> 5) == 6 ? 1 : (c >> 4) == 14 ? 2 : (c >> 3) == 30 ? 3 : -1;
        if (extra > 0 && i + extra >= len) extra = -1; // truncated sequence
        if (extra < 0) { out.push_back((int)c); i++; continue; }
        unsigned cp = extra == 0 ? c : c & (0x3Fu >> extra);
        bool ok = true;
        for (int k = 1; k <= extra; k++) {
            if ((p[i + k] & 0xC0) != 0x80) { ok = false; break; }
            cp = (cp << 6) | (p[i + k] & 0x3Fu);
        }
        if (!ok) { out.push_back((int)c); i++; continue; }
        out.push_back((int)cp);
        i += extra + 1;
    }
}

} // namespace

struct b200_tok {
    int kind = 0;
    std::vector<std::string> tokens;
    std::unordered_map<std::string, int> index;
    int byte_token[256];                              // token id of each single mapped byte
    std::unordered_map<uint64_t, int> merges;         // (left id, right id) -> merged token id (= the priority, :57-62 / :213)
    std::vector<std::vector<uint8_t>> token_bytes;    // decode: raw bytes of each token (mapped code points -> bytes)
    std::vector<uint8_t> decodable;                   // 0: the token holds a code point outside the byte alphabet
};

namespace {

inline uint64_t key(int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }

// encodeChunk (LlamaTokenizer.java:199-225): ids in place
void bpe_chunk(const b200_tok *t, std::vector<int> &ids, std::vector<int> &scratch) {
    while (ids.size() >= 2) {
        int best = INT32_MAX, ba = 0, bb = 0;
        for (size_t i = 0; i + 1 < ids.size(); i++) {
            auto it = t->merges.find(key(ids[i], ids[i + 1]));
            if (it != t->merges.end() && it->second < best) { best = it->second; ba = ids[i]; bb = ids[i + 1]; } // strict <: leftmost first occurrence wins ties
        }
        if (best == INT32_MAX) break;
        scratch.clear();
        for (size_t i = 0; i < ids.size();) {
            if (ids[i] == ba && i + 1 < ids.size() && ids[i + 1] == bb) { scratch.push_back(best); i += 2; }
            else { scratch.push_back(ids[i]); i += 1; }
        }
        ids.swap(scratch);
    }
}

inline bool is_contraction(const uint8_t *b, size_t n, size_t i, size_t &len) {
    // (?i:'s|'t|'re|'ve|'m|'ll|'d) in this order; ASCII case-insensitive (bytes of ASCII letters map to themselves)
    if (b[i] != '\'' || i + 1 >= n) return false;
    const int c1 = b[i + 1] | 0x20, c2 = i + 2 < n ? (b[i + 2] | 0x20) : 0;
    const bool a1 = (b[i + 1] >= 'A' && b[i + 1] <= 'Z') || (b[i + 1] >= 'a' && b[i + 1] <= 'z');
    const bool a2 = i + 2 < n && ((b[i + 2] >= 'A' && b[i + 2] <= 'Z') || (b[i + 2] >= 'a' && b[i + 2] <= 'z'));
    if (!a1) return false;
    if (c1 == 's' || c1 == 't') { len = 2; return true; }
    if (c1 == 'r' && a2 && c2 == 'e') { len = 3; return true; }
    if (c1 == 'v' && a2 && c2 == 'e') { len = 3; return true; }
    if (c1 == 'm') { len = 2; return true; }
    if (c1 == 'l' && a2 && c2 == 'l') { len = 3; return true; }
    if (c1 == 'd') { len = 2; return true; }
    return false;
}

// Length of the regex match that starts at byte i (findAll never skips a symbol: one of the alternatives always matches).
size_t next_chunk(int kind, const uint8_t *b, size_t n, size_t i) {
    const ByteTable &T = table();
    size_t len = 0;
    {
        if (!is_contraction(b, n, i, len)) {
            const Cls c = T.cls_of_byte[b[i]];
            if (c == LETTER || (c == OTHER && i + 1 < n && T.cls_of_byte[b[i + 1]] == LETTER)) { // [^\r\n\p{L}\p{N}]?\p{L}+
                size_t j = c == LETTER ? i : i + 1;
                while (j < n && T.cls_of_byte[b[j]] == LETTER) j++;
                len = j - i;
            } else if (c == NUMBER) { // \p{N}{1,3}  |  \p{N}
                const size_t maxd = kind == B200_TOK_QWEN3 ? 1 : 3;
                size_t j = i;
                while (j < n && j - i < maxd && T.cls_of_byte[b[j]] == NUMBER) j++;
                len = j - i;
            } else { //  ?[^\s\p{L}\p{N}]+[\r\n]*  (no literal space, '\r' or '\n' exists in mapped text)
                size_t j = i;
                while (j < n && T.cls_of_byte[b[j]] == OTHER) j++;
                len = j - i;
            }
        }
    }
    return len;
}

// encodeOrdinary on the byte string (each byte IS one mapped symbol)
int encode_bytes(const b200_tok *t, const uint8_t *b, size_t n, std::vector<int> &out) {
    std::vector<int> ids, scratch;
    size_t i = 0;
    while (i < n) {
        const size_t len = next_chunk(t->kind, b, n, i);
        ids.clear();
        for (size_t k = 0; k < len; k++) {
            const int id = t->byte_token[b[i + k]];
            if (id < 0) return B200_TOK_ERR_VOCAB; // vocabulary.getIndex(...).orElseThrow()
            ids.push_back(id);
        }
        bpe_chunk(t, ids, scratch);
        out.insert(out.end(), ids.begin(), ids.end());
        i += len;
    }
    return B200_TOK_OK;
}

int deliver(const std::vector<int> &v, int32_t *ids, size_t cap, size_t *n_out) {
    if (n_out) *n_out = v.size();
    if (v.size() > cap) return B200_TOK_ERR_SPACE;
    if (!v.empty()) memcpy(ids, v.data(), v.size() * sizeof(int32_t));
    return B200_TOK_OK;
}

} // namespace

extern "C" {

int b200_tok_create(const char *const *tokens, int32_t n_tokens, const char *const *merges, int32_t n_merges, int32_t kind, b200_tok **out,
                    char *err, size_t err_len) {
    auto fail = [&](int code, const std::string &msg) {
        if (err && err_len) snprintf(err, err_len, "%s", msg.c_str());
        return code;
    };
    if (out) *out = nullptr;
    if (!tokens || !out || n_tokens <= 0 || n_merges < 0 || (n_merges && !merges) || (kind != B200_TOK_LLAMA3 && kind != B200_TOK_QWEN3))
        return fail(B200_TOK_ERR_BAD_ARG, "bad argument");
    const ByteTable &T = table();
    b200_tok *t = new b200_tok();
    t->kind = kind;
    t->tokens.reserve(n_tokens);
    t->index.reserve((size_t)n_tokens * 2);
    for (int i = 0; i < n_tokens; i++) {
        t->tokens.emplace_back(tokens[i] ? tokens[i] : "");
        t->index[t->tokens.back()] = i; // Collectors.toMap would throw on duplicates; the last one wins here
    }
    for (int b = 0; b < 256; b++) {
        const int cp = T.cp_of_byte[b];
        char buf[4];
        int l = 0;
        if (cp < 0x80) buf[l++] = (char)cp;
        else { buf[l++] = (char)(0xC0 | (cp >> 6)); buf[l++] = (char)(0x80 | (cp & 0x3F)); }
        auto it = t->index.find(std::string(buf, l));
        t->byte_token[b] = it == t->index.end() ? -1 : it->second;
    }
    t->merges.reserve((size_t)n_merges * 2);
    for (int m = 0; m < n_merges; m++) {
        const std::string line = merges[m] ? merges[m] : "";
        const size_t sp = line.find(' ');
        if (sp == std::string::npos) { delete t; return fail(B200_TOK_ERR_VOCAB, "merge line without a space: " + line); }
        const std::string a = line.substr(0, sp), b = line.substr(sp + 1);
        auto ia = t->index.find(a), ib = t->index.find(b), im = t->index.find(a + b);
        if (ia == t->index.end() || ib == t->index.end() || im == t->index.end()) { delete t; return fail(B200_TOK_ERR_VOCAB, "merge refers to a token missing from the vocabulary: " + line); }
        t->merges[key(ia->second, ib->second)] = im->second;
    }
    t->token_bytes.resize(n_tokens);
    t->decodable.assign(n_tokens, 1);
    std::vector<int> cps;
    for (int i = 0; i < n_tokens; i++) {
        cps.clear();
        utf8_to_cps(t->tokens[i].data(), t->tokens[i].size(), cps);
        for (int cp : cps) {
            if (cp >= 0 && cp < 0x144 && T.byte_of_cp[cp] >= 0) t->token_bytes[i].push_back((uint8_t)T.byte_of_cp[cp]);
            else if (kind == B200_TOK_QWEN3 && cp > 512) t->token_bytes[i].push_back((uint8_t)(cp & 0xFF)); // Qwen3Tokenizer.java:311-317
            else { t->decodable[i] = 0; t->token_bytes[i].push_back((uint8_t)'?'); } // BYTE_DECODER.get(cp) == null -> NPE in the reference
        }
    }
    *out = t;
    return B200_TOK_OK;
}

int b200_tok_encode(const b200_tok *t, const char *utf8, size_t len, int32_t *ids, size_t cap, size_t *n_out) {
    if (!t || (!utf8 && len) || (!ids && cap)) return B200_TOK_ERR_BAD_ARG;
    std::vector<int> out;
    out.reserve(len / 3 + 8);
    int rc = encode_bytes(t, reinterpret_cast<const uint8_t *>(utf8), len, out); // byte b <-> mapped symbol cp_of_byte[b]
    if (rc) return rc;
    return deliver(out, ids, cap, n_out);
}

int b200_tok_encode_mapped(const b200_tok *t, const char *mapped, size_t len, int32_t *ids, size_t cap, size_t *n_out) {
    if (!t || (!mapped && len) || (!ids && cap)) return B200_TOK_ERR_BAD_ARG;
    const ByteTable &T = table();
    std::vector<int> cps;
    utf8_to_cps(mapped, len, cps);
    std::vector<uint8_t> bytes;
    bytes.reserve(cps.size());
    for (int cp : cps) {
        if (cp < 0 || cp >= 0x144 || T.byte_of_cp[cp] < 0) return B200_TOK_ERR_VOCAB; // a character that is not a byte symbol: orElseThrow()
        bytes.push_back((uint8_t)T.byte_of_cp[cp]);
    }
    std::vector<int> out;
    int rc = encode_bytes(t, bytes.data(), bytes.size(), out);
    if (rc) return rc;
    return deliver(out, ids, cap, n_out);
}

int b200_tok_decode(const b200_tok *t, const int32_t *ids, size_t n, char *out, size_t cap, size_t *n_out) {
    if (!t || (!ids && n) || (!out && cap)) return B200_TOK_ERR_BAD_ARG;
    size_t need = 0;
    for (size_t i = 0; i < n; i++) {
        if (ids[i] < 0 || (size_t)ids[i] >= t->tokens.size()) return B200_TOK_ERR_BAD_ARG;
        need += t->token_bytes[ids[i]].size();
    }
    if (n_out) *n_out = need;
    if (need > cap) return B200_TOK_ERR_SPACE;
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        const auto &b = t->token_bytes[ids[i]];
        if (!b.empty()) memcpy(out + o, b.data(), b.size());
        o += b.size();
    }
    return B200_TOK_OK;
}

int b200_tok_split(int32_t kind, const char *utf8, size_t len, int32_t *chunk_lens, size_t cap, size_t *n_out) {
    if ((!utf8 && len) || (!chunk_lens && cap) || (kind != B200_TOK_LLAMA3 && kind != B200_TOK_QWEN3)) return B200_TOK_ERR_BAD_ARG;
    const uint8_t *b = reinterpret_cast<const uint8_t *>(utf8);
    size_t i = 0, k = 0;
    while (i < len) {
        const size_t l = next_chunk(kind, b, len, i);
        if (k < cap) chunk_lens[k] = (int32_t)l;
        k++;
        i += l;
    }
    if (n_out) *n_out = k;
    return k > cap ? B200_TOK_ERR_SPACE : B200_TOK_OK;
}

int32_t b200_tok_index(const b200_tok *t, const char *token) {
    if (!t || !token) return -1;
    auto it = t->index.find(token);
    return it == t->index.end() ? -1 : it->second;
}

int32_t b200_tok_vocab_size(const b200_tok *t) { return t ? (int32_t)t->tokens.size() : 0; }
void b200_tok_free(b200_tok *t) { delete t; }

} // extern "C"

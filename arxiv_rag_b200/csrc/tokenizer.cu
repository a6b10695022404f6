// Host-side WordPiece tokenizer behind the C ABI (arb_tokenizer_*): the string half of
// `SentenceTransformer.encode` (generate_embeddings_parallel.py:146-153, text_processor.py:1383-1396),
// multi-threaded so that one process per GPU can feed ~13k chunks/s without a tokenizer farm.
//
// Specification: arxiv_rag_b200/tokenizer.py (itself checked against transformers' MPNet / BERT
// tokenizers); this file must produce the same ids for every row it does not flag. Rows holding one
// of the few code points whose normalisation depends on neighbours (unicode_tables.h: kHard) or
// malformed UTF-8 are flagged and re-done by the caller in Python. No CUDA in this file.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "unicode_tables.h"

namespace {

constexpr int kMaxCharsPerWord = 100;  // tokenizer.py:_MAX_CHARS_PER_WORD (WordPiece max_input_chars_per_word)

inline uint64_t fnv1a(uint64_t h, const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001B3ull;
    return h;
}
constexpr uint64_t kFnvBasis = 0xCBF29CE484222325ull;

struct Vocab {
    // open addressing over token bytes; a continuation piece is stored with its "##" prefix, and
    // looked up by hashing on from the state after "##" so no piece string is ever built
    struct Slot { uint64_t hash; uint32_t off; uint32_t len; int32_t id; };
    std::vector<Slot> slots;
    std::vector<uint8_t> arena;
    uint64_t mask = 0, cont_basis = 0;
    uint32_t max_len = 0;

    void build(const uint8_t* bytes, const int64_t* offs, const int32_t* ids, int32_t n) {
        size_t cap = 16;
        while (cap < size_t(n) * 3) cap <<= 1;
        slots.assign(cap, Slot{0, 0, 0, -1});
        mask = cap - 1;
        arena.assign(bytes, bytes + offs[n]);
        const uint8_t pp[2] = {'#', '#'};
        cont_basis = fnv1a(kFnvBasis, pp, 2);
        for (int32_t i = 0; i < n; ++i) {
            uint32_t off = uint32_t(offs[i]), len = uint32_t(offs[i + 1] - offs[i]);
            uint64_t h = fnv1a(kFnvBasis, arena.data() + off, len);
            size_t s = h & mask;
            while (slots[s].id >= 0 &&
                   !(slots[s].hash == h && slots[s].len == len && !memcmp(arena.data() + slots[s].off, arena.data() + off, len)))
                s = (s + 1) & mask;
            slots[s] = Slot{h, off, len, ids[i]};  // a repeated token keeps the LAST id, like the dict in tokenizer.py
            if (len > max_len) max_len = len;
        }
    }
    inline int32_t find(bool cont, const uint8_t* p, uint32_t n) const {
        const uint32_t len = n + (cont ? 2u : 0u);
        if (len > max_len) return -1;
        const uint64_t h = fnv1a(cont ? cont_basis : kFnvBasis, p, n);
        for (size_t s = h & mask;; s = (s + 1) & mask) {
            const Slot& e = slots[s];
            if (e.id < 0) return -1;
            if (e.hash == h && e.len == len) {
                const uint8_t* k = arena.data() + e.off;
                if (cont ? (k[0] == '#' && k[1] == '#' && !memcmp(k + 2, p, n)) : !memcmp(k, p, n)) return e.id;
            }
        }
    }
};

struct Tokenizer {
    Vocab vocab;
    int32_t cls_id, sep_id, pad_id, unk_id;
    bool lower;
};

inline uint8_t class_of(uint32_t cp) {
    if (cp < 0x80) return arb_uni::kAsciiClass[cp];
    int lo = 0, hi = arb_uni::kNumRanges - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        const arb_uni::Range& r = arb_uni::kRanges[mid];
        if (cp < r.lo) hi = mid - 1;
        else if (cp > r.hi) lo = mid + 1;
        else return r.cls;
    }
    return 0;
}

inline const arb_uni::Fold* fold_of(uint32_t cp) {
    int lo = 0, hi = arb_uni::kNumFolds - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        uint32_t c = arb_uni::kFolds[mid].cp;
        if (cp < c) hi = mid - 1;
        else if (cp > c) lo = mid + 1;
        else return &arb_uni::kFolds[mid];
    }
    return nullptr;
}

inline bool is_cjk(uint32_t cp) {  // tokenizer.py:_is_cjk
    return (cp >= 0x4E00 && cp <= 0x9FFF) || (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x20000 && cp <= 0x2A6DF) ||
           (cp >= 0x2A700 && cp <= 0x2B73F) || (cp >= 0x2B740 && cp <= 0x2B81F) || (cp >= 0x2B820 && cp <= 0x2CEAF) ||
           (cp >= 0xF900 && cp <= 0xFAFF) || (cp >= 0x2F800 && cp <= 0x2FA1F);
}

// One row's state: the word being collected (UTF-8 bytes + the byte offset of every character) and
// the ids so far.
struct Row {
    const Tokenizer& t;
    int32_t* out;      // ids after the leading special token
    int32_t budget;    // max_length - 2
    int32_t n = 0;
    bool full = false;
    uint8_t word[kMaxCharsPerWord * 4 + 8];
    uint16_t start[kMaxCharsPerWord + 2];
    int chars = 0, bytes = 0;
    bool too_long = false;

    Row(const Tokenizer& tk, int32_t* o, int32_t b) : t(tk), out(o), budget(b) { full = budget <= 0; }

    inline void emit(int32_t id) {
        if (n < budget) out[n] = id;
        ++n;
    }
    void flush() {
        if (chars == 0 && !too_long) return;
        if (too_long) {
            emit(t.unk_id);
        } else {
            // greedy longest match first (tokenizer.py:_wordpiece); a word with any unmatched
            // remainder is the unknown token as a whole
            int32_t pieces[kMaxCharsPerWord];
            int np = 0, s = 0;
            bool ok = true;
            start[chars] = uint16_t(bytes);
            while (s < chars) {
                int e = chars, id = -1;
                for (; e > s; --e) {
                    id = t.vocab.find(s > 0, word + start[s], uint32_t(start[e] - start[s]));
                    if (id >= 0) break;
                }
                if (id < 0) { ok = false; break; }
                pieces[np++] = id;
                s = e;
            }
            if (ok) for (int i = 0; i < np; ++i) emit(pieces[i]);
            else emit(t.unk_id);
        }
        chars = bytes = 0;
        too_long = false;
        if (n >= budget) full = true;  // tokenizer.py:encode stops after the word that fills the budget
    }
    inline void push(uint32_t cp) {  // one character of the current word
        if (chars >= kMaxCharsPerWord) { too_long = true; return; }
        start[chars++] = uint16_t(bytes);
        uint8_t* w = word + bytes;
        if (cp < 0x80) { w[0] = uint8_t(cp); bytes += 1; }
        else if (cp < 0x800) { w[0] = 0xC0 | (cp >> 6); w[1] = 0x80 | (cp & 63); bytes += 2; }
        else if (cp < 0x10000) { w[0] = 0xE0 | (cp >> 12); w[1] = 0x80 | ((cp >> 6) & 63); w[2] = 0x80 | (cp & 63); bytes += 3; }
        else { w[0] = 0xF0 | (cp >> 18); w[1] = 0x80 | ((cp >> 12) & 63); w[2] = 0x80 | ((cp >> 6) & 63); w[3] = 0x80 | (cp & 63); bytes += 4; }
    }
    // a character of the NORMALISED text: separator, punctuation (a word of its own) or word character
    inline void normalised(uint32_t cp, uint8_t cls) {
        if (cp == ' ' || (cls & (arb_uni::kWhitespace | arb_uni::kSplit))) flush();
        else if (cls & arb_uni::kPunct) { flush(); if (!full) { push(cp); flush(); } }
        else push(cp);
    }
};

// -> number of ids (without the two specials, before truncation to the budget), or -1: hand the row
// back to the Python implementation
int32_t tokenize_row(const Tokenizer& t, const uint8_t* p, const uint8_t* end, int32_t* out, int32_t budget) {
    Row row(t, out, budget);
    const bool lower = t.lower;
    while (p < end && !row.full) {
        uint32_t cp = *p;
        if (cp < 0x80) {  // ASCII: no table search, no decomposition
            ++p;
            if (cp == ' ' || cp == '\t' || cp == '\n' || cp == '\r') { row.flush(); continue; }
            const uint8_t cls = arb_uni::kAsciiClass[cp];
            if (cp == 0 || (cls & arb_uni::kControl)) continue;
            if (cls & arb_uni::kPunct) { row.flush(); if (!row.full) { row.push(cp); row.flush(); } }
            else row.push(lower ? arb_uni::kAsciiLower[cp] : cp);
            continue;
        }
        int extra;
        if ((cp & 0xE0) == 0xC0) { cp &= 0x1F; extra = 1; }
        else if ((cp & 0xF0) == 0xE0) { cp &= 0x0F; extra = 2; }
        else if ((cp & 0xF8) == 0xF0) { cp &= 0x07; extra = 3; }
        else return -1;
        if (end - p <= extra) return -1;
        for (int i = 1; i <= extra; ++i) {
            if ((p[i] & 0xC0) != 0x80) return -1;
            cp = (cp << 6) | (p[i] & 0x3F);
        }
        p += extra + 1;
        if (cp < 0x80 || cp > 0x10FFFF) return -1;
        const uint8_t cls = class_of(cp);
        if (cls & arb_uni::kHard) return -1;
        if (cp == 0xFFFD || (cls & arb_uni::kControl)) continue;          // step 1: dropped
        if (cls & arb_uni::kWhitespace) { row.flush(); continue; }        //         -> ' '
        const bool cjk = is_cjk(cp);                                       // step 2 looks at the ORIGINAL character
        if (cjk) row.flush();
        if (!lower) {
            row.normalised(cp, cls);
        } else if (cp >= 0xAC00 && cp <= 0xD7A3) {  // Hangul syllable: algorithmic NFD (jamo have no case, none is Mn)
            const uint32_t s = cp - 0xAC00;
            row.push(0x1100 + s / 588);
            row.push(0x1161 + (s % 588) / 28);
            if (s % 28) row.push(0x11A7 + s % 28);
        } else if (const arb_uni::Fold* f = fold_of(cp)) {  // steps 3-4: lower(strip Mn(NFD(c)))
            for (int i = 0; i < f->len && !row.full; ++i) {
                const uint32_t c = arb_uni::kFoldPool[f->off + i];
                row.normalised(c, class_of(c));
            }
        } else {
            row.normalised(cp, cls);
        }
        if (cjk) row.flush();
    }
    if (!row.full) row.flush();
    return row.n;
}

}  // namespace

extern "C" {

int arb_tokenizer_create(const char* token_bytes, const int64_t* token_offsets, const int32_t* token_ids, int32_t n_tokens,
                         int32_t cls_id, int32_t sep_id, int32_t pad_id, int32_t unk_id, int32_t do_lower_case,
                         void** handle_out) {
    ARB_REQUIRE(token_bytes && token_offsets && token_ids && handle_out, "arb_tokenizer_create: null pointer");
    ARB_REQUIRE(n_tokens > 0, "arb_tokenizer_create: empty vocabulary");
    for (int32_t i = 0; i < n_tokens; ++i)
        ARB_REQUIRE(token_offsets[i + 1] >= token_offsets[i] && token_ids[i] >= 0, "arb_tokenizer_create: bad token table at %d", i);
    ARB_REQUIRE(token_offsets[n_tokens] < (int64_t(1) << 31), "arb_tokenizer_create: vocabulary larger than 2 GiB");
    Tokenizer* t = new Tokenizer();
    t->vocab.build(reinterpret_cast<const uint8_t*>(token_bytes), token_offsets, token_ids, n_tokens);
    t->cls_id = cls_id; t->sep_id = sep_id; t->pad_id = pad_id; t->unk_id = unk_id;
    t->lower = do_lower_case != 0;
    *handle_out = t;
    return ARB_OK;
}

int arb_tokenizer_destroy(void* handle) {
    delete static_cast<Tokenizer*>(handle);
    return ARB_OK;
}

int arb_tokenizer_encode(void* handle, const char* text_bytes, const int64_t* text_offsets, int64_t n_texts,
                         int32_t max_length, int32_t num_threads, int32_t* out_ids, int64_t out_stride,
                         int32_t* out_lens, uint8_t* out_fallback) {
    ARB_REQUIRE(handle && text_offsets && out_ids && out_lens && out_fallback, "arb_tokenizer_encode: null pointer");
    ARB_REQUIRE(n_texts >= 0 && (n_texts == 0 || text_bytes), "arb_tokenizer_encode: bad text table");
    ARB_REQUIRE(max_length >= 0 && out_stride >= (max_length > 2 ? max_length : 2),
                "arb_tokenizer_encode: out_stride %lld < max(max_length, 2)", (long long)out_stride);
    const Tokenizer& t = *static_cast<Tokenizer*>(handle);
    const int32_t budget = max_length > 2 ? max_length - 2 : 0;
    std::atomic<int64_t> next{0};
    constexpr int64_t kGrain = 8;
    auto work = [&]() {
        for (;;) {
            const int64_t lo = next.fetch_add(kGrain, std::memory_order_relaxed);
            if (lo >= n_texts) return;
            const int64_t hi = lo + kGrain < n_texts ? lo + kGrain : n_texts;
            for (int64_t r = lo; r < hi; ++r) {
                int32_t* row = out_ids + r * out_stride;
                const uint8_t* b = reinterpret_cast<const uint8_t*>(text_bytes) + text_offsets[r];
                const uint8_t* e = reinterpret_cast<const uint8_t*>(text_bytes) + text_offsets[r + 1];
                int32_t n = e >= b ? tokenize_row(t, b, e, row + 1, budget) : -1;
                out_fallback[r] = n < 0;
                if (n < 0) n = 0;
                if (n > budget) n = budget;
                row[0] = t.cls_id;
                row[n + 1] = t.sep_id;
                for (int64_t j = n + 2; j < out_stride; ++j) row[j] = t.pad_id;
                out_lens[r] = n + 2;
            }
        }
    };
    int nt = num_threads > 0 ? num_threads : int(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (int64_t(nt) * kGrain > n_texts) nt = int((n_texts + kGrain - 1) / kGrain);
    if (nt <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        pool.reserve(nt - 1);
        for (int i = 1; i < nt; ++i) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    return ARB_OK;
}

}  // extern "C"

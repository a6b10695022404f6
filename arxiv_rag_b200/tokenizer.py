"""In-tree WordPiece tokenizer for the two encoders of the reference (host side of `encode`).

`SentenceTransformer.encode` (called at generate_embeddings_parallel.py:146-153 and
text_processor.py:1383-1396) tokenises strings before the forward pass; the drop-in therefore needs
a tokenizer of its own, driven by the model's `vocab.txt` (no vocabulary ships offline, so the
tests build a synthetic one and compare with `transformers`' tokenizers on it).

What is restated (tokenizers' BertNormalizer + BertPreTokenizer + WordPiece + template, as wired by
transformers `MPNetTokenizer` / `BertTokenizer`):
  1. clean: drop NUL, U+FFFD and control characters; every whitespace character becomes ' ';
  2. CJK ideographs get a space on both sides;
  3. strip accents (NFD, drop category Mn) — on whenever lower-casing is on;
  4. lower-case;
  5. split on whitespace, then isolate every punctuation character (ASCII symbols or category P*);
  6. WordPiece per word: greedy longest match, continuation pieces prefixed '##'; a word longer
     than 100 characters, or with any unmatched remainder, becomes the unknown token;
  7. `<s> pieces </s>` (MPNet) / `[CLS] pieces [SEP]` (BERT), truncated to `max_length`.

`tokenize_batch` returns padded int32 `input_ids` / `attention_mask` arrays, like
`tokenizer(texts, padding=True, truncation=True, max_length=..., return_tensors='np')`.

`WordPieceTokenizer` is the pure-Python statement of those steps (~0.5k chunks/s per thread).
`NativeWordPieceTokenizer` runs the same steps in the C-ABI library (`arb_tokenizer_*`,
csrc/tokenizer.cu: multi-threaded, no GIL) and hands the rare rows whose normalisation depends on
context (U+03A3, non-Mn combining marks, lone surrogates) back to the Python class, so the two
return identical arrays; it is what `B200SentenceEncoder(vocab_file=...)` uses.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, Iterable, List, Sequence

import numpy as np

_MAX_CHARS_PER_WORD = 100


def _is_whitespace(ch: str) -> bool:
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in "\t\n\r":
        return False
    return unicodedata.category(ch) in ("Cc", "Cf", "Cn", "Co")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class WordPieceTokenizer:
    """vocab : path of a `vocab.txt` (one token per line, id = line number) or a {token: id} dict.
    kind  : 'mpnet' (<s> </s> <pad>, unknown '[UNK]') or 'bert' ([CLS] [SEP] [PAD] [UNK])."""

    def __init__(self, vocab, kind: str = "mpnet", do_lower_case: bool = True, max_length: int = 384):
        if isinstance(vocab, dict):
            self.vocab: Dict[str, int] = dict(vocab)
        else:
            with open(vocab, "r", encoding="utf-8") as f:
                self.vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        if kind not in ("mpnet", "bert"):
            raise ValueError("kind must be 'mpnet' or 'bert'")
        cls_tok, sep_tok, pad_tok = ("<s>", "</s>", "<pad>") if kind == "mpnet" else ("[CLS]", "[SEP]", "[PAD]")
        for t in (cls_tok, sep_tok, pad_tok, "[UNK]"):
            if t not in self.vocab:
                raise ValueError(f"vocabulary lacks the special token {t!r}")
        self.kind = kind
        self.cls_id, self.sep_id, self.pad_id, self.unk_id = (self.vocab[t] for t in (cls_tok, sep_tok, pad_tok, "[UNK]"))
        self.do_lower_case = do_lower_case
        self.max_length = int(max_length)
        self._word_cache: Dict[str, List[int]] = {}

    # ------------------------------------------------------------------ steps 1-5
    def _normalize(self, text: str) -> str:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_whitespace(ch):
                out.append(" ")
            elif _is_cjk(cp):
                out.extend((" ", ch, " "))
            else:
                out.append(ch)
        s = "".join(out)
        if self.do_lower_case:
            s = "".join(c for c in unicodedata.normalize("NFD", s) if unicodedata.category(c) != "Mn")
            s = s.lower()
        return s

    @staticmethod
    def _pre_tokenize(text: str) -> List[str]:
        words: List[str] = []
        for chunk in text.split():
            cur = []
            for ch in chunk:
                if _is_punctuation(ch):
                    if cur:
                        words.append("".join(cur))
                        cur = []
                    words.append(ch)
                else:
                    cur.append(ch)
            if cur:
                words.append("".join(cur))
        return words

    # ------------------------------------------------------------------ step 6
    def _wordpiece(self, word: str) -> List[int]:
        hit = self._word_cache.get(word)
        if hit is not None:
            return hit
        if len(word) > _MAX_CHARS_PER_WORD:
            ids = [self.unk_id]
        else:
            ids, start, n = [], 0, len(word)
            while start < n:
                end, cur = n, None
                while start < end:
                    piece = word[start:end] if start == 0 else "##" + word[start:end]
                    cur = self.vocab.get(piece)
                    if cur is not None:
                        break
                    end -= 1
                if cur is None:
                    ids = [self.unk_id]
                    break
                ids.append(cur)
                start = end
        if len(self._word_cache) < 1_000_000:
            self._word_cache[word] = ids
        return ids

    # ------------------------------------------------------------------ public
    def encode(self, text: str, max_length: int | None = None) -> List[int]:
        """ids of one text with the special tokens, truncated to `max_length`."""
        max_length = self.max_length if max_length is None else int(max_length)
        ids: List[int] = []
        budget = max(max_length - 2, 0)
        for w in self._pre_tokenize(self._normalize(text)):
            ids.extend(self._wordpiece(w))
            if len(ids) >= budget:
                break
        return [self.cls_id] + ids[:budget] + [self.sep_id]

    def tokenize_batch(self, texts: Sequence[str], max_length: int | None = None):
        """-> (input_ids int32 [n, S], attention_mask int32 [n, S]) padded to the longest row."""
        rows = [self.encode(t, max_length) for t in texts]
        S = max((len(r) for r in rows), default=1)
        ids = np.full((len(rows), S), self.pad_id, np.int32)
        mask = np.zeros((len(rows), S), np.int32)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = r
            mask[i, :len(r)] = 1
        return ids, mask

    def __call__(self, texts, padding=True, truncation=True, max_length=None, return_tensors="np", **_):
        """The call shape `B200SentenceEncoder` uses for an injected HF tokenizer."""
        if isinstance(texts, str):
            texts = [texts]
        ids, mask = self.tokenize_batch(list(texts), max_length if truncation else 1 << 30)
        return {"input_ids": ids, "attention_mask": mask}


class NativeWordPieceTokenizer(WordPieceTokenizer):
    """Same constructor and results as `WordPieceTokenizer`; `tokenize_batch` runs in the library.
    num_threads: worker threads per call (default: min(16, cores / LOCAL_WORLD_SIZE))."""

    def __init__(self, vocab, kind: str = "mpnet", do_lower_case: bool = True, max_length: int = 384,
                 num_threads: int | None = None):
        super().__init__(vocab, kind, do_lower_case, max_length)
        import ctypes as C
        import os

        from . import _lib

        self._libmod, self._C = _lib, C
        # default: up to 16 threads, but only this rank's share of the cores (one process per GPU under torchrun)
        share = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        self.num_threads = int(num_threads) if num_threads else min(16, share)
        toks = [t.encode("utf-8", "surrogatepass") for t in self.vocab]
        offs = np.zeros(len(toks) + 1, np.int64)
        np.cumsum([len(t) for t in toks], out=offs[1:])
        ids = np.fromiter(self.vocab.values(), np.int32, len(toks))
        blob = b"".join(toks)
        h = C.c_void_p()
        _lib.check(_lib.lib().arb_tokenizer_create(blob, offs.ctypes.data, ids.ctypes.data, len(toks), self.cls_id,
                                                   self.sep_id, self.pad_id, self.unk_id, int(do_lower_case), C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                self._libmod.lib().arb_tokenizer_destroy(h)
            except Exception:
                pass
            self._handle = None

    def tokenize_batch(self, texts: Sequence[str], max_length: int | None = None):
        max_length = self.max_length if max_length is None else int(max_length)
        n = len(texts)
        if n == 0:
            return np.full((0, 1), self.pad_id, np.int32), np.zeros((0, 1), np.int32)
        if max_length > 1 << 20:  # "no truncation": the row buffer is sized from the text instead
            max_length = max(len(t) for t in texts) + 2
        raw = [t.encode("utf-8", "surrogatepass") for t in texts]
        offs = np.zeros(n + 1, np.int64)
        np.cumsum([len(b) for b in raw], out=offs[1:])
        stride = max(max_length, 2)
        ids = np.empty((n, stride), np.int32)
        lens = np.empty(n, np.int32)
        fb = np.empty(n, np.uint8)
        self._libmod.check(self._libmod.lib().arb_tokenizer_encode(
            self._handle, b"".join(raw), offs.ctypes.data, n, max_length, self.num_threads, ids.ctypes.data, stride,
            lens.ctypes.data, fb.ctypes.data))
        for r in np.flatnonzero(fb):  # context-dependent rows: the Python statement of the same steps
            row = self.encode(texts[r], max_length)
            ids[r, :len(row)] = row
            ids[r, len(row):] = self.pad_id
            lens[r] = len(row)
        S = int(lens.max())
        mask = (np.arange(S, dtype=np.int32)[None, :] < lens[:, None]).astype(np.int32)
        return np.ascontiguousarray(ids[:, :S]), mask


def text_lengths(texts: Iterable[str]) -> np.ndarray:
    """sentence-transformers sorts by `len(text)` before batching (`_text_length`)."""
    return np.fromiter((len(t) for t in texts), dtype=np.int64)

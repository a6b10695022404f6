"""CPU: the in-tree WordPiece tokenizer (arxiv_rag_b200/tokenizer.py) against transformers'
MPNetTokenizer / BertTokenizer on a synthetic vocabulary (no vocabulary ships offline). This is the
host half of `SentenceTransformer.encode(List[str])` at generate_embeddings_parallel.py:146-153."""
import random

import numpy as np
import pytest

from arxiv_rag_b200.tokenizer import WordPieceTokenizer, text_lengths

BASE = ["<s>", "<pad>", "</s>", "<unk>", "[UNK]", "[CLS]", "[SEP]", "[PAD]", "<mask>", "[MASK]"]
WORDS = ("the quick brown fox jump over lazy dog un believ able cafe resume naive a b c d e f g h i j k l m n o p q r s t u v w x y z "
         "0 1 2 3 4 5 6 7 8 9 . , ! ? - ' \" ( ) [ ] { } / : ; $ % & * + = < > @ # ^ _ ` | ~ 你 好 世 界 arxiv quantum ization transform er "
         "attention istanbul ss").split()
PIECES = ["##" + w for w in "s es ed ing able believ e a b c d 1 2 3 ization er ers ly tion".split()]
VOCAB = {t: i for i, t in enumerate(BASE + WORDS + PIECES)}

TEXTS = [
    "The quick brown foxes jumped over the lazy dog.", "Unbelievable café! 你好 abc123 xyz", "", "   ",
    "résumé naïve — “quoted” text…", "a" * 120 + " b", "arXiv:2101.00001v2 [quant-ph] quantization transformers' attention",
    "tab\tnew\nline\r\x00\x07ctrl ​ zero�width", "İstanbul ǅ ß ÅÉÎÕÜ", "x" * 5 + " " + "the " * 50,
    "$100 (50%) a+b=c <tag> a_b `c` |d| ~e", "世界 你好the", "áè ö",
]


def _random_texts(n, seed=0):
    rng = random.Random(seed)
    alphabet = list("abcdefghij ABC.,!-'éüß你好\t\n ") + ["##", "the", "quick", " ", " ", " ", "…"]
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 80))) for _ in range(n)]


@pytest.mark.parametrize("kind", ["mpnet", "bert"])
def test_matches_transformers_tokenizer(kind):
    from transformers import BertTokenizer, MPNetTokenizer

    hf = MPNetTokenizer(vocab=VOCAB) if kind == "mpnet" else BertTokenizer(vocab=VOCAB)
    mine = WordPieceTokenizer(VOCAB, kind=kind, max_length=32)
    texts = TEXTS + _random_texts(400)
    for t in texts:
        assert mine.encode(t, 32) == hf(t, truncation=True, max_length=32)["input_ids"], repr(t)
    a = mine(texts[:40], padding=True, truncation=True, max_length=24)
    b = hf(texts[:40], padding=True, truncation=True, max_length=24, return_tensors="np")
    assert np.array_equal(a["input_ids"], b["input_ids"]) and np.array_equal(a["attention_mask"], b["attention_mask"])
    assert a["input_ids"].dtype == np.int32


def test_special_tokens_and_truncation():
    tok = WordPieceTokenizer(VOCAB, kind="mpnet", max_length=8)
    ids = tok.encode("the " * 30)
    assert len(ids) == 8 and ids[0] == VOCAB["<s>"] and ids[-1] == VOCAB["</s>"]
    assert tok.encode("") == [VOCAB["<s>"], VOCAB["</s>"]]
    assert tok.encode("zzzz" + "q" * 200) == [VOCAB["<s>"], VOCAB["[UNK]"], VOCAB["</s>"]]  # > 100 characters
    ids, mask = tok.tokenize_batch(["the fox", "a"])
    assert ids.shape == mask.shape == (2, 4) and ids[1, 3] == VOCAB["<pad>"] and mask[1].tolist() == [1, 1, 1, 0]
    with pytest.raises(ValueError):
        WordPieceTokenizer({"a": 0}, kind="mpnet")
    assert text_lengths(["ab", "", "你好!"]).tolist() == [2, 0, 3]


def test_vocab_file_round_trip(tmp_path):
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(VOCAB) + "\n", encoding="utf-8")
    a = WordPieceTokenizer(str(vf), kind="bert")
    b = WordPieceTokenizer(VOCAB, kind="bert")
    for t in TEXTS:
        assert a.encode(t) == b.encode(t)


# ---------------------------------------------------------------------------- native (C ABI) tokenizer
def _native(**kw):
    from arxiv_rag_b200.tokenizer import NativeWordPieceTokenizer

    return NativeWordPieceTokenizer(VOCAB, **kw)


@pytest.mark.parametrize("kind", ["mpnet", "bert"])
def test_native_matches_transformers_tokenizer(kind):
    """arb_tokenizer_encode (csrc/tokenizer.cu) against the HF tokenizers directly."""
    from transformers import BertTokenizer, MPNetTokenizer

    hf = MPNetTokenizer(vocab=VOCAB) if kind == "mpnet" else BertTokenizer(vocab=VOCAB)
    nat = _native(kind=kind, max_length=32, num_threads=3)
    texts = TEXTS + _random_texts(400, seed=5)
    a = nat(texts, padding=True, truncation=True, max_length=24)
    b = hf(texts, padding=True, truncation=True, max_length=24, return_tensors="np")
    assert np.array_equal(a["input_ids"], b["input_ids"]) and np.array_equal(a["attention_mask"], b["attention_mask"])
    assert a["input_ids"].dtype == np.int32 and a["input_ids"].flags.c_contiguous


@pytest.mark.parametrize("lower", [True, False])
def test_native_equals_python_on_every_code_point(lower):
    """Each of the 1,114,112 code points inside and at the end of a word, then random mixed-script
    text (Latin-1, Greek with both sigmas, Cyrillic, CJK + compatibility ideographs, Hangul, combining
    marks incl. the reorderable non-Mn ones, format/control characters, emoji, ligatures, lone
    surrogates): the native rows — including the ones it hands back — equal the Python ones."""
    py = WordPieceTokenizer(VOCAB, "mpnet", lower, 48)
    nat = _native(kind="mpnet", do_lower_case=lower, max_length=48, num_threads=4)
    every = ["a" + chr(cp) + "b c" + chr(cp) for cp in range(0x110000)]
    a, b = py.tokenize_batch(every), nat.tokenize_batch(every)
    assert a[0].shape == b[0].shape and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rng = random.Random(11)
    pools = ["abcdefghij ABCXYZ  \t\n.,;!?()'\"-", "ÀÉÎÕÜßàéîõüÿĀİıŁłŒœſ", "ΑΣΩασςω", "АЯая中文漢字かなカナ한국어豈更",
             "̀́̈ःाாௗ〮⃝", "    　​‍﻿­\x00\x01\x0b\x1c\x7f\x85�",
             "“”—…·、。「」¡¿", "😀🚀\U0001d15e\U0001d1bb𐏿", "ﬁǅẞΐᾈͅǰ"]
    texts = []
    for _ in range(6000):
        pool = pools[0] * 3 + rng.choice(pools) if rng.random() < 0.7 else "".join(pools)
        t = "".join(rng.choice(pool) for _ in range(rng.randint(0, 120)))
        texts.append(t + "x" * rng.randint(95, 105) if rng.random() < 0.05 else t)
    a, b = py.tokenize_batch(texts), nat.tokenize_batch(texts)
    assert a[0].shape == b[0].shape and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_native_edges_and_errors():
    from arxiv_rag_b200 import _lib

    nat = _native(kind="mpnet", max_length=8, num_threads=2)
    ids, mask = nat.tokenize_batch([])
    assert ids.shape == (0, 1) and mask.shape == (0, 1)
    ids, mask = nat.tokenize_batch(["", "the " * 30, "zzzz" + "q" * 200, "the fox"])
    assert ids.shape == (4, 8) and mask.sum(1).tolist() == [2, 8, 3, 4]
    assert ids[0, :3].tolist() == [VOCAB["<s>"], VOCAB["</s>"], VOCAB["<pad>"]] and ids[2, 1] == VOCAB["[UNK]"]
    assert np.array_equal(nat.tokenize_batch(["the fox"], max_length=1)[0], [[VOCAB["<s>"], VOCAB["</s>"]]])  # like the Python class
    # untruncated call shape (truncation=False)
    long = "the " * 700
    assert nat([long], truncation=False)["input_ids"].shape == (1, 702)
    # the C ABI itself: a row that is not UTF-8 is flagged, not tokenised; a short stride is refused
    import ctypes as C
    raw = b"the \xff fox"
    offs = np.array([0, len(raw)], np.int64)
    out, lens, fb = np.zeros((1, 8), np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint8)
    lib = _lib.lib()
    _lib.check(lib.arb_tokenizer_encode(nat._handle, raw, offs.ctypes.data, 1, 8, 1, out.ctypes.data, 8, lens.ctypes.data, fb.ctypes.data))
    assert fb[0] == 1 and lens[0] == 2
    with pytest.raises(_lib.ArbError):
        _lib.check(lib.arb_tokenizer_encode(nat._handle, raw, offs.ctypes.data, 1, 8, 1, out.ctypes.data, 4, lens.ctypes.data, fb.ctypes.data))
    h = C.c_void_p()
    with pytest.raises(_lib.ArbError):
        _lib.check(lib.arb_tokenizer_create(b"", offs.ctypes.data, offs.ctypes.data, 0, 0, 1, 2, 3, 1, C.byref(h)))

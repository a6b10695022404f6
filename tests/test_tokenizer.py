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

"""CLIP tokenizer (csrc/tokenizer.h) against Hugging Face's own CLIPTokenizer -- the class the
reference instantiates through AutoTokenizer (R/scripts/evaluate.py:155-156) -- on SYNTHETIC vocabularies: no real
vocab.json / merges.txt exists in the image (no network), so a small byte-level BPE is trained here, handed to both
implementations, and ids + attention masks must be identical on ASCII, accented, composed / decomposed, CJK, Arabic,
emoji, Greek (final sigma), Turkish dotted I, contractions, digits, white-space varieties, embedded special tokens,
empty and over-long texts.  CPU only: tokenisation is host work in the reference too."""
import collections
import json
import random

import pytest
import torch


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


CORPUS = ("the quick brown fox jumps over the lazy dog and then the dog's owner said it's fine they're all here we've "
          "seen this before i'm sure you'll agree he'd rather not 12 34 2024 covid19 #hashtag @user http://t.co/abc :) :( "
          "!!! ... café naïve résumé über straße ΣΑΣ σας "
          "ελληνικά русский текст "
          "日本語 のテキスト 中文 文本 العربية "
          "\U0001f600 \U0001f602 ✌\U0001f3fd \U0001f44d\U0001f3ff istanbul i̇stanbul hate speech is not ok racist "
          "sexist homophobe religion other hate meme tweet image text multimodal content moderation").split()


def _train_bpe(words, n_merges):
    """A minimal byte-level BPE trainer (frequency-greedy) -- only to obtain a plausible vocab / merge list."""
    b2u = _bytes_to_unicode()
    vocab = collections.Counter()
    for w in words:
        sym = [b2u[b] for b in w.lower().encode("utf-8")]
        sym[-1] += "</w>"
        vocab[tuple(sym)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for sym, f in vocab.items():
            for a, b in zip(sym, sym[1:]):
                pairs[(a, b)] += f
        if not pairs:
            break
        (a, b), _f = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append((a, b))
        new = collections.Counter()
        for sym, f in vocab.items():
            out, i = [], 0
            while i < len(sym):
                if i + 1 < len(sym) and sym[i] == a and sym[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(sym[i])
                    i += 1
            new[tuple(out)] += f
        vocab = new
    return merges


@pytest.fixture(scope="module")
def toks(tmp_path_factory):
    from transformers import CLIPTokenizer
    import mmcm_b200 as P
    b2u = _bytes_to_unicode()
    merges = _train_bpe(CORPUS * 3, 600)
    vocab = {}
    for c in b2u.values():
        vocab[c] = len(vocab)
    for c in b2u.values():
        vocab[c + "</w>"] = len(vocab)
    for a, b in merges:
        if a + b not in vocab:
            vocab[a + b] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    d = tmp_path_factory.mktemp("clip_tok")
    (d / "vocab.json").write_text(json.dumps(vocab, ensure_ascii=True))          # \uXXXX escapes, like the hub's file
    (d / "merges.txt").write_text("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in merges) + "\n", encoding="utf-8")
    hf = CLIPTokenizer(vocab=vocab, merges=merges)
    mine = P.ClipTokenizer(str(d / "vocab.json"), str(d / "merges.txt"))
    return hf, mine


EDGE_TEXTS = [
    "", " ", "   \t\n ", "a", "Hello, World!", "it's they're we've I'm you'll he'd 'S 'T 'LL don't DON'T",
    "the dog's owner's dogs'", "x<|startoftext|>y<|endoftext|>z", "<|endoftext|>", "<|STARTOFTEXT|> upper",
    "İstanbul İİ I ı",                                       # Turkish dotted / dotless i
    "ΣΑΣ ΟΔΟΣ Σ ΑΣ. ΣΣ",  # Greek capital sigma, final-sigma rule
    "ÅÅ é é ñ ñ ṩ ṩ ǆ ﬁ ㎒",   # composed vs decomposed forms
    "한국어 한글 각",                 # Hangul syllables and conjoining jamo
    "日本語のテキスト 中文文本", "العربية نص",
    "\U0001f600\U0001f602 ✌\U0001f3fd\U0001f44d\U0001f3ff \U0001f1e9\U0001f1ea",
    "12 3.14 1,000 ②③ ½ ٣٤",
    "tab\tsep nbsp em ideo　ls nel",
    "a" * 300, "word " * 200, "!!!???...---___", "#hash @user http://t.co/AbC123 :-) ;) <3",
    "mixed CASE Text WITH Ünïcödé ÀÉÎÕÜ", "́̀ combining first",
    "ẹ́ ạ́ q̣̇",                                 # canonical reordering of marks
    "zero​width‍joiner﻿bom", "\x01 control \x7f",
]


def _random_texts(n, seed):
    rnd = random.Random(seed)
    pools = ["abcdefghijklmnopqrstuvwxyz", "ABCDEFGHIJKLMNOPQRSTUVWXYZ", "0123456789", " \t\n  ", ".,!?'\"#@:;()-_/<>|",
             "àáâãäåæçèéêëìíîïñòóôõöùúûüýÿßœ",
             "ÀÁÂÃÄÅÆÇÈÉÊËÌÍÎÏÑÒÓÔÕÖÙÚÛÜÝŸŒ",
             "αβγδεζηθικλμνξοπρστυφχψωςΣΑΒΓΔ",
             "абвгдежзийклмнопАБВГДЕЖЗ",
             "日本語中文한국어テキストのは",
             "ابتثجحخدذرزسش",
             "\U0001f600\U0001f602\U0001f923\U0001f60d\U0001f525\U0001f4af\U0001f44d\U0001f3fd✌\U0001f3ff",
             "̧̣̀́̂̃̇̈", "İıǅǆﬁﬂ½²③"]
    out = []
    for _ in range(n):
        k = rnd.randint(0, 60)
        s = []
        for _ in range(k):
            pool = rnd.choice(pools[:5]) if rnd.random() < 0.7 else rnd.choice(pools)
            s.append("".join(rnd.choice(pool) for _ in range(rnd.randint(1, 6))))
            if rnd.random() < 0.5:
                s.append(" ")
            if rnd.random() < 0.05:
                s.append(rnd.choice(["'s", "'re", "'ll", "<|endoftext|>", "<|startoftext|>", "it's"]))
        out.append("".join(s))
    return out


@pytest.mark.parametrize("max_len", [77, 16, 2])
def test_matches_hf_clip_tokenizer(toks, max_len):
    hf, mine = toks
    assert (mine.bos_token_id, mine.eos_token_id, mine.pad_token_id) == (hf.bos_token_id, hf.eos_token_id, hf.pad_token_id)
    texts = EDGE_TEXTS + _random_texts(1500, 7)
    want = hf(texts, padding="max_length", truncation=True, max_length=max_len, return_attention_mask=True,
              return_tensors="pt")
    got = mine(texts, padding="max_length", truncation=True, max_length=max_len, return_attention_mask=True,
               return_tensors="pt")
    bad = [(i, texts[i]) for i in range(len(texts)) if not torch.equal(got["input_ids"][i], want["input_ids"][i])
           or not torch.equal(got["attention_mask"][i], want["attention_mask"][i])]
    assert not bad, f"{len(bad)} texts differ, e.g. {bad[0][1]!r}: {got['input_ids'][bad[0][0]].tolist()} vs " \
                    f"{want['input_ids'][bad[0][0]].tolist()}"
    assert got["input_ids"].dtype == torch.int64 and got["input_ids"].shape == (len(texts), max_len)


def test_single_text_call_has_the_references_shape(toks):
    """R/src/data/dataset.py:148-157: `tok(text, ...)["input_ids"][0]` is the sample's row."""
    hf, mine = toks
    a = mine("the dog's owner", padding="max_length", truncation=True, max_length=77, return_attention_mask=True,
             return_tensors="pt")
    b = hf("the dog's owner", padding="max_length", truncation=True, max_length=77, return_attention_mask=True,
           return_tensors="pt")
    assert a["input_ids"].shape == (1, 77) and torch.equal(a["input_ids"], b["input_ids"])
    assert torch.equal(a["attention_mask"], b["attention_mask"])
    # the EOS-pooling contract of the text tower: first id == eos marks the pooled position, padding repeats eos
    row = a["input_ids"][0]
    first_eos = int((row == mine.eos_token_id).nonzero()[0])
    assert int(a["attention_mask"][0].sum()) == first_eos + 1 and (row[first_eos:] == mine.eos_token_id).all()


def test_threads_and_errors(toks, tmp_path):
    import mmcm_b200 as P
    hf, mine = toks
    texts = _random_texts(700, 11)
    mine.n_threads = 1
    a = mine.encode_batch(texts, 77)
    mine.n_threads = 8
    b = mine.encode_batch(texts, 77)
    mine.n_threads = 0
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    with pytest.raises(ValueError, match="cannot read vocabulary"):
        P.ClipTokenizer(str(tmp_path / "missing.json"), str(tmp_path / "missing.txt"))
    (tmp_path / "v.json").write_text('{"a": 0, "b": 1}')
    (tmp_path / "m.txt").write_text("a b\n")
    with pytest.raises(ValueError, match="startoftext"):
        P.ClipTokenizer(str(tmp_path / "v.json"), str(tmp_path / "m.txt"))
    with pytest.raises(ValueError, match="padding='max_length'"):
        mine("x", padding=True)

#!/usr/bin/env python
"""Texts/s of the C++ CLIP tokenizer (mmcm_tokenizer_encode) against Hugging Face's CLIPTokenizer (`tokenizers`, Rust,
batched) on tweet-like synthetic texts, pad-to-77 -- the call of R/src/data/dataset.py:148-155.  Host only.

    python tools/tokenizer_bench.py [n_texts=20000]
"""
import json
import os
import random
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
import torch  # noqa: E402
from transformers import CLIPTokenizer  # noqa: E402
from test_tokenizer_cpu import CORPUS, _bytes_to_unicode, _train_bpe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b2u = _bytes_to_unicode()
merges = _train_bpe(CORPUS * 3, 600)
vocab = {}
for c in b2u.values():
    vocab[c] = len(vocab)
for c in b2u.values():
    vocab[c + "</w>"] = len(vocab)
for a, b in merges:
    vocab.setdefault(a + b, len(vocab))
vocab["<|startoftext|>"] = len(vocab)
vocab["<|endoftext|>"] = len(vocab)
d = tempfile.mkdtemp()
open(os.path.join(d, "vocab.json"), "w").write(json.dumps(vocab))
open(os.path.join(d, "merges.txt"), "w", encoding="utf-8").write("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in merges) + "\n")
hf = CLIPTokenizer(vocab=vocab, merges=merges)
mine = P.ClipTokenizer(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"))
rnd = random.Random(0)
texts = [" ".join(rnd.choice(CORPUS) for _ in range(rnd.randint(3, 40))) for _ in range(n)]
kw = dict(padding="max_length", truncation=True, max_length=77, return_attention_mask=True, return_tensors="pt")


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


t_hf, a = best(lambda: hf(texts, **kw))
t_loop, _ = best(lambda: [hf(t, **kw) for t in texts[:2000]], reps=2)       # the reference's per-sample call
print(f"cores {os.cpu_count()}  texts {n}")
print(f"HF CLIPTokenizer, one batched call      : {n / t_hf:10.0f} texts/s")
print(f"HF CLIPTokenizer, per-sample calls (ref): {2000 / t_loop:10.0f} texts/s")
for th in (1, 4, 0):
    mine.n_threads = th
    t, b = best(lambda: mine(texts, **kw))
    assert torch.equal(a["input_ids"], b["input_ids"]) and torch.equal(a["attention_mask"], b["attention_mask"])
    print(f"mmcm_tokenizer_encode, threads={th or os.cpu_count():<3d}        : {n / t:10.0f} texts/s")

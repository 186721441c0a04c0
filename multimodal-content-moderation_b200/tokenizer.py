"""CLIP tokenizer in the callers' place: `mmcm_tokenizer_*` of include/mmcm.h behind the call
signature the reference uses.

    tok = ClipTokenizer("vocab.json", "merges.txt")            # the checkpoint's own tokenizer files
    enc = tok(texts, padding="max_length", truncation=True, max_length=77, return_attention_mask=True, return_tensors="pt")
    enc["input_ids"], enc["attention_mask"]                      # int64 [N, 77], as R/src/data/dataset.py:148-165 expects

A single string behaves like the reference's per-sample call (`tok["input_ids"][0]` is that sample's row); a list is
encoded as one multi-threaded batch.  Only what the reference's callers use is supported: fixed-length padding with
truncation.  The algorithm is the one Hugging Face's CLIPTokenizer configures (csrc/tokenizer.h).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Union

import torch

from . import lib as L


class ClipTokenizer:
    model_input_names = ["input_ids", "attention_mask"]

    def __init__(self, vocab_file: str, merges_file: str, n_threads: int = 0):
        self.lib = L.load()
        self._h = C.c_void_p()
        L.check(self.lib.mmcm_tokenizer_create(str(vocab_file).encode(), str(merges_file).encode(), C.byref(self._h)))
        v, b, e, p = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        L.check(self.lib.mmcm_tokenizer_info(self._h, C.byref(v), C.byref(b), C.byref(e), C.byref(p)))
        self.vocab_size, self.bos_token_id, self.eos_token_id, self.pad_token_id = v.value, b.value, e.value, p.value
        self.unk_token_id = e.value
        self.n_threads = int(n_threads)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mmcm_tokenizer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def encode_batch(self, texts: Sequence[str], max_length: int = 77):
        """-> (input_ids, attention_mask), int64 [N, max_length] host tensors."""
        n = len(texts)
        raw: List[bytes] = [(t or "").encode("utf-8") for t in texts]
        arr = (C.c_char_p * max(n, 1))(*raw)
        lens = (C.c_int64 * max(n, 1))(*[len(r) for r in raw])
        ids = torch.empty((n, max_length), dtype=torch.int64)
        mask = torch.empty((n, max_length), dtype=torch.int64)
        L.check(self.lib.mmcm_tokenizer_encode(self._h, arr, lens, n, int(max_length), C.c_void_p(ids.data_ptr()),
                                               C.c_void_p(mask.data_ptr()), self.n_threads))
        return ids, mask

    def __call__(self, text: Union[str, Sequence[str]], padding="max_length", truncation=True, max_length: int = 77,
                 return_attention_mask: bool = True, return_tensors="pt", **unused):
        if padding != "max_length" or not truncation:
            raise ValueError("ClipTokenizer supports padding='max_length' with truncation=True (what the reference's "
                             "callers use: R/src/data/dataset.py:148-155)")
        texts = [text] if isinstance(text, str) else list(text)
        ids, mask = self.encode_batch(texts, max_length)
        if return_tensors not in ("pt", None):
            raise ValueError("return_tensors must be 'pt' or None")
        if return_tensors is None:
            ids, mask = ids.tolist(), mask.tolist()
            if isinstance(text, str):
                ids, mask = ids[0], mask[0]
        out = {"input_ids": ids}
        if return_attention_mask:
            out["attention_mask"] = mask
        return out

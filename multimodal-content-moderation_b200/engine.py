"""Thin Python owner of one `mmcm_handle` (include/mmcm.h).

Holds no arithmetic: it converts torch tensors to raw pointers, forwards the call on torch's current CUDA
stream and maps C status codes to the exceptions the reference raises for the same conditions.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import arch as A
from . import lib as L


def _cfg_struct(a: A.ArchCfg, head: int, num_outputs: int, fusion_dim: int, head_hidden_dim: int) -> L.MmcmConfig:
    c = L.MmcmConfig()
    c.backend, c.head = a.backend, head
    c.text_hidden, c.text_heads, c.text_layers, c.text_ffn, c.text_act = (
        a.text.hidden, a.text.heads, a.text.layers, a.text.ffn, a.text.act)
    c.vis_hidden, c.vis_heads, c.vis_layers, c.vis_ffn, c.vis_act = (
        a.vision.hidden, a.vision.heads, a.vision.layers, a.vision.ffn, a.vision.act)
    c.text_eps, c.vis_eps = a.text.eps, a.vision.eps
    c.vocab, c.max_pos, c.eos_id = a.vocab, a.max_pos, a.eos_id
    c.image, c.patch, c.proj_dim = a.image, a.patch, a.proj_dim
    c.fusion_dim, c.num_outputs, c.head_hidden_dim = fusion_dim, num_outputs, head_hidden_dim or 0
    return c


def _arch_from_struct(c: L.MmcmConfig) -> A.ArchCfg:
    """Inverse of `_cfg_struct` for the shape fields (a packed weight file carries the struct, not the dataclass)."""
    return A.ArchCfg(
        backend=c.backend,
        text=A.TowerCfg(c.text_hidden, c.text_heads, c.text_layers, c.text_ffn, c.text_eps, c.text_act),
        vision=A.TowerCfg(c.vis_hidden, c.vis_heads, c.vis_layers, c.vis_ffn, c.vis_eps, c.vis_act),
        vocab=c.vocab, max_pos=c.max_pos, eos_id=c.eos_id, image=c.image, patch=c.patch, proj_dim=c.proj_dim)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One scoring engine bound to one CUDA device."""

    def __init__(self, a: A.ArchCfg, head: int, num_outputs: int, fusion_dim: int = 512,
                 head_hidden_dim: int = 0, device: int = 0):
        self.lib = L.load()
        self.arch, self.head, self.num_outputs = a, head, num_outputs
        self.device = int(device)
        self._h = C.c_void_p()
        cfg = _cfg_struct(a, head, num_outputs, fusion_dim, head_hidden_dim)
        L.check(self.lib.mmcm_create(C.byref(cfg), self.device, C.byref(self._h)))

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mmcm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Push every floating-point entry of a reference state dict; the library repacks its own copy."""
        for key, t in sd.items():
            if not torch.is_tensor(t) or not t.is_floating_point():
                continue  # e.g. `embeddings.position_ids` (int64 buffer in transformers<4.31 checkpoints)
            src = t.detach().to(torch.float32).contiguous()
            L.check(self.lib.mmcm_load_weight(self._h, key.encode(), src.data_ptr(), src.numel()))
        L.check(self.lib.mmcm_finalize_weights(self._h))

    def set_option(self, name: str, value: int) -> None:
        L.check(self.lib.mmcm_set_option(self._h, name.encode(), int(value)))

    # ------------------------------------------------------------------ packed weight file
    def save_packed(self, path: str) -> None:
        """Write the repacked (bf16 / fp32) weight set of this engine as one blob (`mmcm_save_packed`)."""
        L.check(self.lib.mmcm_save_packed(self._h, str(path).encode()))

    def load_packed(self, path: str) -> None:
        """Fill this engine from a blob written by `save_packed` for the same configuration."""
        L.check(self.lib.mmcm_load_packed(self._h, str(path).encode()))

    @classmethod
    def from_packed(cls, path: str, device: int = 0) -> "Engine":
        """Build an engine from a packed weight file alone: configuration from its header, weights by one copy."""
        lib = L.load()
        cfg = L.MmcmConfig()
        L.check(lib.mmcm_packed_config(str(path).encode(), C.byref(cfg)))
        eng = cls(_arch_from_struct(cfg), cfg.head, cfg.num_outputs, cfg.fusion_dim, cfg.head_hidden_dim, device)
        eng.load_packed(path)
        return eng

    # ------------------------------------------------------------------ hot path
    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], pixel_values: torch.Tensor,
                text_present: torch.Tensor, image_present: torch.Tensor, want_probs: bool = False):
        """Device tensors in, device logits (and optionally sigmoid probabilities) out; enqueued on the current stream."""
        for nm, t in (("input_ids", input_ids), ("pixel_values", pixel_values), ("text_present", text_present),
                      ("image_present", image_present), ("attention_mask", attention_mask)):
            if t is not None and not t.is_cuda:
                raise RuntimeError(f"{nm} must be a CUDA tensor: the B200 scoring path has no CPU fallback")
        a = self.arch
        if pixel_values.dim() != 4 or pixel_values.shape[1] != 3 or pixel_values.shape[2] != a.image \
                or pixel_values.shape[3] != a.image:
            # HF/models/clip/modeling_clip.py:204-207
            raise ValueError(f"Input image size ({pixel_values.shape[-2]}*{pixel_values.shape[-1]}) doesn't match model "
                             f"({a.image}*{a.image}).")
        B, S = input_ids.shape
        dev = input_ids.device
        ids = input_ids.to(torch.int64).contiguous()
        mask = None if attention_mask is None else attention_mask.to(torch.int64).contiguous()
        px = pixel_values.to(torch.float32).contiguous()
        tp = text_present.to(torch.float32).contiguous()
        ip = image_present.to(torch.float32).contiguous()
        if px.shape[0] != B or tp.numel() != B or ip.numel() != B or (mask is not None and mask.shape != ids.shape):
            raise ValueError("batch dimensions of the inputs disagree")
        logits = torch.empty((B, self.num_outputs), dtype=torch.float32, device=dev)
        probs = torch.empty_like(logits) if want_probs else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        L.check(self.lib.mmcm_forward(self._h, ids.data_ptr(), _ptr(mask), px.data_ptr(), tp.data_ptr(), ip.data_ptr(),
                                      B, S, logits.data_ptr(), _ptr(probs), C.c_void_p(stream)))
        # the inputs above may be temporaries (dtype casts): keep them alive until the stream has consumed them
        for t in (ids, mask, px, tp, ip):
            if t is not None:
                t.record_stream(torch.cuda.current_stream(dev))
        return (logits, probs) if want_probs else logits

    def forward_host(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], pixel_values: torch.Tensor,
                     text_present: torch.Tensor, image_present: torch.Tensor, want_probs: bool = False,
                     out: Optional[torch.Tensor] = None):
        """HOST tensors in (pinned recommended), HOST logits out: H2D, towers, head and D2H inside one call."""
        for t in (input_ids, attention_mask, pixel_values, text_present, image_present):
            if t is not None and t.is_cuda:
                raise RuntimeError("forward_host takes host tensors")
        B, S = input_ids.shape
        ids = input_ids.to(torch.int64).contiguous()
        mask = None if attention_mask is None else attention_mask.to(torch.int64).contiguous()
        px = pixel_values.to(torch.float32).contiguous()
        tp = text_present.to(torch.float32).contiguous()
        ip = image_present.to(torch.float32).contiguous()
        logits = out if out is not None else torch.empty((B, self.num_outputs), dtype=torch.float32).pin_memory()
        probs = torch.empty((B, self.num_outputs), dtype=torch.float32).pin_memory() if want_probs else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.mmcm_forward_host(self._h, ids.data_ptr(), _ptr(mask), px.data_ptr(), tp.data_ptr(),
                                           ip.data_ptr(), B, S, logits.data_ptr(), _ptr(probs), C.c_void_p(stream)))
        return (logits, probs) if want_probs else logits

    def prefetch_host(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], pixels: torch.Tensor,
                      text_present: torch.Tensor, image_present: torch.Tensor) -> None:
        """Ship the NEXT batch to the device now (copy engine), then call `forward_host` / `forward_host_u8` for the
        current one: the following forward_host* call on these SAME tensors finds its inputs on the device.  `pixels`
        is the fp32 `pixel_values` tensor or the uint8 HWC image tensor.  The tensors must be host tensors of the final
        dtype (int64 ids / mask, fp32 flags), contiguous -- no conversion copy may sit between them and the call that
        consumes them -- and are kept alive by the engine until then."""
        want = ((input_ids, torch.int64), (attention_mask, torch.int64), (text_present, torch.float32),
                (image_present, torch.float32))
        for t, dt in want:
            if t is not None and (t.is_cuda or t.dtype != dt or not t.is_contiguous()):
                raise ValueError("prefetch_host needs contiguous host tensors of dtype int64 (ids, mask) / float32 (flags)")
        if pixels.is_cuda or not pixels.is_contiguous() or pixels.dtype not in (torch.float32, torch.uint8):
            raise ValueError("prefetch_host needs a contiguous host fp32 pixel_values or uint8 image tensor")
        B, S = input_ids.shape
        if pixels.dtype == torch.uint8:
            self._check_u8(pixels, B)
            fn = self.lib.mmcm_prefetch_host_u8
        else:
            fn = self.lib.mmcm_prefetch_host
        L.check(fn(self._h, input_ids.data_ptr(), _ptr(attention_mask), pixels.data_ptr(), text_present.data_ptr(),
                   image_present.data_ptr(), B, S))
        self._prefetched = (input_ids, attention_mask, pixels, text_present, image_present)

    def score_host_batches(self, batches, want_probs: bool = False):
        """The reference's evaluation loop (R/scripts/evaluate.py:163-183: `for batch in loader: model(**batch)`) over
        HOST batches -- dicts with the collate_fn keys of R/src/data/dataset.py:171-193, pinned tensors of the final
        dtypes (what a `DataLoader(pin_memory=True)` yields) -- as a generator of host logits (or (logits, probs)), one
        per batch: batch i+1 is shipped with `prefetch_host` while batch i is scored, so the H2D copies run on the copy
        engine behind the towers.  Same logits, bit for bit, as one `forward_host` per batch."""
        order = ("input_ids", "attention_mask", "pixel_values", "text_present", "image_present")
        it = iter(batches)
        cur = next(it, None)
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                self.prefetch_host(*[nxt.get(k) for k in order])
            yield self.forward_host(*[cur.get(k) for k in order], want_probs=want_probs)
            cur = nxt

    # ------------------------------------------------------------------ uint8 pixel source
    def _check_u8(self, images_u8: torch.Tensor, B: int) -> torch.Tensor:
        a = self.arch
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("images_u8 must be a uint8 [B, H, W, 3] tensor")
        if images_u8.shape[1] != a.image or images_u8.shape[2] != a.image:
            raise ValueError(f"Input image size ({images_u8.shape[1]}*{images_u8.shape[2]}) doesn't match model "
                             f"({a.image}*{a.image}).")
        if images_u8.shape[0] != B:
            raise ValueError("batch dimensions of the inputs disagree")
        return images_u8.contiguous()

    @staticmethod
    def _norm_consts(mean, std):
        if len(mean) != 3 or len(std) != 3:
            raise ValueError("mean and std need three entries")
        return (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])

    def forward_u8(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], images_u8: torch.Tensor,
                   mean, std, text_present: torch.Tensor, image_present: torch.Tensor, want_probs: bool = False):
        """`forward` on raw uint8 HWC images: ToTensor + Normalize happen inside the patch im2col (bit-identical to
        `forward(pixel_values=prepost.preprocess_u8(images_u8, mean, std))`)."""
        for nm, t in (("input_ids", input_ids), ("images_u8", images_u8), ("text_present", text_present),
                      ("image_present", image_present), ("attention_mask", attention_mask)):
            if t is not None and not t.is_cuda:
                raise RuntimeError(f"{nm} must be a CUDA tensor: the B200 scoring path has no CPU fallback")
        B, S = input_ids.shape
        dev = input_ids.device
        px = self._check_u8(images_u8, B)
        ids = input_ids.to(torch.int64).contiguous()
        mask = None if attention_mask is None else attention_mask.to(torch.int64).contiguous()
        tp = text_present.to(torch.float32).contiguous()
        ip = image_present.to(torch.float32).contiguous()
        if tp.numel() != B or ip.numel() != B or (mask is not None and mask.shape != ids.shape):
            raise ValueError("batch dimensions of the inputs disagree")
        m3, s3 = self._norm_consts(mean, std)
        logits = torch.empty((B, self.num_outputs), dtype=torch.float32, device=dev)
        probs = torch.empty_like(logits) if want_probs else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        L.check(self.lib.mmcm_forward_u8(self._h, ids.data_ptr(), _ptr(mask), px.data_ptr(), m3, s3, tp.data_ptr(),
                                         ip.data_ptr(), B, S, logits.data_ptr(), _ptr(probs), C.c_void_p(stream)))
        for t in (ids, mask, px, tp, ip):
            if t is not None:
                t.record_stream(torch.cuda.current_stream(dev))
        return (logits, probs) if want_probs else logits

    def forward_host_u8(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], images_u8: torch.Tensor,
                        mean, std, text_present: torch.Tensor, image_present: torch.Tensor, want_probs: bool = False,
                        out: Optional[torch.Tensor] = None):
        """`forward_host` on raw uint8 HWC host images (pinned recommended): 4x less H2D traffic than fp32 pixels."""
        for t in (input_ids, attention_mask, images_u8, text_present, image_present):
            if t is not None and t.is_cuda:
                raise RuntimeError("forward_host_u8 takes host tensors")
        B, S = input_ids.shape
        px = self._check_u8(images_u8, B)
        ids = input_ids.to(torch.int64).contiguous()
        mask = None if attention_mask is None else attention_mask.to(torch.int64).contiguous()
        tp = text_present.to(torch.float32).contiguous()
        ip = image_present.to(torch.float32).contiguous()
        m3, s3 = self._norm_consts(mean, std)
        logits = out if out is not None else torch.empty((B, self.num_outputs), dtype=torch.float32).pin_memory()
        probs = torch.empty((B, self.num_outputs), dtype=torch.float32).pin_memory() if want_probs else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.mmcm_forward_host_u8(self._h, ids.data_ptr(), _ptr(mask), px.data_ptr(), m3, s3,
                                              tp.data_ptr(), ip.data_ptr(), B, S, logits.data_ptr(), _ptr(probs),
                                              C.c_void_p(stream)))
        return (logits, probs) if want_probs else logits

    # ------------------------------------------------------------------ introspection
    def stage(self, name: str) -> torch.Tensor:
        n = C.c_int64(0)
        L.check(self.lib.mmcm_get_stage(self._h, name.encode(), None, 0, C.byref(n), None))
        out = torch.empty(max(n.value, 1), dtype=torch.float32, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(out.device).cuda_stream
        L.check(self.lib.mmcm_get_stage(self._h, name.encode(), out.data_ptr(), out.numel(), C.byref(n),
                                        C.c_void_p(stream)))
        return out[: n.value]

    def last_launch_count(self) -> int:
        return int(self.lib.mmcm_last_launch_count(self._h))

    def last_host_copy_share(self) -> float:
        """Share of the last forward_host* call during which its H2D pixel copies were running (> 0.85: H2D bound)."""
        return float(self.lib.mmcm_last_host_copy_share(self._h))

    def last_chunks(self):
        """(text, vision) micro-batch sizes the last forward split the batch into."""
        t, v = C.c_int32(0), C.c_int32(0)
        L.check(self.lib.mmcm_last_chunks(self._h, C.byref(t), C.byref(v)))
        return t.value, v.value

    def gemm_time(self, epilogue: Optional[int] = None, N: int = 0, K: int = 0):
        """(ms, executed FLOPs, launches) of the GEMM launches of the last forward run with option time_gemms = 1;
        `epilogue` restricts the sum to one MMCM_EPI_* kind (then algorithmic bytes are returned too), `N` / `K` to
        one weight shape."""
        ms, fl, by, n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
        if N or K:
            L.check(self.lib.mmcm_gemm_time_shape(self._h, -1 if epilogue is None else int(epilogue), int(N), int(K),
                                                  C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)))
            return ms.value, fl.value, by.value, n.value
        if epilogue is None:
            L.check(self.lib.mmcm_gemm_time(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
            return ms.value, fl.value, n.value
        L.check(self.lib.mmcm_gemm_time_epi(self._h, int(epilogue), C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)))
        return ms.value, fl.value, by.value, n.value

"""Throughput of mmcm_resize_crop_u8 on a batch of decoded photos.  Usage: python tools/resize_bench.py [B] [H] [W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
import torch  # noqa: E402
from mmcm_b200 import prepost  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 375
W = int(sys.argv[3]) if len(sys.argv) > 3 else 500
imgs = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(B)]
for _ in range(3):
    out = prepost.resize_crop_u8(imgs, 224)
torch.cuda.synchronize()
# time the kernel alone: call the C entry on a pre-concatenated buffer
import ctypes as C  # noqa: E402
from mmcm_b200 import lib as L  # noqa: E402
flat = torch.cat([i.reshape(-1) for i in imgs])
offs = (C.c_int64 * B)(*[i * H * W * 3 for i in range(B)])
hs = (C.c_int32 * B)(*[H] * B)
ws = (C.c_int32 * B)(*[W] * B)
out = torch.empty((B, 224, 224, 3), dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib = L.load()
for _ in range(3):
    L.check(lib.mmcm_resize_crop_u8(flat.data_ptr(), offs, hs, ws, B, 224, out.data_ptr(), st))
torch.cuda.synchronize()
e0.record()
N = 20
for _ in range(N):
    L.check(lib.mmcm_resize_crop_u8(flat.data_ptr(), offs, hs, ws, B, 224, out.data_ptr(), st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(f"resize_crop_u8: {B} images {H}x{W} -> 224x224: {ms:.3f} ms/batch = {B / ms * 1e3:,.0f} images/s, "
      f"{(flat.numel() + out.numel()) / ms / 1e6:.0f} GB/s of source + crop bytes")

"""Host-buffer end-to-end throughput vs the H2D pipeline stage size (`host_chunk`).  Dev tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
import torch  # noqa: E402
import mmcm_b200 as P  # noqa: E402
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

a = A.CLIP_B32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
u8 = len(sys.argv) > 2 and sys.argv[2] == "u8"
m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
m.load_state_dict(syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0))
m = m.to("cuda:0").eval()
m.set_option("varlen_text", 0)
eng = m._ensure_engine(0)
host = {k: v.pin_memory() for k, v in syn.make_inputs(a, B, seed=1).items()}
img = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8).pin_memory()
mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
out = torch.empty((B, 5)).pin_memory()


def call():
    if u8:
        eng.forward_host_u8(host["input_ids"], host["attention_mask"], img, mean, std, host["text_present"],
                            host["image_present"], out=out)
    else:
        eng.forward_host(host["input_ids"], host["attention_mask"], host["pixel_values"], host["text_present"],
                         host["image_present"], out=out)


chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [128, 192, 256, 342, 512, 1024]
for chunk in chunks:
    m.set_option("host_chunk", chunk)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"host_chunk {chunk:5d}: {dt * 1e3:7.2f} ms/step  {B / dt:9.0f} samples/s")

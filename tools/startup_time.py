"""Start-up cost of the two weight paths: reference-style state dict (fp32 H2D + repack kernels per tensor) vs the
packed weight file (mmcm_load_packed: one mmap + copies).  Usage: python tools/startup_time.py [model]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
import torch  # noqa: E402
import mmcm_b200 as P  # noqa: E402
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402


def main():
    a = A.CLIP_B32
    spec = A.fusion_spec(a, 5, 512)
    sd = syn.make_state_dict(spec, a, seed=0)
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    t0 = time.perf_counter()
    m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
    t1 = time.perf_counter()
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    eng = m._ensure_engine(0)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    path = os.path.join(tempfile.mkdtemp(), "packed.bin")
    m.save_packed(path)
    t3 = time.perf_counter()
    s = P.PackedScorer(path, "cuda:0")
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    s2 = P.PackedScorer(path, "cuda:0")          # page cache warm
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    batch = {k: v.cuda() for k, v in syn.make_inputs(a, 8, seed=1).items()}
    assert torch.equal(m(**batch)["logits"], s2(**batch)["logits"])
    print(f"module construction (random init)      {t1 - t0:7.2f} s")
    print(f"load_state_dict + .to(cuda) + repack   {t2 - t1:7.2f} s")
    print(f"save_packed ({os.path.getsize(path) / 1e6:.0f} MB)                  {t3 - t2:7.2f} s")
    print(f"PackedScorer from file (first)         {t4 - t3:7.2f} s")
    print(f"PackedScorer from file (second)        {t5 - t4:7.2f} s")


if __name__ == "__main__":
    main()

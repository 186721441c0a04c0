"""CPU checks for the pre-/post-processing rows: the oracle against the libraries the reference itself calls
(torchvision transforms, sklearn metrics), and the host-side metric arithmetic.  No GPU needed."""
import numpy as np
import pytest
import torch

from oracle import prepost_oracle as orc


def test_oracle_matches_torchvision_eval_transform_tail():
    """R/src/data/dataset.py:106-111: ToTensor + Normalize applied to an already 224x224 RGB image."""
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8)
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    tf = T.Compose([T.ToTensor(), T.Normalize(mean, std)])
    ref = tf(Image.fromarray(arr))
    got = orc.to_tensor_normalize(torch.from_numpy(arr)[None], mean, std)[0]
    assert torch.equal(got, ref)


def test_metrics_from_confusion_matches_sklearn():
    from mmcm_b200 import prepost
    rng = np.random.default_rng(1)
    N, C = 999, 5
    probs = rng.random((N, C)).astype(np.float32)
    y = (rng.random((N, C)) < 0.25).astype(np.float32)
    y[:, 4] = 0                                       # a class without positives: zero_division=0 paths
    thr = np.array([0.2, 0.35, 0.5, 0.8, 0.95], dtype=np.float32)
    pred = probs >= thr[None]
    conf = torch.zeros(C, 4, dtype=torch.int64)
    for c in range(C):
        conf[c, 0] = int((pred[:, c] & (y[:, c] > 0.5)).sum())
        conf[c, 1] = int((pred[:, c] & (y[:, c] < 0.5)).sum())
        conf[c, 2] = int((~pred[:, c] & (y[:, c] > 0.5)).sum())
        conf[c, 3] = int((~pred[:, c] & (y[:, c] < 0.5)).sum())
    m = prepost.metrics_from_confusion(conf)
    ref = orc.detailed_metrics(probs, y, thr)
    for k in ("f1_macro", "f1_micro", "precision_macro", "recall_macro"):
        assert abs(m[k] - ref[k]) < 1e-12, k
    np.testing.assert_allclose(m["per_class"]["f1"], ref["per_class_f1"], atol=1e-12)
    assert m["per_class"]["support"][4] == 0


def test_postprocess_requires_gpu():
    from mmcm_b200 import prepost
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        prepost.postprocess(torch.zeros(2, 5), torch.full((5,), 0.5))
    with pytest.raises(ValueError):
        prepost.preprocess_u8(torch.zeros(1, 224, 224, 3, dtype=torch.uint8), [0.5] * 3, [0.5] * 3)


def test_resize_oracle_is_pinned_against_pillow():
    """oracle.prepost_oracle.resize_center_crop (numpy restatement of Pillow's Resample.c + torchvision's size / crop
    arithmetic) against Pillow + torchvision themselves, the libraries the reference's eval_tf calls."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms as T
    from oracle import prepost_oracle as orc
    rng = np.random.default_rng(5)
    tf = T.Compose([T.Resize(224, antialias=True), T.CenterCrop((224, 224))])
    for h, w in [(375, 500), (500, 375), (224, 224), (224, 300), (100, 80), (231, 517), (225, 224), (449, 449), (7, 9)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(orc.resize_center_crop(img, 224), np.asarray(tf(Image.fromarray(img)))), (h, w)
    assert orc.center_crop_origin(299, 225, 224) == (38, 0) and orc.center_crop_origin(297, 224, 224) == (36, 0)


def test_resize_oracle_random_geometries_against_pillow():
    """Seeded sweep over awkward geometries (1-pixel sides, primes, near-square, 10x up / down scaling, both crop
    parities) for two crop sizes: the restatement must equal Pillow + torchvision everywhere."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms as T
    from oracle import prepost_oracle as orc
    rng = np.random.default_rng(2026)
    shapes = [(1, 1), (1, 37), (41, 1), (2, 3), (31, 33), (64, 64), (65, 64), (64, 67), (97, 211), (640, 61), (59, 600)]
    shapes += [tuple(int(v) for v in rng.integers(3, 300, 2)) for _ in range(25)]
    for size in (32, 56):
        tf = T.Compose([T.Resize(size, antialias=True), T.CenterCrop((size, size))])
        for h, w in shapes:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            ref = np.asarray(tf(Image.fromarray(img)))
            assert np.array_equal(orc.resize_center_crop(img, size), ref), (h, w, size)

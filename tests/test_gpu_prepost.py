"""uint8 pre-processing and fused post-processing against the oracle (the reference's literal
numpy / sklearn expressions)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_preprocess_u8_is_bit_identical_to_totensor_normalize():
    from mmcm_b200 import prepost
    from oracle import prepost_oracle as orc
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (5, 224, 224, 3), generator=g, dtype=torch.uint8)
    for mean, std in (([0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]),   # CLIP processor
                      ([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])):                                             # SigLIP / default
        ref = orc.to_tensor_normalize(img, mean, std)
        got = prepost.preprocess_u8(img.cuda(), mean, std).cpu()
        assert got.shape == (5, 3, 224, 224)
        assert torch.equal(got, ref)


def test_postprocess_matches_numpy_and_sklearn():
    from mmcm_b200 import prepost
    from oracle import prepost_oracle as orc
    g = torch.Generator().manual_seed(1)
    B, C = 4099, 5
    logits = torch.randn(B, C, generator=g) * 3
    thr = torch.tensor([0.2, 0.35, 0.5, 0.8, 0.95])            # R/runs/*/inference_config.json range
    y = (torch.rand(B, C, generator=g) < 0.3).float()
    p_ref, l_ref, a_ref = orc.postprocess(logits.numpy(), thr.numpy())
    out = prepost.postprocess(logits.cuda(), thr, labels=y)
    probs = out["probs"].cpu().numpy()
    np.testing.assert_allclose(probs, p_ref, rtol=0, atol=2e-7)
    safe = np.abs(p_ref - thr.numpy()[None]) > 1e-6              # decisions are exact away from a 1-ulp band
    assert (out["labels"].cpu().numpy() == l_ref)[safe].all()
    rows_safe = safe.all(axis=1)
    assert (out["any_harmful"].cpu().numpy() == a_ref)[rows_safe].all()
    conf = out["confusion"]
    assert int(conf.sum()) == B * C
    # accumulate a second batch into the same counters, then compare the derived metrics with sklearn on both
    out2 = prepost.postprocess(logits.flip(0).cuda(), thr, labels=y.flip(0), confusion=conf)
    m = prepost.metrics_from_confusion(out2["confusion"])
    ref = orc.detailed_metrics(np.concatenate([p_ref, p_ref[::-1]]), np.concatenate([y.numpy(), y.numpy()[::-1]]),
                               thr.numpy())
    for k in ("f1_macro", "f1_micro", "precision_macro", "recall_macro"):
        assert abs(m[k] - ref[k]) < 1e-9, k
    np.testing.assert_allclose(m["per_class"]["f1"], ref["per_class_f1"], atol=1e-9)


def test_u8_pipeline_feeds_the_forward():
    """uint8 images -> preprocess_u8 -> model == fp32 pixel_values computed the reference's way -> model."""
    from conftest import build_case
    from test_gpu_forward import _make_module
    from mmcm_b200 import prepost, synthetic as syn
    from oracle import prepost_oracle as orc
    kind, a, kw, sd, _, _ = build_case("clip_fusion_hardened")
    m = _make_module(kind, a, kw, sd)
    batch = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, 8, seed=50).items()}
    img = torch.randint(0, 256, (8, 224, 224, 3), generator=torch.Generator().manual_seed(2), dtype=torch.uint8)
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    b_ref = dict(batch, pixel_values=orc.to_tensor_normalize(img, mean, std).cuda())
    b_u8 = dict(batch, pixel_values=prepost.preprocess_u8(img.cuda(), mean, std))
    assert torch.equal(m(**b_ref)["logits"], m(**b_u8)["logits"])


def test_batched_scorer_equals_per_sample_reference_flow():
    """One batched call == the reference's per-request loop (tokenised ids + uint8 crops in, records out)."""
    import numpy as np
    import mmcm_b200 as P
    from conftest import build_case
    from test_gpu_forward import _make_module
    from mmcm_b200 import synthetic as syn
    from oracle import prepost_oracle as orc
    from oracle import scoring_oracle as so
    kind, a, kw, sd, _, _ = build_case("clip_fusion_hardened")
    m = _make_module(kind, a, kw, sd)
    N = 21
    b = syn.make_inputs(a, N, seed=80)
    img = torch.randint(0, 256, (N, 224, 224, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    names = ["racist", "sexist", "homophobe", "religion", "otherhate"]
    thr = [0.2, 0.35, 0.5, 0.8, 0.95]
    scorer = P.BatchedScorer(m, names, thr, max_batch=8)
    res = scorer.score(b["input_ids"], b["attention_mask"], img)
    recs = scorer.as_records(res)
    assert len(recs) == N and set(recs[0]["predictions"]) == set(names)
    # oracle: reference transform tail -> reference forward -> reference post-processing, on the CPU
    mean, std = scorer.mean, scorer.std
    batch = dict(b, pixel_values=orc.to_tensor_normalize(img, mean, std), image_present=torch.ones(N))
    with torch.no_grad():
        ref_logits = so.fusion_forward(sd, batch, "clip", a.patch, a.eos_id)
    p_ref, l_ref, a_ref = orc.postprocess(ref_logits.numpy(), np.array(thr, dtype=np.float32))
    probs = res["probs"].cpu().numpy()
    assert np.abs(probs - p_ref).max() <= 0.25 * 0.05 * float(ref_logits.std())      # sigmoid-Lipschitz of the logit gate
    far = np.abs(p_ref - np.array(thr)[None]) > 0.25 * 0.05 * float(ref_logits.std())
    assert (res["labels"].cpu().numpy() == l_ref)[far].all()
    # no image at all -> image_present = 0 path
    res2 = scorer.score(b["input_ids"][:4], b["attention_mask"][:4], None)
    assert res2["probs"].shape == (4, 5)


@pytest.mark.parametrize("name", ["clip_fusion_hardened", "siglip_fusion_hardened", "clip_mtl_h256_hardened"])
def test_forward_u8_is_bit_identical_to_transform_then_forward(name):
    """mmcm_forward_u8 / mmcm_forward_host_u8 (ToTensor + Normalize inside the im2col) == the reference's transform tail
    (R/src/data/dataset.py:106-111, computed by torch on the CPU) followed by the fp32-pixel forward -- every bit."""
    from conftest import build_case
    from test_gpu_forward import _make_module
    from mmcm_b200 import synthetic as syn
    from oracle import prepost_oracle as orc
    kind, a, kw, sd, _, _ = build_case(name)
    m = _make_module(kind, a, kw, sd)
    B = 11
    b = syn.make_inputs(a, B, seed=70, edge_rows=True)
    img = torch.randint(0, 256, (B, a.image, a.image, 3), generator=torch.Generator().manual_seed(4), dtype=torch.uint8)
    img[0] = 0
    img[1] = 255
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    dbatch = {k: v.to("cuda:0") for k, v in b.items()}
    ref = m(**dict(dbatch, pixel_values=orc.to_tensor_normalize(img, mean, std).cuda()))["logits"]
    got, probs = m.forward_u8(dbatch["input_ids"], dbatch["attention_mask"], img.cuda(), dbatch["text_present"],
                              dbatch["image_present"], mean, std, want_probs=True)
    assert torch.equal(got, ref)
    assert torch.equal(probs, torch.sigmoid(ref)) or (probs - torch.sigmoid(ref)).abs().max().item() < 1e-6
    # a pixel buffer that is not 8-byte aligned takes the byte-load branch of the kernel
    flat = torch.empty(img.numel() + 3, dtype=torch.uint8, device="cuda:0")
    odd = flat[3:].view_as(img)
    odd.copy_(img)
    assert odd.data_ptr() % 8 != 0
    assert torch.equal(m.forward_u8(dbatch["input_ids"], dbatch["attention_mask"], odd, dbatch["text_present"],
                                    dbatch["image_present"], mean, std), ref)
    # host entry point, chunked H2D of the uint8 images
    eng = m._engine
    m.set_option("micro_batch", 4)
    host = eng.forward_host_u8(b["input_ids"], b["attention_mask"], img.pin_memory(), mean, std, b["text_present"],
                               b["image_present"])
    assert torch.equal(host, ref.cpu())
    # argument errors: wrong size mirrors HF clip :204-207, wrong dtype / std
    with pytest.raises(ValueError, match="doesn't match model"):
        m.forward_u8(dbatch["input_ids"], dbatch["attention_mask"], img[:, :32].cuda(), dbatch["text_present"],
                     dbatch["image_present"], mean, std)
    with pytest.raises(ValueError):
        m.forward_u8(dbatch["input_ids"], dbatch["attention_mask"], img.cuda().float(), dbatch["text_present"],
                     dbatch["image_present"], mean, std)
    with pytest.raises(ValueError, match="std"):
        m.forward_u8(dbatch["input_ids"], dbatch["attention_mask"], img.cuda(), dbatch["text_present"],
                     dbatch["image_present"], mean, [0.0, 1.0, 1.0])


RESIZE_SHAPES = [(375, 500), (500, 375), (224, 224), (224, 300), (1080, 1920), (100, 80), (231, 517), (640, 427),
                 (225, 224), (3000, 223), (449, 449), (450, 675), (7, 9), (2048, 1365)]


def test_resize_crop_u8_is_byte_identical_to_pillow():
    """mmcm_resize_crop_u8 == T.CenterCrop(224)(T.Resize(224, antialias=True)(PIL image)) (dataset.py:106-108) for a
    ragged batch: down-, up- and no scaling, extreme aspect ratios, odd crop offsets (round-half-even)."""
    from PIL import Image
    from torchvision import transforms as T
    from mmcm_b200 import prepost
    from oracle import prepost_oracle as orc
    rng = np.random.default_rng(11)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in RESIZE_SHAPES]
    imgs[0][:] = 255                                            # saturated image: rounding must not overflow 255
    imgs[1][::2] = 0                                            # high-frequency stripes
    tf = T.Compose([T.Resize(224, antialias=True), T.CenterCrop((224, 224))])
    got = prepost.resize_crop_u8([torch.from_numpy(i) for i in imgs], 224, device="cuda:0").cpu().numpy()
    for i, im in enumerate(imgs):
        ref = np.asarray(tf(Image.fromarray(im)))
        assert np.array_equal(got[i], ref), f"image {i} {im.shape}: max diff {np.abs(got[i].astype(int) - ref).max()}"
        if im.shape[0] * im.shape[1] < 400 * 600:               # the numpy restatement is slow on the big ones
            assert np.array_equal(orc.resize_center_crop(im, 224), ref)
    # another crop size, CUDA inputs
    tf96 = T.Compose([T.Resize(96, antialias=True), T.CenterCrop((96, 96))])
    got96 = prepost.resize_crop_u8([torch.from_numpy(i).cuda() for i in imgs[:6]], 96).cpu().numpy()
    for i in range(6):
        assert np.array_equal(got96[i], np.asarray(tf96(Image.fromarray(imgs[i]))))


def test_resize_crop_u8_awkward_geometries():
    """1-pixel sides, primes, 10x up / down scaling, both crop parities, two crop sizes -- against Pillow."""
    from PIL import Image
    from torchvision import transforms as T
    from mmcm_b200 import prepost
    rng = np.random.default_rng(2026)
    shapes = [(1, 1), (1, 37), (41, 1), (2, 3), (31, 33), (64, 64), (65, 64), (64, 67), (97, 211), (640, 61), (59, 600)]
    shapes += [tuple(int(v) for v in rng.integers(3, 300, 2)) for _ in range(40)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    for size in (32, 56, 224):
        tf = T.Compose([T.Resize(size, antialias=True), T.CenterCrop((size, size))])
        got = prepost.resize_crop_u8([torch.from_numpy(i) for i in imgs], size, device="cuda:0").cpu().numpy()
        for i, im in enumerate(imgs):
            assert np.array_equal(got[i], np.asarray(tf(Image.fromarray(im)))), (im.shape, size)


def test_raw_images_to_logits_equals_the_reference_transform_then_forward():
    """decoded images -> resize_crop_u8 -> forward_u8  ==  eval_tf on the CPU (PIL + torchvision) -> forward."""
    from PIL import Image
    from torchvision import transforms as T
    from conftest import build_case
    from test_gpu_forward import _make_module
    from mmcm_b200 import prepost, synthetic as syn
    kind, a, kw, sd, _, _ = build_case("clip_fusion_hardened")
    m = _make_module(kind, a, kw, sd)
    rng = np.random.default_rng(12)
    shapes = RESIZE_SHAPES[:8]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    eval_tf = T.Compose([T.Resize(224, antialias=True), T.CenterCrop((224, 224)), T.ToTensor(), T.Normalize(mean, std)])
    px = torch.stack([eval_tf(Image.fromarray(i)) for i in imgs])
    b = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, len(imgs), seed=71).items()}
    ref = m(**dict(b, pixel_values=px.cuda()))["logits"]
    crops = prepost.resize_crop_u8([torch.from_numpy(i) for i in imgs], 224, device="cuda:0")
    got = m.forward_u8(b["input_ids"], b["attention_mask"], crops, b["text_present"], b["image_present"], mean, std)
    assert torch.equal(got, ref)
    # the same through BatchedScorer with a ragged list (two internal batches)
    import mmcm_b200 as P
    thr = [0.2, 0.35, 0.5, 0.8, 0.95]
    scorer = P.BatchedScorer(m, ["racist", "sexist", "homophobe", "religion", "otherhate"], thr, mean, std, max_batch=5)
    res = scorer.score(b["input_ids"].cpu(), b["attention_mask"].cpu(), [torch.from_numpy(i) for i in imgs])
    assert torch.equal(res["probs"], prepost.postprocess(ref, torch.tensor(thr))["probs"])


def test_resize_crop_u8_argument_errors():
    from mmcm_b200 import prepost
    with pytest.raises(ValueError):
        prepost.resize_crop_u8([], 224)
    with pytest.raises(ValueError):
        prepost.resize_crop_u8([torch.zeros(4, 4, 3)], 224)
    with pytest.raises(ValueError, match="bad geometry"):
        prepost.resize_crop_u8([torch.zeros(4, 4, 3, dtype=torch.uint8), torch.zeros(0, 5, 3, dtype=torch.uint8)], 8,
                               device="cuda:0")
    with pytest.raises(ValueError, match="size"):
        prepost.resize_crop_u8([torch.zeros(4, 4, 3, dtype=torch.uint8)], 4096, device="cuda:0")

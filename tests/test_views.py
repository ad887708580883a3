"""SURVEY 8 f4, data side: flip + reflect-padded crop + ToTensor + Normalize of the reference's view transforms
(code/dataset.py:24-53) on the device, against torchvision (the library the reference calls) on the same seeded images."""
import numpy as np
import pytest
import torch


def _torchvision_views(imgs_u8, size, padding, seed, crop=True):
    """The reference's Compose, image by image, on the host: RandomHorizontalFlip -> RandomCrop(reflect) -> ToTensor -> Normalize."""
    from PIL import Image
    from torchvision import transforms
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    ops = [transforms.RandomHorizontalFlip()]
    if crop:
        ops.append(transforms.RandomCrop(size=size, padding=padding, padding_mode="reflect"))
    tf = transforms.Compose(ops + [transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])
    torch.manual_seed(seed)
    return torch.stack([tf(Image.fromarray(im)) for im in imgs_u8])


def test_draw_view_params_follow_torchvision_rng_order():
    from endoscopy_image_classification_b200.views import draw_view_params
    from torchvision import transforms
    n, H, W, size, pad = 6, 20, 24, 16, 3
    torch.manual_seed(11)
    flips, xy = draw_view_params(n, H, W, size, pad)
    torch.manual_seed(11)
    want_f, want_xy = [], []
    for _ in range(n):
        want_f.append(int(torch.rand(1) < 0.5))
        i, j, _, _ = transforms.RandomCrop.get_params(torch.zeros(3, H + 2 * pad, W + 2 * pad), (size, size))
        want_xy.append((j, i))
    assert flips.tolist() == want_f and xy.tolist() == [list(t) for t in want_xy]


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,size,pad,crop", [(32, 32, 32, 4, True), (224, 224, 224, 28, True), (40, 56, 32, 0, True), (64, 64, 64, 0, False)])
def test_normalize_views_bit_exact_with_torchvision(H, W, size, pad, crop):
    from endoscopy_image_classification_b200.views import draw_view_params, normalize_views
    n = 5
    rng = np.random.default_rng(H + W + pad)
    imgs = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    ref = _torchvision_views(imgs, size, pad, seed=3, crop=crop)
    torch.manual_seed(3)
    flips, xy = draw_view_params(n, H, W, size, pad, crop=crop)
    got = normalize_views(torch.from_numpy(imgs).cuda(), size=size, padding=pad, flip=flips.cuda(),
                          crop_xy=xy.cuda() if xy is not None else None)
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)          # fp32: every rounding like ToTensor / Normalize
    got16 = normalize_views(torch.from_numpy(imgs).cuda(), size=size, padding=pad, flip=flips.cuda(),
                            crop_xy=xy.cuda() if xy is not None else None, dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), ref.to(torch.bfloat16))
    # no flip / centred window defaults: plain ToTensor + Normalize
    if not crop:
        from torchvision import transforms
        plain = torch.stack([transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(transforms.ToTensor()(im)) for im in imgs])
        assert torch.equal(normalize_views(torch.from_numpy(imgs).cuda()).cpu(), plain)
    with pytest.raises(ValueError):
        normalize_views(torch.zeros(2, 3, 8, 8, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        normalize_views(torch.zeros(2, 8, 8, 3, dtype=torch.uint8))           # CPU tensor: no CPU path

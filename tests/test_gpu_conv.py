"""Plane convolutions on the tcgen05 pipeline vs ATen convolutions evaluated in fp64."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _rel(a, ref):
    return (a.double() - ref).abs().max().item() / ref.abs().max().item()


@pytest.mark.parametrize("B,cin,cout,H,W", [(1, 32, 32, 8, 16), (2, 32, 64, 32, 32), (1, 64, 128, 64, 48), (2, 128, 64, 16, 16),
                                           (1, 256, 256, 32, 32), (1, 512, 256, 16, 32)])
@pytest.mark.parametrize("relu_in", [False, True])
@pytest.mark.parametrize("flavour", ["f16", "tf32"])
def test_conv3x3_matches_fp64(B, cin, cout, H, W, relu_in, flavour, monkeypatch):
    from tomosar2height_b200.conv import conv3x3
    from tomosar2height_b200 import linear as L
    monkeypatch.setattr(L, "USE_F16", flavour == "f16")        # 3xFP16 (default) or 3xTF32 entry points
    monkeypatch.setattr(L, "USE_F16_WGRAD", flavour == "f16")
    g = torch.Generator().manual_seed(cin + cout + H)
    x = _cl(torch.randn(B, cin, H, W, generator=g).cuda()).requires_grad_(True)
    w = _cl((torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()).requires_grad_(True)
    b = torch.randn(cout, generator=g).cuda().requires_grad_(True)
    gy = _cl(torch.randn(B, cout, H, W, generator=g).cuda())
    from tomosar2height_b200 import _lib
    before = _lib.launch_count
    y = conv3x3(x, w, b, relu_in=relu_in)
    assert _lib.launch_count > before, "expected the tensor-core path"
    y.backward(gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = F.conv2d(xd.relu() if relu_in else xd, wd, bd, padding=1)
    yd.backward(gy.double())
    tol = 1e-5 * max(1.0, 9 * cin / 256)
    assert _rel(y, yd) < tol, _rel(y, yd)
    assert _rel(x.grad, xd.grad) < tol, _rel(x.grad, xd.grad)
    assert _rel(w.grad, wd.grad) < tol * 4, _rel(w.grad, wd.grad)
    assert _rel(b.grad, bd.grad) < 1e-5


def test_conv1x1_and_transpose_match_fp64():
    from tomosar2height_b200.conv import conv1x1, conv_transpose2x2
    g = torch.Generator().manual_seed(3)
    x = _cl(torch.randn(2, 64, 16, 16, generator=g).cuda()).requires_grad_(True)
    w1 = (torch.randn(32, 64, 1, 1, generator=g) / 8).cuda().requires_grad_(True)
    b1 = torch.randn(32, generator=g).cuda().requires_grad_(True)
    wt = (torch.randn(64, 32, 2, 2, generator=g) / 8).cuda().requires_grad_(True)
    bt = torch.randn(32, generator=g).cuda().requires_grad_(True)
    y1 = conv1x1(x, w1, b1)
    yt = conv_transpose2x2(x, wt, bt)
    (y1.square().sum() + yt.square().sum()).backward()
    xd, w1d, b1d, wtd, btd = (t.detach().double().requires_grad_(True) for t in (x, w1, b1, wt, bt))
    y1d = F.conv2d(xd, w1d, b1d)
    ytd = F.conv_transpose2d(xd, wtd, btd, stride=2)
    (y1d.square().sum() + ytd.square().sum()).backward()
    for got, ref in ((y1, y1d), (yt, ytd), (x.grad, xd.grad), (w1.grad, w1d.grad), (b1.grad, b1d.grad),
                     (wt.grad, wtd.grad), (bt.grad, btd.grad)):
        assert got.shape == ref.shape and _rel(got, ref) < 2e-5, _rel(got, ref)


def test_conv_fallback_for_uncovered_shapes():
    from tomosar2height_b200.conv import conv3x3
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3, 20, 20, generator=g).cuda()       # image-branch style: 3 input channels
    w = torch.randn(8, 3, 3, 3, generator=g).cuda()
    assert torch.allclose(conv3x3(x, w), F.conv2d(x, w, padding=1), rtol=1e-4, atol=1e-5)

"""One-call block entry points of the C ABI (t2h_resblock_fwd/bwd, t2h_comm_mlp_fwd/bwd) against the
oracle's restatement of the reference blocks (block/resnet.py:46-54, encoder/alto.py:63-69,123-128) in fp64."""
import pytest
import torch

import oracle
from oracle import model as omodel

pytestmark = pytest.mark.gpu

TOL = 1e-5  # fp32-grade: 3xTF32 / 3xFP16 GEMMs against an fp64 evaluation, relative to max |reference|


def _rel(a, ref):
    return (a.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def _workspace(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("rows,k1,k2,n_h,n_out,shortcut", [
    (5000, 32, 32, 32, 32, True),     # the encoder's blocks: [net | pooled] -> 32 (pointnet.py:73-79)
    (777, 64, 0, 32, 32, True),       # first block, one source (pointnet.py:37-39)
    (3000, 32, 0, 32, 32, False),     # identity shortcut (size_in == size_out)
    (2500, 128, 128, 128, 256, True),  # wide: the 3xFP16 flavour, CTA-pair kernels
    (1, 32, 32, 32, 32, True),
])
def test_resblock_entry_points_match_the_reference_block(rows, k1, k2, n_h, n_out, shortcut):
    from tomosar2height_b200 import _lib
    ptr, lib = _lib.ptr, _lib.load()
    g = torch.Generator().manual_seed(rows + k1 + n_out)
    n_in = k1 + k2
    x = torch.randn(rows, n_in, generator=g)
    P = {"blk.fc_0.weight": torch.randn(n_h, n_in, generator=g) / n_in ** 0.5, "blk.fc_0.bias": torch.randn(n_h, generator=g),
         "blk.fc_1.weight": torch.randn(n_out, n_h, generator=g) / n_h ** 0.5, "blk.fc_1.bias": torch.randn(n_out, generator=g)}
    if shortcut:
        P["blk.shortcut.weight"] = torch.randn(n_out, n_in, generator=g) / n_in ** 0.5
    gy = torch.randn(rows, n_out, generator=g)

    # the oracle's block (fp64) and its gradients
    Pd = {k: v.double().requires_grad_(True) for k, v in P.items()}
    xd = x.double().requires_grad_(True)
    ref = omodel._resblock(Pd, "blk", xd)
    ref.backward(gy.double())
    net_ref = torch.relu(xd.detach()) @ Pd["blk.fc_0.weight"].detach().t() + Pd["blk.fc_0.bias"].detach()

    dev = {k: v.cuda() for k, v in P.items()}
    xc = x.cuda()
    x1, x2 = xc[:, :k1], (xc[:, k1:] if k2 else None)     # column views of one buffer: pitches differ from widths
    ws_bytes = lib.t2h_resblock_workspace_bytes(rows, k1, k2, n_h, n_out, int(shortcut))
    ws = _workspace(ws_bytes)
    net = torch.empty(rows, n_h, device="cuda")
    out = torch.empty(rows, n_out, device="cuda")
    wsc = dev.get("blk.shortcut.weight")
    _lib.call("t2h_resblock_fwd", ptr(x1), x1.stride(0), k1, ptr(x2), 0 if x2 is None else x2.stride(0), k2, rows,
              ptr(dev["blk.fc_0.weight"]), ptr(dev["blk.fc_0.bias"]), n_h, ptr(dev["blk.fc_1.weight"]),
              ptr(dev["blk.fc_1.bias"]), ptr(wsc), n_out, ptr(ws), ws_bytes, ptr(net), net.stride(0), ptr(out), out.stride(0))
    assert _rel(out, ref.detach()) < TOL
    assert _rel(net, net_ref) < TOL

    gyc = gy.cuda()
    dx = torch.empty(rows, n_in, device="cuda")
    d_x1, d_x2 = dx[:, :k1], (dx[:, k1:] if k2 else None)
    d_w0, d_b0 = torch.empty(n_h, n_in, device="cuda"), torch.empty(n_h, device="cuda")
    d_w1, d_b1 = torch.empty(n_out, n_h, device="cuda"), torch.empty(n_out, device="cuda")
    d_ws = torch.empty(n_out, n_in, device="cuda") if shortcut else None
    _lib.call("t2h_resblock_bwd", ptr(gyc), gyc.stride(0), ptr(x1), x1.stride(0), k1, ptr(x2), 0 if x2 is None else x2.stride(0),
              k2, ptr(net), net.stride(0), rows, ptr(dev["blk.fc_0.weight"]), n_h, ptr(dev["blk.fc_1.weight"]), ptr(wsc),
              n_out, ptr(ws), ws_bytes, ptr(d_x1), d_x1.stride(0), ptr(d_x2), 0 if d_x2 is None else d_x2.stride(0),
              ptr(d_w0), ptr(d_b0), ptr(d_w1), ptr(d_b1), ptr(d_ws))
    assert _rel(dx, xd.grad) < TOL
    assert _rel(d_w0, Pd["blk.fc_0.weight"].grad) < TOL and _rel(d_b0, Pd["blk.fc_0.bias"].grad) < TOL
    assert _rel(d_w1, Pd["blk.fc_1.weight"].grad) < TOL and _rel(d_b1, Pd["blk.fc_1.bias"].grad) < TOL
    if shortcut:
        assert _rel(d_ws, Pd["blk.shortcut.weight"].grad) < TOL


@pytest.mark.parametrize("rows,C,C_prev", [(4000, 32, 0), (3000, 64, 32), (2600, 128, 64), (2000, 256, 128), (1300, 512, 256)])
def test_comm_mlp_entry_points_match_the_reference_block(rows, C, C_prev):
    from tomosar2height_b200 import _lib
    ptr, lib = _lib.ptr, _lib.load()
    g = torch.Generator().manual_seed(rows + C)
    c = torch.randn(rows, C, generator=g)
    c_last = torch.randn(rows, C_prev, generator=g) if C_prev else None
    W = {"w0": torch.randn(2 * C, C, generator=g) / C ** 0.5, "b0": torch.randn(2 * C, generator=g),
         "w2": torch.randn(C, 2 * C, generator=g) / (2 * C) ** 0.5, "b2": torch.randn(C, generator=g)}
    if C_prev:
        W["wc"] = torch.randn(C, C_prev, generator=g) / C_prev ** 0.5
        W["bc"] = torch.randn(C, generator=g)
    gy = torch.randn(rows, C, generator=g)

    # alto.py:123-128: c = fc_comm(c); c = c + fc_c(c_last)   (fc_comm = Linear, ReLU, Linear: alto.py:63-67)
    Wd = {k: v.double().requires_grad_(True) for k, v in W.items()}
    cd = c.double().requires_grad_(True)
    cld = c_last.double().requires_grad_(True) if C_prev else None
    hid_ref = cd @ Wd["w0"].t() + Wd["b0"]
    ref = torch.relu(hid_ref) @ Wd["w2"].t() + Wd["b2"]
    if C_prev:
        ref = ref + cld @ Wd["wc"].t() + Wd["bc"]
    ref.backward(gy.double())

    dev = {k: v.cuda() for k, v in W.items()}
    cc, clc, gyc = c.cuda(), (c_last.cuda() if C_prev else None), gy.cuda()
    ws_bytes = lib.t2h_comm_mlp_workspace_bytes(rows, C, C_prev)
    ws = _workspace(ws_bytes)
    hidden = torch.empty(rows, 2 * C, device="cuda")
    out = torch.empty(rows, C, device="cuda")
    _lib.call("t2h_comm_mlp_fwd", ptr(cc), cc.stride(0), C, ptr(clc), C_prev, C_prev, rows, ptr(dev["w0"]), ptr(dev["b0"]),
              ptr(dev["w2"]), ptr(dev["b2"]), ptr(dev.get("wc")), ptr(dev.get("bc")), ptr(ws), ws_bytes, ptr(hidden),
              hidden.stride(0), ptr(out), out.stride(0))
    assert _rel(out, ref.detach()) < TOL
    assert _rel(hidden, hid_ref.detach()) < TOL

    d_c = torch.empty_like(cc)
    d_cl = torch.empty_like(clc) if C_prev else None
    grads = {k: torch.empty_like(v) for k, v in dev.items()}
    _lib.call("t2h_comm_mlp_bwd", ptr(gyc), gyc.stride(0), ptr(cc), cc.stride(0), C, ptr(clc), C_prev, C_prev, ptr(hidden),
              hidden.stride(0), rows, ptr(dev["w0"]), ptr(dev["w2"]), ptr(dev.get("wc")), ptr(ws), ws_bytes, ptr(d_c),
              d_c.stride(0), ptr(d_cl), C_prev, ptr(grads["w0"]), ptr(grads["b0"]), ptr(grads["w2"]), ptr(grads["b2"]),
              ptr(grads.get("wc")), ptr(grads.get("bc")))
    tol = TOL * max(1.0, C / 128)   # fp32 accumulation over 2C-long rows
    assert _rel(d_c, cd.grad) < tol
    for k in W:
        assert _rel(grads[k], Wd[k].grad) < tol, k
    if C_prev:
        assert _rel(d_cl, cld.grad) < tol


def test_block_entry_points_reject_bad_arguments():
    from tomosar2height_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(64, 30, device="cuda")
    ws = _workspace(1 << 20)
    p = _lib.ptr
    # a width that is not a multiple of 4, and a workspace that is too small
    assert lib.t2h_resblock_fwd(p(x), 30, 30, None, 0, 0, 64, p(x), None, 32, p(x), None, p(x), 32, p(ws), 1 << 20, p(x), 32, p(x), 32,
                                None) == 1
    x = torch.zeros(4096, 64, device="cuda")
    w = torch.zeros(64, 64, device="cuda")
    out = torch.zeros(4096, 64, device="cuda")
    assert lib.t2h_resblock_fwd(p(x), 64, 64, None, 0, 0, 4096, p(w), None, 64, p(w), None, p(w), 64, p(ws), 512, p(out), 64, p(out),
                                64, None) == 4

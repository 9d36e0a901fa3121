"""GPU parity of every C-ABI operator against the CPU oracle (same seeded inputs).

Bar (north star): bit-exact cell indices / argmax, <= 1e-4 relative for fp32 features and gradients.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from cases import synthetic_cloud

pytestmark = pytest.mark.gpu

REL = 1e-4  # north-star tolerance for fp32 values


def _close(got, want, rel=REL, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().clamp(min=1e-30)
    err = (got - want).abs().max()
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="module")
def T():
    import tomosar2height_b200.functional as F
    return F


def _topo(cloud, reso):
    from tomosar2height_b200.topology import Topology
    return Topology(cloud.cuda(), reso)


def test_golden_op_vectors(T, golden_dir):
    from tomosar2height_b200 import scatter as S
    from tomosar2height_b200.utils import coordinate2index
    v = np.load(os.path.join(golden_dir, "op_vectors.npz"))
    xy = torch.from_numpy(v["g1_xy"]).cuda()
    idx = coordinate2index(xy, 2)
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), v["g1_index"])
    plane = S.scatter_mean(xy.permute(0, 2, 1), idx, out=xy.new_zeros(1, 2, 4)).reshape(1, 2, 2, 2)
    np.testing.assert_allclose(plane.cpu().numpy(), v["g1_plane"], rtol=1e-6)
    out, arg = S.scatter_max(torch.from_numpy(v["g2_src"]).cuda(), torch.from_numpy(v["g2_index"]).cuda(), dim_size=5)
    assert np.array_equal(out.cpu().numpy(), v["g2_out"]) and np.array_equal(arg.cpu().numpy(), v["g2_arg"])
    edge = torch.from_numpy(v["g3_xy"]).cuda()
    for r in (256, 100):
        assert np.array_equal(coordinate2index(edge, r).cpu().numpy(), v[f"g3_index_{r}"])
    # G-4 through a topology
    p = torch.from_numpy(v["g4_p"]).cuda()
    topo = _topo(p.cpu(), 4)
    plane = torch.from_numpy(v["g4_plane"]).cuda().repeat(1, 4, 1, 1)  # C=4 copies
    rows = T.bilinear_sample(T.nchw_to_plane(plane), topo.level(4))
    got = topo.unsort_rows(rows)[:, 0].cpu().numpy()
    np.testing.assert_allclose(got, v["g4_out"].reshape(-1), rtol=1e-6, atol=1e-6)
    # G-5
    t = torch.from_numpy(v["g5_in"]).cuda().repeat(1, 2, 1, 1)  # C=4
    same = T.upsample_bilinear(T.nchw_to_plane(t), 6).permute(0, 3, 1, 2)[:, :2]
    assert np.array_equal(same.cpu().numpy(), v["g5_same"])
    up = T.upsample_bilinear(T.nchw_to_plane(t), 12).permute(0, 3, 1, 2)[:, :2]
    np.testing.assert_allclose(up.cpu().numpy(), v["g5_up"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("reso", [2, 64, 100, 256, 512])
def test_cell_index_bit_exact(T, reso):
    g = torch.Generator().manual_seed(reso)
    xy = torch.rand(3, 50000, 2, generator=g)
    xy[0, :6] = torch.tensor([[2.0 ** -24, 1 - 2.0 ** -24], [1 - 2.0 ** -24, 2.0 ** -24], [0.5, 0.5],
                              [0.25, 0.75], [0.999999, 0.000001], [1 / 3, 2 / 3]])
    want = oracle.cell_index(xy, reso)
    got = T.cell_index(xy.cuda(), reso)
    assert got.dtype == torch.int64 and got.shape == want.shape
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("B,N,reso", [(1, 5000, 64), (3, 4000, 32), (2, 3000, 100), (1, 1, 8), (2, 777, 256)])
def test_topology_sort(B, N, reso):
    cloud = synthetic_cloud(B, N, seed=B * 1000 + N)
    topo = _topo(cloud, reso)
    perm = topo.perm.cpu().long()
    # a permutation that keeps tiles separate
    assert torch.equal(perm.sort().values, torch.arange(B * N))
    assert torch.equal(perm // N, torch.arange(B * N) // N)
    assert torch.equal(topo.xyz_sorted.cpu()[:, :3], cloud.view(-1, 3)[perm])  # rows are padded to 16 bytes
    keys = topo.keys_sorted.cpu().long()
    assert bool((keys[1:] >= keys[:-1]).all())
    # stable: equal keys keep increasing point index
    same = keys[1:] == keys[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    # cell_start is the exclusive histogram scan of the keys
    n_keys = B * reso * reso
    hist = torch.bincount(keys, minlength=n_keys)
    want = torch.cat([torch.zeros(1, dtype=torch.long), hist.cumsum(0)])
    assert torch.equal(topo.cell_start.cpu().long(), want)
    # every level's cell id agrees with the reference rule (floor(p*r) == floor(p*R) >> k)
    for r in ([reso, reso // 2, reso // 4] if topo.morton and reso >= 8 else [reso]):
        lvl = topo.level(r)
        ref_idx = oracle.cell_index(topo.xyz_sorted.cpu().view(B, N, -1)[..., :2], r).view(-1)
        seg_of = torch.bucketize(torch.arange(B * N), topo.cell_start.cpu().long()[:: (1 << lvl.shift)][1:], right=True)
        if topo.morton:
            from tomosar2height_b200.tests_support import demorton
            b = seg_of // (r * r)
            ix, iy = demorton(seg_of % (r * r))
            got = ix + r * iy
        else:
            got = seg_of % (r * r)
        assert torch.equal(got, ref_idx), r


def _oracle_rows(fn, rows, cloud, reso, B, N):
    """run an oracle (B, C, N) op on (B*N, C) rows given in ORIGINAL point order"""
    src = rows.view(B, N, -1).permute(0, 2, 1)
    idx = oracle.cell_index(cloud[..., :2], reso)
    return fn(src, idx, reso * reso)


@pytest.mark.parametrize("C", [4, 32, 64, 128, 256])
@pytest.mark.parametrize("sorted_rows", [True, False])
def test_seg_max_pool_fwd_bwd(T, C, sorted_rows):
    B, N, R = 2, 6000, 64
    cloud = synthetic_cloud(B, N, seed=C)
    # a facade: 2500 points of tile 0 in ONE fine cell and 900 of tile 1 in another (the kernels hand cells with
    # more than 128 rows to the whole CTA), with exact ties among them
    cloud[0, 500:3000, :2] = torch.tensor([0.4005, 0.7003]) + 1e-4 * torch.rand(2500, 2, generator=torch.Generator().manual_seed(7))
    cloud[1, 100:1000, :2] = torch.tensor([0.9101, 0.0303]) + 1e-4 * torch.rand(900, 2, generator=torch.Generator().manual_seed(8))
    topo = _topo(cloud, R)
    g = torch.Generator().manual_seed(C + 1)
    feat = torch.randn(B * N, C, generator=g)
    feat[:, 0] = (feat[:, 0] * 2).round() / 2  # many exact ties
    feat[:, 1] = 0.25                            # all equal
    w = torch.randn(B * N, C, generator=g)
    for r in (R, R // 4):
        lvl = topo.level(r, rows_sorted=sorted_rows)
        perm = topo.perm.cpu().long()
        rows_dev = (feat[perm] if sorted_rows else feat).cuda().requires_grad_(True)
        pooled, arg = T.seg_max_pool(rows_dev, lvl, return_arg=True)
        (pooled * (w[perm] if sorted_rows else w).cuda()).sum().backward()
        # oracle in original order
        f_cpu = feat.clone().requires_grad_(True)
        cells, arg_ref = _oracle_rows(oracle.segment_max, f_cpu, cloud, r, B, N)
        idx = oracle.cell_index(cloud[..., :2], r)
        pooled_ref = cells.gather(2, idx.expand(-1, C, -1)).permute(0, 2, 1).reshape(B * N, C)
        (pooled_ref * w).sum().backward()
        got_pooled = pooled.detach().cpu()
        got_grad = rows_dev.grad.cpu()
        if sorted_rows:
            inv = torch.empty_like(perm); inv[perm] = torch.arange(B * N)
            got_pooled, got_grad = got_pooled[inv], got_grad[inv]
        assert torch.equal(got_pooled, pooled_ref.detach()), "max values are exact"
        # argmax: row index -> point index within the tile; bit-exact incl. ties and empty cells
        a = arg.cpu().long().view(B, r * r, C)
        rows_to_point = perm if sorted_rows else torch.arange(B * N)
        a_pt = torch.where(a < 0, torch.full_like(a, N), rows_to_point[a.clamp(min=0)] - (torch.arange(B) * N).view(B, 1, 1))
        assert torch.equal(a_pt.permute(0, 2, 1), arg_ref), f"argmax mismatch at r={r}"
        _close(got_grad, f_cpu.grad, what=f"seg_max grad r={r}")


@pytest.mark.parametrize("C", [8, 32, 128, 512])
@pytest.mark.parametrize("sorted_rows", [True, False])
def test_seg_mean_fwd_bwd(T, C, sorted_rows):
    B, N, R = 2, 5000, 32
    cloud = synthetic_cloud(B, N, seed=100 + C)
    topo = _topo(cloud, R)
    g = torch.Generator().manual_seed(C + 7)
    feat = torch.randn(B * N, C, generator=g)
    perm = topo.perm.cpu().long()
    for r in (R, R // 2, R // 8):
        lvl = topo.level(r, rows_sorted=sorted_rows)
        w = torch.randn(B, C, r * r, generator=g)
        rows_dev = (feat[perm] if sorted_rows else feat).cuda().requires_grad_(True)
        plane = T.seg_mean(rows_dev, lvl)                      # (B*r*r, C)
        (plane.view(B, r * r, C).permute(0, 2, 1) * w.cuda()).sum().backward()
        f_cpu = feat.clone().requires_grad_(True)
        ref = _oracle_rows(oracle.segment_mean, f_cpu, cloud, r, B, N)
        (ref * w).sum().backward()
        _close(plane.view(B, r * r, C).permute(0, 2, 1), ref, what=f"seg_mean r={r}")
        assert bool(((plane.view(B, r * r, C).abs().sum(-1) == 0).cpu() == (ref.abs().sum(1) == 0)).all()), "empty cells are exactly 0"
        got_grad = rows_dev.grad.cpu()
        if sorted_rows:
            inv = torch.empty_like(perm); inv[perm] = torch.arange(B * N)
            got_grad = got_grad[inv]
        _close(got_grad, f_cpu.grad, what=f"seg_mean grad r={r}")


def test_seg_broadcast_mean_pool(T):
    """pool_local with scatter_type='mean': mean then gather-back, and its backward."""
    B, N, R, C = 2, 3000, 32, 32
    cloud = synthetic_cloud(B, N, seed=5)
    topo = _topo(cloud, R)
    lvl = topo.level(R)
    perm = topo.perm.cpu().long()
    g = torch.Generator().manual_seed(9)
    feat, w = torch.randn(B * N, C, generator=g), torch.randn(B * N, C, generator=g)
    rows_dev = feat[perm].cuda().requires_grad_(True)
    pooled = T.seg_broadcast(T.seg_mean(rows_dev, lvl), lvl)
    (pooled * w[perm].cuda()).sum().backward()
    f_cpu = feat.clone().requires_grad_(True)
    idx = oracle.cell_index(cloud[..., :2], R)
    cells = oracle.segment_mean(f_cpu.view(B, N, C).permute(0, 2, 1), idx, R * R)
    ref = cells.gather(2, idx.expand(-1, C, -1)).permute(0, 2, 1).reshape(B * N, C)
    (ref * w).sum().backward()
    inv = torch.empty_like(perm); inv[perm] = torch.arange(B * N)
    _close(pooled.detach().cpu()[inv], ref, what="mean pool")
    _close(rows_dev.grad.cpu()[inv], f_cpu.grad, what="mean pool grad")


@pytest.mark.parametrize("C,reso", [(32, 64), (64, 64), (128, 32), (256, 16), (512, 8), (4, 100)])
@pytest.mark.parametrize("sorted_rows", [True, False])
def test_bilinear_sample_fwd_bwd(T, C, reso, sorted_rows):
    B, N = 2, 4000
    R = 64 if reso != 100 else 100
    cloud = synthetic_cloud(B, N, seed=300 + C)
    cloud[0, 0, :2] = torch.tensor([2.0 ** -24, 1 - 2.0 ** -24])
    cloud[0, 1, :2] = torch.tensor([1 - 2.0 ** -24, 1 - 2.0 ** -24])
    topo = _topo(cloud, R)
    lvl = topo.level(reso, rows_sorted=sorted_rows)
    perm = topo.perm.cpu().long()
    g = torch.Generator().manual_seed(C)
    plane = torch.randn(B, C, reso, reso, generator=g)
    w = torch.randn(B * N, C, generator=g)
    plane_dev = plane.cuda().requires_grad_(True)
    rows = T.bilinear_sample(T.nchw_to_plane(plane_dev), lvl)
    (rows * (w[perm] if sorted_rows else w).cuda()).sum().backward()
    p_cpu = plane.clone().requires_grad_(True)
    ref = oracle.bilinear_sample_points(p_cpu, cloud[..., :2]).permute(0, 2, 1).reshape(B * N, C)
    (ref * w).sum().backward()
    got = rows.detach().cpu()
    if sorted_rows:
        inv = torch.empty_like(perm); inv[perm] = torch.arange(B * N)
        got = got[inv]
    _close(got, ref, what="sample fwd")
    _close(plane_dev.grad, p_cpu.grad, what="sample bwd")
    # and against ATen itself (the reference's dependency)
    aten = torch.nn.functional.grid_sample(plane, (2 * cloud[..., :2] - 1)[:, :, None], padding_mode="border",
                                           align_corners=True).squeeze(-1).permute(0, 2, 1).reshape(B * N, C)
    _close(got, aten, what="sample fwd vs ATen")


@pytest.mark.parametrize("C,h,size", [(32, 64, 128), (32, 16, 16), (4, 7, 20), (64, 32, 64), (32, 33, 50)])
def test_upsample_fwd_bwd(T, C, h, size):
    B = 2
    g = torch.Generator().manual_seed(h)
    plane = torch.randn(B, C, h, h, generator=g)
    w = torch.randn(B, C, size, size, generator=g)
    dev = plane.cuda().requires_grad_(True)
    out = T.upsample_bilinear(T.nchw_to_plane(dev), size).permute(0, 3, 1, 2)
    (out * w.cuda()).sum().backward()
    cpu = plane.clone().requires_grad_(True)
    ref = torch.nn.functional.interpolate(cpu, size=size, mode="bilinear", align_corners=True)
    (ref * w).sum().backward()
    _close(out, ref, what="upsample fwd")
    _close(dev.grad, cpu.grad, what="upsample bwd")
    _close(out, oracle.upsample_bilinear_align(plane, size), what="upsample vs explicit oracle")


def test_scatter_api_matches_oracle():
    """torch_scatter-style calls in arbitrary point order (pointnet.py:95,109), incl. odd channel counts."""
    from tomosar2height_b200 import scatter as S
    B, C, N, M = 2, 5, 3000, 97
    g = torch.Generator().manual_seed(1)
    src = torch.randn(B, C, N, generator=g)
    src[:, 0] = src[:, 0].round()
    index = torch.randint(0, M, (B, 1, N), generator=g)
    index[index == 13] = 12  # an empty segment
    s_dev = src.cuda().requires_grad_(True)
    out, arg = S.scatter_max(s_dev, index.cuda(), dim_size=M)
    w = torch.randn(B, C, M, generator=g)
    (out * w.cuda()).sum().backward()
    s_cpu = src.clone().requires_grad_(True)
    ref, arg_ref = oracle.segment_max(s_cpu, index, M)
    (ref * w).sum().backward()
    assert torch.equal(out.detach().cpu(), ref.detach()) and torch.equal(arg.cpu(), arg_ref)
    _close(s_dev.grad, s_cpu.grad, what="scatter_max grad")
    mean = S.scatter_mean(src.cuda(), index.cuda(), dim_size=M)
    _close(mean, oracle.segment_mean(src, index, M), what="scatter_mean")
    with pytest.raises(IndexError):
        S.scatter_mean(src.cuda(), (index + M).cuda(), dim_size=M)


def test_empty_and_single_point_tiles(T):
    cloud = torch.tensor([[[0.3, 0.7, 0.1]]])
    topo = _topo(cloud, 8)
    lvl = topo.level(8)
    rows = torch.arange(32, dtype=torch.float32).view(1, 32).cuda()
    plane = T.seg_mean(rows, lvl).view(8, 8, 32)
    assert float(plane.abs().sum()) == float(rows.abs().sum())
    assert torch.equal(plane[5, 2].cpu(), rows[0].cpu())  # iy = floor(0.7*8) = 5, ix = 2
    pooled, arg = T.seg_max_pool(rows, lvl, return_arg=True)
    assert torch.equal(pooled, rows) and int((arg >= 0).sum()) == 32


def test_determinism(T):
    B, N, R, C = 2, 20000, 64, 64
    cloud = synthetic_cloud(B, N, seed=77)
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(B * N, C, generator=g).cuda()
    outs = []
    for _ in range(3):
        topo = _topo(cloud, R)
        lvl = topo.level(R // 2)
        plane = T.seg_mean(topo.sort_rows(feat), lvl)
        x = plane.view(B, R // 2, R // 2, C).detach().requires_grad_(True)
        T.bilinear_sample(x, lvl).square().sum().backward()
        outs.append((plane.clone(), x.grad.clone()))
    for p, gq in outs[1:]:
        assert torch.equal(p, outs[0][0]) and torch.equal(gq, outs[0][1]), "bitwise run-to-run determinism"


@pytest.mark.parametrize("n,n_keys", [(1, 5), (33, 2), (5000, 1), (100000, 255), (100000, 257), (1 << 20, 4 * 65536),
                                      (300000, (1 << 24) + 3), (70001, (1 << 31) - 2)])
def test_radix_sort_matches_torch_stable_sort(n, n_keys):
    """The hand-written LSD radix sort (t2h_sort_by_cell): keys sorted, STABLE (equal keys keep input order),
    cell_start = exclusive histogram scan; 1 to 4 eight-bit passes."""
    from tomosar2height_b200.topology import sort_keys
    g = torch.Generator().manual_seed(n % 1000 + 1)
    keys = torch.randint(0, min(n_keys, 1 << 31) , (n,), generator=g, dtype=torch.int64)
    keys[: n // 3] = keys[0]  # long runs of equal keys
    k32 = keys.to(torch.int32).cuda()
    keys_sorted, perm, cell_start = sort_keys(k32, n_keys, with_cell_start=n_keys <= 4 * 65536)
    want_k, want_p = torch.sort(keys, stable=True)
    assert torch.equal(keys_sorted.cpu().long(), want_k)
    assert torch.equal(perm.cpu().long(), want_p)
    if n_keys <= 4 * 65536:
        hist = torch.bincount(keys, minlength=n_keys)
        want = torch.cat([torch.zeros(1, dtype=torch.long), hist.cumsum(0)])
        assert torch.equal(cell_start.cpu().long(), want)


def test_out_of_range_points_are_flagged():
    """coordinate2index does not clamp and torch_scatter would fail on such an index; the topology bins the point
    into a border cell (no kernel can run out of bounds) and raises on request."""
    from tomosar2height_b200.topology import Topology
    cloud = synthetic_cloud(1, 500, seed=2)
    ok = Topology(cloud.cuda(), 32, check_range=True)
    assert int(ok.range_flag.item()) == 0
    for bad in (1.0, -1e-3, float("nan")):
        c = cloud.clone()
        c[0, 7, 0] = bad
        t = Topology(c.cuda(), 32)
        assert int(t.range_flag.item()) == 1 and int(t.cell_start[-1]) == 500
        with pytest.raises(IndexError):
            Topology(c.cuda(), 32, check_range=True)

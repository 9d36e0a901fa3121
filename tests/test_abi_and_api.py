"""CPU-side checks: the C-ABI library exports what include/t2h.h declares, the Python mirror keeps
the reference's API surface / state_dict layout, and the product path refuses to run without CUDA."""
import json
import os
import re

import pytest
import torch

import tomosar2height_b200 as t2h
from tomosar2height_b200 import _lib
from cases import CASES, make_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "t2h.h")).read()
    declared = set(re.findall(r"\b(t2h_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.t2h_abi_version() == 1
    assert lib.t2h_status_string(0) == b"ok"
    assert lib.t2h_status_string(2) == b"unsupported shape"
    assert lib.t2h_sort_workspace_bytes(1 << 20) > (1 << 22)


def test_header_cites_reference():
    header = open(os.path.join(ROOT, "include", "t2h.h")).read()
    for cite in ("utils/coordinate.py:12-28", "pointnet.py:92-99", "alto.py:90-95", "pixel.py:105-111"):
        assert cite in header


def test_argument_validation_without_gpu(lib):
    # invalid arguments are rejected before any launch, so this is safe on a CPU box
    assert lib.t2h_cell_index(None, 10, 2, 256, None, None) == 1
    # rows, n_rows, perm, row_keys, cell_start, n_seg, shift, C, morton, reso, mean, ws, ws_bytes, plane, stream
    assert lib.t2h_seg_reduce_fwd(None, 8, None, None, None, 8, 0, 32, 0, 4, 1, None, 0, None, None) == 1
    assert lib.t2h_seg_mean_fwd(None, 8, None, None, None, 8, 0, 32, 0, 4, None, 0, None, None) == 1
    assert lib.t2h_seg_workspace_bytes(1 << 20, 1 << 16, 32) >= (1 << 20) // 32 * 2 * 32 * 8
    assert lib.t2h_xy_keys(None, 0, 3, 1, 100, 1, None, None, None) == 1


def test_block_workspace_queries_plan_without_a_gpu(lib):
    """t2h_resblock / t2h_comm_mlp: the workspace query runs the launch plan dry (host only)"""
    small = lib.t2h_resblock_workspace_bytes(1000, 32, 32, 32, 32, 1)
    large = lib.t2h_resblock_workspace_bytes(100000, 32, 32, 32, 32, 1)
    assert 256 < small < large                        # g_net scratch grows with the rows
    assert lib.t2h_resblock_workspace_bytes(1000, 30, 0, 32, 32, 1) == 256     # unsupported width: nothing to plan
    assert lib.t2h_resblock_workspace_bytes(1000, 64, 0, 32, 32, 0) == 256     # identity shortcut needs size_in == size_out
    first = lib.t2h_comm_mlp_workspace_bytes(1000, 128, 0)
    later = lib.t2h_comm_mlp_workspace_bytes(1000, 128, 64)
    assert 256 < first < later                        # fc_c adds a weight gradient and a transposed weight
    # bad arguments are refused before anything is launched
    assert lib.t2h_resblock_fwd(None, 0, 32, None, 0, 0, 10, None, None, 32, None, None, None, 32, None, 0, None, 0, None, 0, None) == 1
    assert lib.t2h_comm_mlp_fwd(None, 0, 32, None, 0, 0, 10, None, None, None, None, None, None, None, 0, None, 0, None, 0, None) == 1


@pytest.mark.parametrize("name", list(CASES) + ["berlin_full", "berlin_image_full", "munich_full", "munich_image_full"])
def test_state_dict_matches_reference(golden_dir, name):
    with open(os.path.join(golden_dir, f"state_dict_{name}.json")) as fh:
        ref = {k: tuple(v) for k, v in json.load(fh).items()}
    if name in CASES:
        cfg = make_cfg(**CASES[name]["cfg"])
    else:
        cfg = (t2h.berlin_config if name.startswith("berlin") else t2h.munich_config)(use_image="image" in name)
    model = t2h.TomoSAR2Height(cfg)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == ref
    assert list(got) == list(ref) or sorted(got) == sorted(ref)


def test_registries_and_errors():
    assert set(t2h.encoder_dict) == {"pointnet_local_pool", "pointnet_plus_plus", "hourglass", "unet"}
    assert set(t2h.decoder_dict) == {"pixel"}
    from tomosar2height_b200.encoder.pointnet import LocalPoolPointnet
    from tomosar2height_b200.decoder.pixel import PixelwiseDecoder
    from tomosar2height_b200.encoder.alto import UNet as Alto
    with pytest.raises(ValueError):
        LocalPoolPointnet(unet_type="nope", unet_kwargs={})
    with pytest.raises(ValueError):
        LocalPoolPointnet(scatter_type="nope", unet_kwargs={"depth": 2, "start_filts": 8}, feature_dim=8, hidden_dim=8)
    with pytest.raises(ValueError):
        PixelwiseDecoder(mode="nope")
    with pytest.raises(ValueError):
        Alto(8, in_channels=8, depth=2, start_filts=8, up_mode="nope")
    with pytest.raises(ValueError):
        Alto(8, in_channels=8, depth=2, start_filts=8, up_mode="upsample", merge_mode="add")
    with pytest.raises(NotImplementedError):
        t2h.encoder_dict["hourglass"]()
    blk = t2h.ResnetBlockFC(64, 32)
    assert blk.shortcut is not None and blk.shortcut.bias is None and t2h.ResnetBlockFC(32).shortcut is None
    assert float(blk.fc_1.weight.abs().max()) == 0.0


def test_init_rule_matches_reference():
    torch.manual_seed(0)
    model = t2h.TomoSAR2Height(make_cfg(**CASES["berlin_small"]["cfg"]))
    for m in model.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and m.bias is not None:
            assert float(m.bias.abs().max()) == 0.0
    # z_scale, threshold
    assert abs(model.z_scale - 190.2) < 1e-9 and model.threshold == 0.5


def test_no_cpu_fallback():
    model = t2h.TomoSAR2Height(make_cfg(**CASES["berlin_small"]["cfg"]))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(input_cloud=torch.rand(1, 100, 3))
    from tomosar2height_b200.utils import coordinate2index
    with pytest.raises(RuntimeError, match="CUDA"):
        coordinate2index(torch.rand(1, 10, 2), 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tomosar2height_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_install_as_reference_alias():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "tomosar2height" or k.startswith("tomosar2height.")}
    try:
        t2h.install_as_reference()
        from tomosar2height import TomoSAR2Height
        from tomosar2height.encoder import encoder_dict  # noqa: F401
        assert TomoSAR2Height is t2h.TomoSAR2Height
    finally:
        for k in [k for k in sys.modules if k == "tomosar2height" or k.startswith("tomosar2height.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_absmax_registry_never_returns_a_stale_entry():
    """host logic of linear._AbsmaxRegistry (CPU tensors stand in; the slot is an opaque object): an entry is valid
    only for the live, unmodified tensor it was registered for, and for views of it."""
    import gc
    import torch
    from tomosar2height_b200.linear import _AbsmaxRegistry
    reg = _AbsmaxRegistry()
    t = torch.randn(64, 8)
    slot = object()
    reg.put(t, slot)
    assert reg.get(t) is slot
    assert reg.get(t.view(32, 16)) is slot            # same storage, same version: the maximum is shape-independent
    assert reg.get(t[:32]) is None                    # different extent
    assert reg.get(t.t()) is None                     # not contiguous
    t.add_(1.0)                                       # modified in place: version counter moved on
    assert reg.get(t) is None
    reg.put(t, slot)
    assert reg.get(t) is slot
    ptr, n = t.data_ptr(), t.numel()
    del t
    gc.collect()
    assert (ptr, n) not in reg._entries               # entry dropped with its owner: a recycled address cannot hit
    a = torch.randn(16, 4)
    b = a.clone()
    reg.put(a, slot)
    assert reg.get(b) is None                         # equal contents, different storage


def test_ctypes_signatures_have_the_arity_of_the_header():
    """A missing argtype (e.g. the trailing stream) makes ctypes pass a 64-bit handle as a C int."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "t2h.h")).read(), flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(t2h_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name]), (name, n, len(_lib.SIGNATURES[name]))
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_custom_ops_propagate_shapes_without_a_gpu():
    """The torch custom-op layer (t2h::*) over the C ABI: every op has a register_fake shape function, so meta / fake
    tensors flow through it with no device (what torch.compile / export need); real CPU tensors still raise."""
    import torch
    import tomosar2height_b200  # noqa: F401  (registers the ops)
    m = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device="meta")
    n, C, M, B, r = 100, 32, 64, 1, 8
    keys, cs = m(n, dt=torch.int32), m(M + 1, dt=torch.int32)
    ops = torch.ops.t2h
    assert ops.seg_reduce(m(n, C), None, keys, cs, M, 0, 1, r, True).shape == (M, C)
    assert ops.seg_broadcast(m(M, C), n, None, keys, cs, M, 0, 1, r, True).shape == (n, C)
    rows, slot = ops.seg_broadcast_add(m(M, C), m(n, C), None, keys, cs, M, 0, 1, r, True)
    assert rows.shape == (n, C) and slot.shape == (1,)
    pooled, plane, arg = ops.seg_max(m(n, C), None, None, keys, cs, M, 0, 1, r, True)
    assert pooled.shape == (n, C) and plane.shape == (M, C) and arg.dtype == torch.int32
    assert ops.seg_max_bwd(m(n, C), None, arg, n, None, keys, cs, M, 0, 1, r).shape == (n, C)
    assert ops.bilinear_sample(m(B, r, r, C), m(n, 4), None, None, keys, cs, M, 0, 1, n).shape == (n, C)
    assert ops.bilinear_sample_bwd(m(n, C), B, r, m(n, 4), None, keys, cs, M, 0, 1).shape == (B, r, r, C)
    assert ops.upsample_bilinear(m(B, r, r, C), 16, 16).shape == (B, 16, 16, C)
    assert ops.upsample_bilinear_bwd(m(B, 16, 16, C), r, r).shape == (B, r, r, C)
    assert ops.cell_index(m(2, n, 2), 256).shape == (2, 1, n)
    assert ops.gather_rows(m(n, C), m(n, dt=torch.int32)).shape == (n, C)
    assert ops.linear(m(n, 64), m(n, 64), m(32, 128), m(32), m(n, 32), True).shape == (n, 32)
    dx1, dx2, dw, db = ops.linear_bwd(m(n, 32), m(n, 64), m(n, 64), m(32, 128), True, True, True, True, True, True)
    assert dx1.shape == (n, 64) and dx2.shape == (n, 64) and dw.shape == (32, 128) and db.shape == (32,)
    assert ops.conv3x3(m(B, 16, 16, C), m(64, C, 3, 3), None, False).shape == (B, 16, 16, 64)
    dx, dwc, dbc = ops.conv3x3_bwd(m(B, 16, 16, 64), m(B, 16, 16, C), m(64, C, 3, 3), False, True, True, True)
    assert dx.shape == (B, 16, 16, C) and dwc.shape == (64, C, 3, 3) and dbc.shape == (64,)
    with pytest.raises(RuntimeError):
        ops.seg_reduce(torch.zeros(n, C), None, torch.zeros(n, dtype=torch.int32), torch.zeros(M + 1, dtype=torch.int32), M, 0, 1, r, True)

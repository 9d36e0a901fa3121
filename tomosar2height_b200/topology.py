"""Point -> cell topology: one stable sort of the tile batch serves every plane level.

Replaces the per-call ``coordinate2index`` + atomic scatter of the reference
(pointnet.py:69-70, alto.py:79-80,189-190).  See include/t2h.h for the key layout.
"""
import os

import torch

from . import _lib


_CHECK_RANGE = os.environ.get("T2H_CHECK_RANGE", "0") == "1"


def _is_pow2(v: int) -> bool:
    return v > 0 and (v & (v - 1)) == 0


class RaggedCloud:
    """A batch of tiles with DIFFERENT point counts: flat ``points`` (P, 3) fp32 + ``offsets`` (B+1,) int64,
    tile b owning points[offsets[b]:offsets[b+1]] (xy normalised to the open unit square per tile).
    The reference is stuck at batch 1 because its dense (B, N, 3) collate cannot hold such tiles
    (conf/model/tomosar2height.yaml:40); every kernel here works on flat rows + cell keys, so ragged
    batches cost nothing extra.  Pass it as ``input_cloud``."""

    def __init__(self, points: torch.Tensor, offsets: torch.Tensor):
        self.points, self.offsets = points, offsets

    @classmethod
    def from_list(cls, tiles):
        sizes = torch.tensor([0] + [int(t.shape[0]) for t in tiles], dtype=torch.int64)
        points = torch.cat([t.reshape(-1, t.shape[-1]) for t in tiles], 0).contiguous()
        return cls(points, sizes.cumsum(0).to(points.device))

    @property
    def n_tiles(self):
        return self.offsets.numel() - 1


class CellLevel:
    """What the segment / sampling kernels need for one plane resolution ``reso``.

    ``perm`` is None when per-point rows are stored in sorted order (the fused model path),
    otherwise it maps sorted position -> row.
    """

    __slots__ = ("topo", "reso", "shift", "morton", "perm", "tie", "keys", "cell_start", "xyz_sorted", "B", "N", "n_seg",
                 "n_points", "tile_ids")

    def __init__(self, topo, reso, shift, perm, tie=None):
        self.topo = topo
        self.reso = reso
        self.shift = shift
        self.morton = int(topo.morton)
        self.perm = perm
        # argmax ties go to the smallest ORIGINAL point index; inside one sort key the stable sort
        # already guarantees that, a coarser segment (shift > 0) needs the explicit rank
        self.tie = tie if tie is not None else (topo.perm if shift > 0 else None)
        self.keys = topo.keys_sorted   # sort key of every sorted position; this level's segment = key >> shift
        self.cell_start = topo.cell_start
        self.xyz_sorted = topo.xyz_sorted
        self.B, self.N = topo.B, topo.N
        self.n_seg = topo.B * reso * reso
        self.n_points = topo.n_points
        self.tile_ids = topo.tile_ids  # None for dense (B, N, 3) batches

    @property
    def n_rows(self):
        return self.n_points


class Topology:
    """Sort the points of a (B, N, 3) tile batch by (tile, cell) once.

    Attributes
      xyz_sorted  (B*N, 4) fp32 : coordinates in sorted order, zero-padded to 16-byte rows
                                  (vector loads in the sampling kernels, TMA-loadable for fc_pos)
      perm        (B*N,) int32  : sorted position -> flat input index  (b*N + n)
      cell_start  (B*R*R + 1,) int32
      morton      bool          : Morton keys (power-of-two R) -> coarser levels share the sort
    """

    def __init__(self, xyz: torch.Tensor, reso: int, offsets: torch.Tensor = None, check_range: bool = False):
        """xyz (B, N, >=2) for a dense batch, or -- with ``offsets`` (B+1,) int64 on the device -- the flat
        (P, >=2) cloud of a RAGGED batch whose tile b owns the points [offsets[b], offsets[b+1]).

        The reference does not clamp cell coordinates: a point outside [0, 1)^2 makes torch_scatter raise on the
        index (coordinate.py:24-26, dataset.py:278 keeps real data inside).  The kernels bin such points into the
        border cells and raise a device flag, ``range_flag``; ``check_range=True`` (or env T2H_CHECK_RANGE=1) reads
        it (one host sync) and raises IndexError like the reference path would."""
        _lib.require_cuda_f32(xyz, "Topology(xyz)")
        ragged = offsets is not None
        if xyz.dim() != (2 if ragged else 3) or xyz.shape[-1] < 2:
            raise RuntimeError(f"Topology: expected {'(P, >=2)' if ragged else '(B, N, >=2)'} points, got {tuple(xyz.shape)}")
        if xyz.shape[-1] > 4:
            raise RuntimeError("Topology: at most 4 coordinates per point are supported")
        if xyz.shape[-1] != 4:
            xyz = torch.nn.functional.pad(xyz, (0, 4 - xyz.shape[-1]))
        xyz = xyz.contiguous()
        self.D = 4
        self.reso = int(reso)
        self.morton = _is_pow2(self.reso)
        dev = xyz.device
        if ragged:
            if offsets.dtype != torch.int64 or not offsets.is_cuda or offsets.dim() != 1 or offsets.numel() < 2:
                raise RuntimeError("Topology: offsets must be a CUDA int64 vector of length B + 1")
            self.B, self.N = offsets.numel() - 1, None
            n = xyz.shape[0]
            keys = torch.empty(n, dtype=torch.int32, device=dev)
            self.range_flag = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("t2h_xy_keys_ragged", _lib.ptr(xyz), n, self.D, _lib.ptr(offsets.contiguous()), self.B, self.reso,
                      int(self.morton), _lib.ptr(keys), _lib.ptr(self.range_flag))
        else:
            self.B, self.N = xyz.shape[0], xyz.shape[1]
            n = self.B * self.N
            keys = torch.empty(n, dtype=torch.int32, device=dev)
            self.range_flag = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("t2h_xy_keys", _lib.ptr(xyz), n, self.D, self.N, self.reso, int(self.morton), _lib.ptr(keys),
                      _lib.ptr(self.range_flag))
        if check_range or _CHECK_RANGE:
            if int(self.range_flag.item()) != 0:
                raise IndexError(f"Topology: a point lies outside the unit square [0, 1)^2 (cell index out of range for reso {self.reso})")
        self.n_points = n
        n_keys = self.B * self.reso * self.reso
        self.keys_sorted, self.perm, self.cell_start = sort_keys(keys, n_keys)
        self.tile_ids = (self.keys_sorted // (self.reso * self.reso)).to(torch.int32) if ragged else None
        self.offsets = offsets
        self.xyz_sorted = torch.empty(n, self.D, dtype=torch.float32, device=dev)
        _lib.call("t2h_gather_rows", _lib.ptr(xyz.view(n, self.D)), _lib.ptr(self.perm), n, self.D,
                  _lib.ptr(self.xyz_sorted))
        self._sub = {}

    # -- levels ------------------------------------------------------------------------------
    def _shift_for(self, reso: int):
        if reso == self.reso:
            return 0
        if self.morton and _is_pow2(reso) and reso < self.reso:
            k = (self.reso // reso).bit_length() - 1
            return 2 * k
        return None

    def level(self, reso: int, rows_sorted: bool = True) -> CellLevel:
        """Level descriptor for plane resolution ``reso``.

        rows_sorted=True : per-point rows live in THIS topology's sorted order.
        rows_sorted=False: rows live in the original input order.
        A resolution that does not nest in the Morton keys (non power-of-two planes) gets its own
        sort of the already-sorted coordinates, cached per resolution.
        """
        shift = self._shift_for(reso)
        if shift is not None:
            return CellLevel(self, reso, shift, None if rows_sorted else self.perm)
        key = (reso, rows_sorted)
        if key not in self._sub:
            if rows_sorted and self.N is None:
                raise RuntimeError("Topology.level: ragged batches need power-of-two plane resolutions")
            if rows_sorted:
                sub = Topology(self.xyz_sorted.view(self.B, self.N, self.D), reso)
            else:
                raise RuntimeError("Topology.level: incompatible resolution for original-order rows; "
                                   "build a Topology at that resolution instead")
            self._sub[key] = sub
        sub = self._sub[key]
        if not hasattr(sub, "origin"):
            sub.origin = self.perm[sub.perm.long()].contiguous()  # sub-sorted position -> original index
        return CellLevel(sub, reso, 0, sub.perm, tie=sub.origin)

    def sort_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """(B*N, C) rows in input order -> sorted order."""
        return gather_rows(rows, self.perm)

    def unsort_rows(self, rows: torch.Tensor) -> torch.Tensor:
        return scatter_rows(rows, self.perm)


def sort_keys(keys: torch.Tensor, n_keys: int, with_cell_start: bool = True):
    """Stable sort of int32 keys in [0, n_keys) -> (keys_sorted, perm, cell_start[n_keys + 1]).
    ``with_cell_start=False``: a plain stable sort (no first-position table)."""
    n = keys.numel()
    dev = keys.device
    lib = _lib.load()
    ws_bytes = int(lib.t2h_sort_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    keys_sorted = torch.empty_like(keys)
    perm = torch.empty_like(keys)
    cell_start = torch.empty(n_keys + 1, dtype=torch.int32, device=dev) if with_cell_start else None
    _lib.call("t2h_sort_by_cell", _lib.ptr(keys), n, n_keys, _lib.ptr(ws), ws_bytes, _lib.ptr(keys_sorted),
              _lib.ptr(perm), _lib.ptr(cell_start))
    return keys_sorted, perm, cell_start


@torch.library.custom_op("t2h::gather_rows", mutates_args=())
def gather_rows(rows: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[i] = rows[perm[i]] for a permutation ``perm``; differentiable (the backward is ``scatter_rows``)."""
    _lib.require_cuda_f32(rows, "gather_rows")
    rows = rows.contiguous()
    out = torch.empty_like(rows)
    _lib.call("t2h_gather_rows", _lib.ptr(rows), _lib.ptr(perm), rows.shape[0], rows.shape[1], _lib.ptr(out))
    return out


@torch.library.custom_op("t2h::scatter_rows", mutates_args=())
def scatter_rows(rows: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[perm[i]] = rows[i] for a permutation ``perm``; differentiable (the backward is ``gather_rows``)."""
    _lib.require_cuda_f32(rows, "scatter_rows")
    rows = rows.contiguous()
    out = torch.empty_like(rows)
    _lib.call("t2h_scatter_rows", _lib.ptr(rows), _lib.ptr(perm), rows.shape[0], rows.shape[1], _lib.ptr(out))
    return out


@gather_rows.register_fake
def _(rows, perm):
    return torch.empty_like(rows)


@scatter_rows.register_fake
def _(rows, perm):
    return torch.empty_like(rows)


def _perm_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])


gather_rows.register_autograd(lambda ctx, g: (scatter_rows(g.contiguous(), ctx.saved_tensors[0]), None), setup_context=_perm_setup)
scatter_rows.register_autograd(lambda ctx, g: (gather_rows(g.contiguous(), ctx.saved_tensors[0]), None), setup_context=_perm_setup)


class IndexLevel:
    """Level descriptor for an arbitrary (B, 1, N) int64 index (torch_scatter-style API)."""

    __slots__ = ("reso", "shift", "morton", "perm", "tie", "keys", "cell_start", "xyz_sorted", "B", "N", "n_seg", "dim_size",
                 "n_points", "tile_ids")

    def __init__(self, index: torch.Tensor, dim_size: int, check: bool = True):
        if not index.is_cuda or index.dtype != torch.int64:
            raise RuntimeError("IndexLevel: expected a CUDA int64 index")
        B = index.shape[0]
        N = index.shape[-1]
        index = index.reshape(B, N).contiguous()
        n = B * N
        keys = torch.empty(n, dtype=torch.int32, device=index.device)
        flag = torch.zeros(1, dtype=torch.int32, device=index.device)
        _lib.call("t2h_index_keys", _lib.ptr(index), n, N, dim_size, _lib.ptr(keys), _lib.ptr(flag))
        if check and int(flag.item()) != 0:
            raise IndexError(f"scatter index out of range [0, {dim_size})")
        self.keys, self.perm, self.cell_start = sort_keys(keys, B * dim_size)
        self.reso, self.shift, self.morton = 1, 0, 0
        self.xyz_sorted, self.tie = None, None
        self.B, self.N, self.dim_size = B, N, dim_size
        self.n_seg = B * dim_size
        self.n_points, self.tile_ids = n, None

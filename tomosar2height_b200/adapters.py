"""Config / IO adapters so that the B200 path can be driven from the reference's own
configuration tree and chunk files in an image without hydra / omegaconf / rasterio / laspy.

* ``load_config(conf_dir, overrides)``   Hydra-free composition of ``conf/`` (conf/config.yaml:5-10 defaults list,
  ``# @package _global_`` group files, a group file's own ``defaults: [base]``, ``${a.b}`` interpolation, the
  ``group=name`` / ``a.b.c=value`` command-line overrides of README.md:46-74).  The result reads by item and by
  attribute like the DictConfig the reference passes around (model.py:18-21).
* ``ChunkCloud``                          the point-cloud part of TomoSARDataset.__init__ (dataset.py:79-83,137-146):
  ``chunk_info.yaml`` + ``<chunk>/input_point_cloud.npz['pts']`` as float64, and the scene bounds DSMGenerator
  derives from them (generator.py:59-70).
* ``write_raster`` / ``read_raster``      a GeoTIFF-free container for the generated nDSM: float32 ``.npy`` bands +
  a JSON side-car with the affine transform (io_raster.py:56-62), pixel size and EPSG code, i.e. everything
  RasterWriter.write_to_file hands to rasterio (io_raster.py:171-212).
"""
import json
import os
import re

import numpy as np
import yaml

from .config import Config, to_config

_INTERP = re.compile(r"\$\{([^}]+)\}")


def _deep_merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _load_group_file(conf_dir, group, name):
    """One group option, its own ``defaults`` (siblings of the same group) merged first; returns (tree, is_global)."""
    path = os.path.join(conf_dir, group, name + ".yaml")
    with open(path) as fh:
        text = fh.read()
    is_global = bool(re.search(r"^#\s*@package\s+_global_\s*$", text, flags=re.M))
    node = yaml.safe_load(text) or {}
    tree = {}
    for entry in node.pop("defaults", []) or []:
        if isinstance(entry, str) and entry != "_self_":
            sub, sub_global = _load_group_file(conf_dir, group, entry)
            is_global = is_global or sub_global
            _deep_merge(tree, sub)
    _deep_merge(tree, node)
    return tree, is_global


def _lookup(root, dotted):
    node = root
    for part in dotted.split("."):
        node = node[part]
    return node


def _resolve(node, root, depth=0):
    if depth > 16:
        raise ValueError("config interpolation does not terminate")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str) and "${" in node:
        whole = _INTERP.fullmatch(node)
        if whole:  # keeps the type of the referenced value (use_footprint: ${use_footprint} -> bool)
            return _resolve(_lookup(root, whole.group(1)), root, depth + 1)
        return _resolve(_INTERP.sub(lambda m: str(_lookup(root, m.group(1))), node), root, depth + 1)
    return node


def load_config(conf_dir, overrides=()):
    """Compose ``conf_dir`` (the reference's ``conf/``) like ``@hydra.main(config_path='conf', config_name='config')``.

    ``overrides``: strings as on the reference's command line -- ``dataset=berlin`` picks a group option,
    ``use_image=true`` / ``model.encoder_kwargs.unet_kwargs.depth=4`` set values (parsed as YAML scalars)."""
    with open(os.path.join(conf_dir, "config.yaml")) as fh:
        primary = yaml.safe_load(fh) or {}
    defaults = primary.pop("defaults", []) or []
    choices, order = {}, []
    for entry in defaults:
        if entry == "_self_":
            order.append("_self_")
        elif isinstance(entry, dict):
            (group, name), = entry.items()
            if group.startswith("override "):
                continue  # hydra's own logging groups
            choices[group] = name
            order.append(group)
    values = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        key = key.lstrip("+")
        if key in choices and "." not in key:
            choices[key] = val
        else:
            values.append((key, yaml.safe_load(val)))
    if "_self_" not in order:
        order.append("_self_")  # hydra's default: the primary config is merged last
    cfg = {}
    for item in order:
        if item == "_self_":
            _deep_merge(cfg, primary)
            continue
        tree, is_global = _load_group_file(conf_dir, item, choices[item])
        _deep_merge(cfg, tree if is_global else {item: tree})
    for key, val in values:
        node = cfg
        parts = key.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = val
    cfg.pop("hydra", None)
    return to_config(_resolve(cfg, cfg))


class ChunkCloud:
    """Chunked scene cloud as the reference stores it (scripts/build_dataset.py -> dataset.py:79-83,137-146)."""

    INPUT_POINT_CLOUD = "input_point_cloud.npz"
    CHUNK_INFO = "chunk_info.yaml"

    def __init__(self, dataset_dir, chunk_ids=None):
        with open(os.path.join(dataset_dir, self.CHUNK_INFO)) as fh:
            self.chunk_info = yaml.safe_load(fh)
        self.chunk_ids = list(self.chunk_info) if chunk_ids is None else list(chunk_ids)
        self.points = {}
        for idx in self.chunk_ids:
            name = self.chunk_info[idx]["name"]
            with np.load(os.path.join(dataset_dir, name, self.INPUT_POINT_CLOUD)) as z:
                self.points[idx] = np.asarray(z["pts"], dtype=np.float64)  # geo-coordinates need float64 (dataset.py:231)

    def bounds(self):
        """generator.py:59-70: union of the chunks' (min_bound, max_bound) in x, y."""
        lo = [min(self.chunk_info[i]["min_bound"][d] for i in self.chunk_ids) for d in (0, 1)]
        hi = [max(self.chunk_info[i]["max_bound"][d] for i in self.chunk_ids) for d in (0, 1)]
        return lo, hi

    def all_points(self):
        return np.concatenate([self.points[i] for i in self.chunk_ids], 0)


def write_raster(path, bands, bl_bound, tr_bound, pixel_size, crs_epsg):
    """``bands``: (rows, cols) array or list of them (row 0 = northern edge).  Writes ``path + '.npy'`` (float32,
    (n_bands, rows, cols)) and ``path + '.json'``: the Affine(px, 0, left, 0, -py, top) of io_raster.py:56-62."""
    arr = np.stack([np.asarray(b, dtype=np.float32) for b in (bands if isinstance(bands, (list, tuple)) else [bands])])
    np.save(path + ".npy", arr)
    px = [float(p) for p in np.atleast_1d(pixel_size)] * (2 if np.ndim(pixel_size) == 0 else 1)
    meta = {"driver": "npy+json", "count": int(arr.shape[0]), "height": int(arr.shape[1]), "width": int(arr.shape[2]),
            "dtype": "float32", "crs_epsg": int(crs_epsg), "pixel_size": px[:2],
            "transform": [px[0], 0.0, float(bl_bound[0]), 0.0, -px[1], float(tr_bound[1])]}
    with open(path + ".json", "w") as fh:
        json.dump(meta, fh)
    return meta


def read_raster(path):
    with open(path + ".json") as fh:
        meta = json.load(fh)
    return np.load(path + ".npy"), meta

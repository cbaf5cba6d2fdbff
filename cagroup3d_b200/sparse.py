"""Device-resident sparse tensors and the functional sparse ops, all through the C ABI.

Host-side mirror of the slice of the MinkowskiEngine API the reference uses:
coordinate maps with hash tables, cached strided maps and rule maps (neighbour tables), and
conv / transposed conv / conv-at-coordinates / avg-pool / trilinear / quantise-average.  torch is
used for device memory and streams only; every computation is a call into libcagroup3d_b200.so.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _lib

ACT = {None: 0, "none": 0, "relu": 1, "elu": 2}

# which conv kernel gemm_rows uses: "simt" (exact fp32 FFMA) or "tc" (tcgen05, bf16 hi/lo split x 3 products, fp32 accumulate)
_CONV_IMPL = {"name": os.environ.get("CG3D_CONV", "tc")}


def set_conv_impl(name: str) -> None:
    assert name in ("simt", "tc")
    _CONV_IMPL["name"] = name


def get_conv_impl() -> str:
    return _CONV_IMPL["name"]


class LaunchCounter:
    """Counts C-ABI calls that launch kernels (bench.py reports it as gpu_launches)."""
    n = 0


class Profile:
    """Optional per-call CUDA-event timing (bench.py roofline accounting).  `active` is a list that
    receives (name, meta, start_event, end_event); None disables it."""
    active = None
    stage = ""
    conv_only = False
    streams = None          # optional list: the CUDA stream of every recorded call (tools/timeline.py)


def _call(name, *args, meta=None):
    LaunchCounter.n += 1
    if Profile.active is None or (Profile.conv_only and not name.startswith(("cg3d_spconv", "cg3d_split"))):
        _lib.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    Profile.active.append((name, Profile.stage, meta, e0, e1))
    if Profile.streams is not None:
        Profile.streams.append(torch.cuda.current_stream().cuda_stream)


def _i32(*shape, device):
    return torch.empty(shape, dtype=torch.int32, device=device)


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


@dataclass
class CoordMap:
    """Unique int32 (b,x,y,z) rows + their hash table (key -> row)."""
    coords: torch.Tensor          # [n, 4] int32
    stride: int
    keys: torch.Tensor            # [capacity] int64 (u64 bit pattern)
    vals: torch.Tensor            # [capacity] int32
    uid: int = field(default=0)
    order: Optional[torch.Tensor] = None          # tile order: position -> row (Morton-sorted), see tile_order()
    ordered_coords: Optional[torch.Tensor] = None # coords[order]

    @property
    def n(self) -> int:
        return self.coords.shape[0]

    @property
    def capacity(self) -> int:
        return self.keys.shape[0]


class Manager:
    """Caches strided maps by tensor stride (ME semantics A6) and rule maps by (in, out, kind)."""

    def __init__(self, batch_bits: int = 8):
        self.by_stride: Dict[int, CoordMap] = {}
        self.tables: Dict[Tuple, torch.Tensor] = {}
        # Coordinate stream (optional): strided maps, rule maps and tile orders depend on coordinates only, so while it
        # is set they are built on this stream -- next to the feature kernels of the calling stream instead of in
        # front of them -- and the host syncs that read a map's size wait for this stream only.  Every cached object
        # carries a CUDA event; whoever fetches it makes ITS stream wait for that event.
        self.stream: Optional["torch.cuda.Stream"] = None
        self.events: Dict[Tuple, "torch.cuda.Event"] = {}
        self.rule_counts: Dict[Tuple, int] = {}
        self.batch_bits = batch_bits          # bits of the largest batch index (bounds the tile-order sort)
        self._uid = 0

    def new_uid(self) -> int:
        self._uid += 1
        return self._uid


@dataclass
class SparseTensor:
    F: torch.Tensor
    cmap: CoordMap
    mgr: Manager

    @property
    def C(self) -> torch.Tensor:
        return self.cmap.coords

    def with_F(self, F: torch.Tensor) -> "SparseTensor":
        return SparseTensor(F, self.cmap, self.mgr)


# ---- coordinate maps --------------------------------------------------------------------------
def unique_first(coords: torch.Tensor, stride: int, mgr: Optional[Manager] = None, want_first=False,
                 want_inverse=False):
    """Hash-unique in first-occurrence order.  One host sync (reads the unique count)."""
    dev = coords.device
    n = coords.shape[0]
    cap = _lib.hash_capacity(n)
    keys = torch.empty((cap,), dtype=torch.int64, device=dev)
    vals = _i32(cap, device=dev)
    out = _i32(max(n, 1), 4, device=dev)
    first = _i32(max(n, 1), device=dev) if want_first else None
    inv = _i32(max(n, 1), device=dev) if want_inverse else None
    nu = _i32(1, device=dev)
    ws = _i32(3 * n + _lib.scan_workspace_ints(n), device=dev)
    _call("cg3d_unique_first", coords, n, keys, vals, cap, out, first, inv, nu, ws)
    u = int(nu.item())
    cm = CoordMap(out[:u], stride, keys, vals, mgr.new_uid() if mgr else 0)
    return cm, (first[:u] if want_first else None), (inv[:n] if want_inverse else None)


# BiResNet + DAPPM coordinate pyramid (biresnet.py:109-127,265-268): (tensor stride, tensor stride of the map it is made
# from).  ME derives a strided map from the rows of the conv's / pool's INPUT map, so the DAPPM pools (strides 2 .. 16 on
# the stride-32 tensor) all start from the stride-32 rows.
BACKBONE_PYRAMID = ((2, 1), (4, 2), (8, 4), (16, 8), (32, 16), (64, 32), (128, 32), (256, 32), (512, 32))


def voxel_pyramid(coords: torch.Tensor, mgr: Manager, plan=BACKBONE_PYRAMID, err: Optional[torch.Tensor] = None):
    """The stride-1 map of `coords` (hash-unique, first occurrence) AND every strided map of `plan`, with ONE host
    read-back for all their sizes: level l is made from the unique rows of its source level while that level's row count
    still lives on the device (cg3d_unique_first_dev, launches sized by the upper bound).  Replaces 1 + len(plan) size
    syncs spread over the backbone by one at its start.  -> (stride-1 map, first_row of its rows); the strided maps are
    left in mgr.by_stride, where strided_map() finds them.  err: optional device counter read back with the sizes."""
    dev, n = coords.device, coords.shape[0]
    L = len(plan)
    cap = _lib.hash_capacity(n)
    cnt = torch.zeros((L + 2,), dtype=torch.int32, device=dev)
    ws = _i32(3 * n + _lib.scan_workspace_ints(n), device=dev)
    keys0 = torch.empty((cap,), dtype=torch.int64, device=dev)
    vals0 = _i32(cap, device=dev)
    out0 = _i32(max(n, 1), 4, device=dev)
    first = _i32(max(n, 1), device=dev)
    _call("cg3d_unique_first", coords, n, keys0, vals0, cap, out0, first, None, cnt[0:1], ws)
    bufs = {1: (out0, keys0, vals0, cnt[0:1])}
    for l, (ts, src) in enumerate(plan):
        so, _, _, sn = bufs[src]
        keys = torch.empty((cap,), dtype=torch.int64, device=dev)
        vals = _i32(cap, device=dev)
        out = _i32(max(n, 1), 4, device=dev)
        _call("cg3d_unique_first_dev", so, sn, n, ts, keys, vals, cap, out, cnt[l + 1:l + 2], ws)
        bufs[ts] = (out, keys, vals, cnt[l + 1:l + 2])
    if err is not None:
        cnt[L + 1:L + 2].copy_(err)
    host = cnt.cpu().tolist()                                    # the one host sync of the backbone's coordinate maps
    maps = {}
    for i, ts in enumerate([1] + [p[0] for p in plan]):
        out, keys, vals, _ = bufs[ts]
        maps[ts] = CoordMap(out[:host[i]], ts, keys, vals, mgr.new_uid())
        mgr.by_stride[ts] = maps[ts]
    return maps[1], first[:host[0]], host[L + 1]


def build_map(coords: torch.Tensor, stride: int, mgr: Optional[Manager] = None) -> CoordMap:
    """Hash table over rows that are already unique (conv(x, coordinates=...), A12)."""
    dev = coords.device
    n = coords.shape[0]
    cap = _lib.hash_capacity(n)
    keys = torch.empty((cap,), dtype=torch.int64, device=dev)
    vals = _i32(cap, device=dev)
    _call("cg3d_hash_build", coords, n, keys, vals, cap)
    return CoordMap(coords, stride, keys, vals, mgr.new_uid() if mgr else 0)


def quantize(points4: torch.Tensor, ld: int, n: int, vs, mul: int = 1) -> torch.Tensor:
    """rows (b,x,y,z) fp32 -> int32 voxel rows; raises if a voxel index overflows the key range."""
    dev = points4.device
    out = _i32(max(n, 1), 4, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    _call("cg3d_quantize", points4, ld, n, float(vs[0]), float(vs[1]), float(vs[2]), mul, out, err)
    return out[:n], err


class _coord_scope:
    """with _coord_scope(mgr, key): body runs on the manager's coordinate stream (if any); on exit an event is recorded
    under `key`.  _fetched(mgr, key) makes the current stream wait for it."""

    def __init__(self, mgr, key):
        self.mgr, self.key = mgr, key
        self.ctx = torch.cuda.stream(mgr.stream) if (mgr is not None and mgr.stream is not None) else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            ev = torch.cuda.Event()
            ev.record(self.mgr.stream)
            self.mgr.events[self.key] = ev
            self.ctx.__exit__(*exc)
        return False


def _fetched(mgr, key):
    ev = mgr.events.get(key) if mgr is not None else None
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)


def strided_map(x_map: CoordMap, mgr: Manager, s: int) -> CoordMap:
    ts = x_map.stride * s
    if ts not in mgr.by_stride:
        with _coord_scope(mgr, ("map", ts)):
            c = _i32(max(x_map.n, 1), 4, device=x_map.coords.device)
            _call("cg3d_stride_coords", x_map.coords, x_map.n, ts, c)
            cm, _, _ = unique_first(c[:x_map.n], ts, mgr)
            mgr.by_stride[ts] = cm
    _fetched(mgr, ("map", ts))
    return mgr.by_stride[ts]


def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, n: int, end_bit: int = 64, begin_bit: int = 0) -> None:
    """in-place stable radix sort of the first n (key, value) pairs on key bits [begin_bit, end_bit)."""
    if n <= 1:
        return
    dev = keys.device
    _call("cg3d_sort_pairs", keys, vals, n, begin_bit, end_bit, torch.empty((n,), dtype=torch.int64, device=dev),
          _i32(n, device=dev), _i32(_lib.sort_workspace_ints(n), device=dev))


# which output rows share a conv tile: "none" = consecutive rows, "morton" = spatially compact patches,
# "mask" = rows with the same set of active taps (rule maps with K <= MASK_MAX_K taps; the conv skips empty taps)
_TILE_ORDER = {"mode": os.environ.get("CG3D_TILE_ORDER", "mask"), "big": os.environ.get("CG3D_TILE_ORDER_BIG", "none")}
MASK_MAX_K = 27
# maps with fewer rows keep their row order (measured 256 / 20000 / 60000: 30.25 / 30.15 / 29.76 ms per step): the three
# radix passes (~55 us of launches) cost more than the few taps
# the grouping saves on a map of a few dozen tiles
MASK_MIN_ROWS = int(os.environ.get("CG3D_MASK_MIN_ROWS", "60000"))


def mask_order(nbr: torch.Tensor, ksize: int, n: int, coords=None, group_div: int = 0, group_bits: int = 0):
    """(positional table, order) with rows grouped by (weight group, tap pattern)."""
    K, dev = nbr.shape[0], nbr.device
    keys = torch.empty((n,), dtype=torch.int64, device=dev)
    order = _i32(n, device=dev)
    _call("cg3d_table_mask_keys", nbr, K, ksize, n, coords if group_div else None, group_div, keys, order)
    # 24 key bits (3 radix passes) group almost as well as all 27 taps
    sort_pairs(keys, order, n, end_bit=(min(K, 27) if not group_div else 27 + group_bits), begin_bit=max(0, min(K, 27) - 24))
    out = torch.empty_like(nbr)
    _call("cg3d_permute_table", nbr, K, n, order, out)
    return out, order


def tile_order(cmap: CoordMap, batch_bits: int = 8):
    """(order, coords[order]): Morton tile order of a map (cached on it); (None, coords) when disabled.

    The map's own row order (ME's first-occurrence order) is untouched: `order` only decides which output
    rows share a conv tile.  Cells of 4^3 voxels are the sort granularity (key bits below 6 are ignored)."""
    if cmap.n < 256:
        return None, cmap.coords
    if cmap.order is None:
        dev, n = cmap.coords.device, cmap.n
        keys = torch.empty((n,), dtype=torch.int64, device=dev)
        vals = _i32(n, device=dev)
        _call("cg3d_morton_keys", cmap.coords, n, cmap.stride, keys, vals)
        sort_pairs(keys, vals, n, end_bit=33 + batch_bits, begin_bit=6)
        oc = _i32(n, 4, device=dev)
        _call("cg3d_gather_coords", cmap.coords, vals, n, oc)
        cmap.order, cmap.ordered_coords = vals, oc
    return cmap.order, cmap.ordered_coords


def neighbor_table(in_map: CoordMap, out_map: CoordMap, k: int, mgr: Optional[Manager], ordered: bool = False,
                   group_div: int = 0, spatial: Optional[bool] = None, coarse_mask: bool = False, wait: bool = True):
    """ME kernel map as a tap-major table.  ordered=False -> nbr[k][row]; ordered=True -> (nbr[k][position], order)
    with position -> row given by the output map's tile order (order is None when positions == rows).

    Tile order: k^3 <= 27 taps -> rows grouped by tap pattern ("mask"); wider kernels (the 9^3 / 5^3 class convs)
    keep the map's row order: their maps are ~13 % occupied volumes, where even a Morton-compact 128-row tile reaches
    ~87 % of the 729 taps (measured: 634 active taps per tile either way, profiles/r1_stage_times_tc.log), so the
    sort does not pay (CG3D_TILE_ORDER_BIG=morton enables it).  spatial=False forces the map's own row order.
    coarse_mask=True (wide kernel over a map where a row has only a handful of neighbours, i.e. the RoI grid conv with
    3.7 of 125): rows are grouped by WHICH of the 3x3x3 coarse blocks of taps they reach, so a tile touches a few
    blocks' taps instead of 89 of 125, and rows without any neighbour form tiles without work."""
    key = ("conv", in_map.uid, out_map.uid, k, ordered)
    if mgr is not None and key in mgr.tables:
        if wait:
            _fetched(mgr, key)
        return mgr.tables[key]
    with _coord_scope(mgr, key):
        res = _build_neighbor_table(in_map, out_map, k, mgr, ordered, group_div, spatial, coarse_mask)
    if mgr is not None:
        mgr.tables[key] = res
    if wait:                    # wait=False: prefetch on the coordinate stream; the later (cached) fetch waits
        _fetched(mgr, key)
    return res


def _build_neighbor_table(in_map, out_map, k, mgr, ordered, group_div, spatial, coarse_mask):
    if spatial is None:
        spatial = _TILE_ORDER["mode"] == "mask" and k ** 3 > MASK_MAX_K and _TILE_ORDER["big"] == "morton"
    use = ordered and (_TILE_ORDER["mode"] == "morton" or spatial)
    order, oc = tile_order(out_map, mgr.batch_bits if mgr else 8) if use else (None, out_map.coords)
    nbr = _i32(k ** 3, max(out_map.n, 1), device=in_map.coords.device)
    if in_map is out_map and oc is out_map.coords and (k & 1) and k > 1 and out_map.n > 0:
        _call("cg3d_neighbor_table_symmetric", oc, out_map.n, in_map.keys, in_map.vals, in_map.capacity, k, in_map.stride, nbr)
    else:
        _call("cg3d_neighbor_table", oc, out_map.n, in_map.keys, in_map.vals, in_map.capacity, k, in_map.stride, nbr)
    if ordered and _TILE_ORDER["mode"] == "mask" and (k ** 3 <= MASK_MAX_K or coarse_mask) and out_map.n >= MASK_MIN_ROWS:
        nbr, order = mask_order(nbr, k, out_map.n, out_map.coords, group_div, mgr.batch_bits if mgr else 8)
    return (nbr, order) if ordered else nbr


def transpose_table(in_map: CoordMap, fine_map: CoordMap, k: int, mgr: Optional[Manager], ordered: bool = False,
                    group_div: int = 0):
    key = ("convT", in_map.uid, fine_map.uid, k, ordered)
    if mgr is not None and key in mgr.tables:
        _fetched(mgr, key)
        return mgr.tables[key]
    with _coord_scope(mgr, key):
        res = _build_transpose_table(in_map, fine_map, k, mgr, ordered, group_div)
    if mgr is not None:
        mgr.tables[key] = res
    _fetched(mgr, key)
    return res


def _build_transpose_table(in_map, fine_map, k, mgr, ordered, group_div):
    use = ordered and _TILE_ORDER["mode"] == "morton"
    order, oc = tile_order(fine_map, mgr.batch_bits if mgr else 8) if use else (None, fine_map.coords)
    nbr = _i32(k ** 3, max(fine_map.n, 1), device=in_map.coords.device)
    _call("cg3d_transpose_table", oc, fine_map.n, in_map.keys, in_map.vals, in_map.capacity, k, in_map.stride, nbr)
    if ordered and _TILE_ORDER["mode"] == "mask" and k ** 3 <= MASK_MAX_K and fine_map.n >= MASK_MIN_ROWS:
        nbr, order = mask_order(nbr, k, fine_map.n, fine_map.coords, group_div, mgr.batch_bits if mgr else 8)
    return (nbr, order) if ordered else nbr


def count_rules(nbr: torch.Tensor) -> int:
    cnt = torch.zeros((1,), dtype=torch.int64, device=nbr.device)
    _call("cg3d_count_rules", nbr, nbr.numel(), cnt)
    return int(cnt.item())


# ---- compute ----------------------------------------------------------------------------------
@dataclass
class Tiles:
    """Grouped-conv tiling: rows [row0, row0+rows) of tile t use weight group `group`."""
    row0: torch.Tensor
    rows: torch.Tensor
    group: torch.Tensor
    n: int
    offsets: Optional[list] = None        # the per-group row ranges the tiling was cut from
    tile: int = 0
    alt: Dict[int, "Tiles"] = field(default_factory=dict)

    def with_tile(self, tile: int) -> "Tiles":
        """the same row ranges cut into tiles of another height (the pair-compacted kernel's 448 rows); cached"""
        if tile == self.tile:
            return self
        if tile not in self.alt:
            self.alt[tile] = make_tiles(self.offsets, self.row0.device, tile)
        return self.alt[tile]


def make_tiles(seg_offsets, device, tile=64) -> Tiles:
    """seg_offsets: python list [g0_start, g1_start, ..., total] of contiguous per-group row ranges."""
    import numpy as np
    off = np.asarray(seg_offsets, dtype=np.int64)
    cnt = np.maximum(-(-(off[1:] - off[:-1]) // tile), 0)                     # tiles per group (vectorised: the head's maps
    gg = np.repeat(np.arange(len(cnt)), cnt)                                   # have thousands of tiles)
    first = np.cumsum(cnt) - cnt
    r0 = off[:-1][gg] + (np.arange(int(cnt.sum())) - first[gg]) * tile
    rn = np.minimum(tile, off[1:][gg] - r0)
    t = torch.from_numpy(np.stack([r0, rn, gg]).astype(np.int32))
    if torch.device(device).type == "cuda":
        # pinned staging: a copy from pageable memory first waits for everything queued on the stream (the host would sit
        # out the kernels launched just before and the GPU would idle after them while the next launches are prepared)
        t = t.pin_memory()
    t = t.to(device, non_blocking=True)
    return Tiles(t[0].contiguous(), t[1].contiguous(), t[2].contiguous(), int(cnt.sum()), list(seg_offsets), tile)


def gemm_rows(Fin: torch.Tensor, nbr: Optional[torch.Tensor], W: torch.Tensor, n_out: int, K: int,
              scale=None, shift=None, residual=None, act=None, tiles: Optional[Tiles] = None,
              impl: Optional[str] = None, in_act=None, out: Optional[torch.Tensor] = None,
              out_rows: Optional[torch.Tensor] = None, split_out: Optional[str] = None,
              algo_cin: Optional[int] = None) -> torch.Tensor:
    """out = act((sum_k in_act(Fin[nbr[k]]) @ W[k]) * scale + shift + residual); W: [(G,) K, Cin, Cout].
    algo_cin: the layer's real input width when Fin / W are zero-padded (only the profiling record uses it: algorithmic
    bytes and FLOPs are those of the unpadded layer).

    Fin / out may be column slices of wider row-major matrices (unit column stride).  out_rows: the table is
    positional (tile order); position j is output row out_rows[j].  split_out ("none" | "relu"): the tensor-core
    kernel also emits the split-bf16 copy of the result a following conv with that input activation will ask for."""
    Cin, Cout = W.shape[-2], W.shape[-1]
    assert Fin.stride(1) == 1 and W.is_contiguous() and Fin.shape[1] == Cin
    if out is None:
        out = _f32(n_out, Cout, device=Fin.device)
    assert out.stride(1) == 1 and out.shape[0] == n_out and out.shape[1] == Cout
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == (n_out, Cout)
    if n_out == 0:
        return out
    name = impl or _CONV_IMPL["name"]
    use_tc = (name == "tc" and tc_supported(Cin, Cout, K) and in_act in (None, "none", "relu")
              and Fin.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0 and Fin.stride(0) % 4 == 0
              and out.stride(0) % 4 == 0)
    meta = None
    if Profile.active is not None and not Profile.conv_only:
        ac = algo_cin or Cin
        meta = dict(n_in=Fin.shape[0], n_out=n_out, Cin=ac, Cout=Cout, K=K, nbr=nbr, w_bytes=W.numel() * 4 * ac // Cin,
                    residual=residual is not None)
    targs = (tiles.row0 if tiles else None, tiles.rows if tiles else None, tiles.group if tiles else None,
             tiles.n if tiles else 0, out_rows)
    if use_tc:
        So = None
        if split_out is not None and Cout % 32 == 0 and out.stride(0) == Cout:
            So = torch.empty((n_out, 2 * Cout), dtype=torch.int16, device=out.device)
            out._cg3d_split = {(out.data_ptr(), out._version, out.stride(0), 1 if split_out == "relu" else 0): So}
        ks = _lib.host("cg3d_spconv_tc_splitk", n_out, Cin, Cout, K, 1 if tiles else 0, tiles.n if tiles else 0)
        ws = torch.empty((ks * n_out * Cout,), dtype=torch.float32, device=out.device) if ks > 1 else None
        _call("cg3d_spconv_tc", split_rows(Fin, in_act), Fin.shape[0], nbr, weight_image(W), out, out.stride(0), n_out, Cin, Cout, K,
              scale, shift, residual, ACT[act], *targs, So, 1 if split_out == "relu" else 0, ws, meta=meta)
    else:
        _call("cg3d_spconv_simt", Fin, Fin.stride(0), ACT[in_act], nbr, W, out, out.stride(0), n_out, Cin, Cout, K,
              scale, shift, residual, ACT[act], *targs, meta=meta)
    return out


def split_rows(F: torch.Tensor, in_act=None) -> torch.Tensor:
    """[n, 2C] bf16 copy (per 32-channel chunk: hi | lo) of an activation matrix for the tensor-core conv, with the
    consumer's input activation applied.  Cached on the tensor object; rebuilt if the tensor is modified in place."""
    relu = 1 if in_act == "relu" else 0
    assert in_act in (None, "none", "relu")
    cache = getattr(F, "_cg3d_split", None)
    if cache is None:
        cache = {}
        try:
            F._cg3d_split = cache
        except AttributeError:
            pass
    key = (F.data_ptr(), F._version, F.stride(0), relu)
    if key not in cache:
        n, C = F.shape
        out = torch.empty((max(n, 1), 2 * C), dtype=torch.int16, device=F.device)
        _call("cg3d_split_bf16", F, F.stride(0), n, C, relu, out)
        if len(cache) > 1:
            cache.clear()
        cache[key] = out
    return cache[key]


def tc_supported(Cin: int, Cout: int, K: int = 1) -> bool:
    return Cin % 32 == 0 and Cout % 64 == 0 and K <= 729


def weight_image(W: torch.Tensor) -> torch.Tensor:
    """bf16 hi/lo split + UMMA-swizzled image of a weight tensor.  Built once and kept ON the tensor
    object (so it dies with it); rebuilt when the tensor is modified in place or moved."""
    cached = getattr(W, "_cg3d_wimg", None)
    if cached is not None and cached[0] == (W.data_ptr(), W._version):
        return cached[1]
    Cin, Cout = W.shape[-2], W.shape[-1]
    K = W.shape[-3] if W.dim() >= 3 else 1
    G = W.shape[0] if W.dim() == 4 else 1
    img = torch.empty((W.numel() * 4,), dtype=torch.uint8, device=W.device)
    _call("cg3d_spconv_tc_prepare", W.detach(), G, K, Cin, Cout, img)
    W._cg3d_wimg = ((W.data_ptr(), W._version), img)
    return img


def conv(x: SparseTensor, W: torch.Tensor, k: int, stride: int = 1, **ep) -> SparseTensor:
    """MinkowskiConvolution forward with a fused epilogue (scale/shift/residual/act)."""
    if k == 1 and stride == 1:
        return x.with_F(gemm_rows(x.F, None, W, x.cmap.n, 1, **ep))
    omap = x.cmap if stride == 1 else strided_map(x.cmap, x.mgr, stride)
    nbr, order = neighbor_table(x.cmap, omap, k, x.mgr, ordered=True)
    return SparseTensor(gemm_rows(x.F, nbr, W, omap.n, k ** 3, out_rows=order, **ep), omap, x.mgr)


def conv_transpose_k2s2(x: SparseTensor, W: torch.Tensor, **ep) -> SparseTensor:
    fmap = x.mgr.by_stride[x.cmap.stride // 2]
    nbr, order = transpose_table(x.cmap, fmap, 2, x.mgr, ordered=True)
    return SparseTensor(gemm_rows(x.F, nbr, W, fmap.n, 8, out_rows=order, **ep), fmap, x.mgr)


def affine_act(x: torch.Tensor, scale=None, shift=None, add=None, act=None, out=None) -> torch.Tensor:
    """act(x * scale + shift + add); x / out may be column slices (unit column stride)."""
    if out is None:
        out = torch.empty((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    assert x.stride(1) == 1 and out.stride(1) == 1
    if add is not None:
        assert add.is_contiguous()
    _call("cg3d_affine_act", x, x.stride(0), scale, shift, add, out, out.stride(0), x.shape[0], x.shape[1], ACT[act])
    return out


def relu_rows(F: torch.Tensor) -> torch.Tensor:
    """relu(F) as a new matrix; a cached split copy of F made for a relu-consuming conv is the plain split of the result."""
    out = affine_act(F, act="relu")
    for (ptr, ver, ld, relu), Sp in (getattr(F, "_cg3d_split", None) or {}).items():
        if relu == 1 and ptr == F.data_ptr() and ver == F._version and ld == F.stride(0):
            out._cg3d_split = {(out.data_ptr(), out._version, out.stride(0), 0): Sp}
    return out


def interp(x: SparseTensor, query: torch.Tensor, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """base + x.features_at_coordinates(query) for integer query rows."""
    nq, C = query.shape[0], x.F.shape[1]
    out = _f32(nq, C, device=x.F.device)
    _call("cg3d_interp_trilinear", query, nq, x.cmap.keys, x.cmap.vals, x.cmap.capacity, x.cmap.stride, x.F, C,
          base, out)
    return out


def avg_pool(x: SparseTensor, k: int, stride: int) -> SparseTensor:
    omap = strided_map(x.cmap, x.mgr, stride)
    C = x.F.shape[1]
    out = _f32(omap.n, C, device=x.F.device)
    _call("cg3d_avgpool_window", omap.coords, omap.n, x.cmap.coords, x.cmap.n, (k // 2) * x.cmap.stride, x.F, C, out)
    return SparseTensor(out, omap, x.mgr)


def segment_mean(srcA, ldA, srcB, ldB, ref, inverse, n, n_unique, C) -> torch.Tensor:
    dev = inverse.device
    out = _f32(n_unique, C, device=dev)
    cnt = _f32(max(n_unique, 1), device=dev)
    ws = torch.empty((_lib.host("cg3d_segment_mean_workspace", n, n_unique),), dtype=torch.int64, device=dev)
    _call("cg3d_segment_mean", srcA, ldA, srcB, ldB, ref, inverse, n, n_unique, C, out, cnt, ws)
    return out


def gather_rows(src: torch.Tensor, col0: int, rows: Optional[torch.Tensor], n: int, C: int, divisor: float = 1.0):
    out = _f32(n, C, device=src.device)
    _call("cg3d_gather_rows", src, src.shape[1], col0, rows, n, C, divisor, out)
    return out


def exclusive_scan(flags: torch.Tensor):
    """returns (positions, total as device int tensor)."""
    n = flags.numel()
    out = _i32(max(n, 1), device=flags.device)
    ws = _i32(_lib.scan_workspace_ints(n), device=flags.device)
    total = _i32(1, device=flags.device)
    _call("cg3d_exclusive_scan_i32", flags, n, out, ws, total)
    return out[:n], total

"""Arrow -> ragged (pinned) buffers: the host side of the grouped-map boundary.

Replaces Spark's shuffle + Arrow hand-off of ``groupby('series_id','dim_id').apply``
(reference src/jobs/prophet_modeler.py:139-141): rows are sorted by (series_id, dim_id, ds)
-- the sort fbprophet's setup_dataframe does per group -- and cut into CSR offsets.
Null ``y`` rows are dropped from the fit buffers (fbprophet: ``df[df['y'].notnull()]``) but
still count for ``history_dates.max()`` (the anchor of make_future_dataframe).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


@dataclass
class PackedGroups:
    series_id: np.ndarray    # [N] int32
    dim_id: np.ndarray       # [N] int32
    offsets: np.ndarray      # [N+1] int64 into ds / y
    ds: np.ndarray           # [R] int64 ns, ascending within a group
    y: np.ndarray            # [R] int32 (or float64 when the input was not integral)
    last_ds: np.ndarray      # [N] int64: max ds of the group INCLUDING null-y rows
    n_rows_in: np.ndarray    # [N] rows of the group before dropping nulls

    @property
    def n(self) -> int:
        return self.offsets.size - 1

    def take(self, lo: int, hi: int) -> "PackedGroups":
        """Groups [lo, hi) (a rank's shard; see dist.shard_bounds)."""
        a, b = int(self.offsets[lo]), int(self.offsets[hi])
        return PackedGroups(self.series_id[lo:hi], self.dim_id[lo:hi], self.offsets[lo:hi + 1] - a,
                            self.ds[a:b], self.y[a:b], self.last_ds[lo:hi], self.n_rows_in[lo:hi])

    @property
    def on_device(self) -> bool:
        return not isinstance(self.ds, np.ndarray)


def _pinned_like(a: np.ndarray) -> np.ndarray:
    """Copy into page-locked memory when a CUDA runtime is usable (faster H2D), else return as is."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
            out = t.numpy()
            out[...] = a
            return _Keep(out, t)
    except Exception:
        pass
    return a


def _group_keys(table: pa.Table, keys):
    """The two int group-key columns as int64 numpy arrays.  A null key is refused: Spark would make null its own
    group, and the reference's fixed schema never produces one (series_id comes from the directory name, an empty
    dim_id field is a malformed row) -- silently merging it into another group would be worse than failing."""
    out = []
    for k in keys[:2]:
        col = table[k]
        if col.null_count:
            raise ValueError(f"group key column {k!r} holds {col.null_count} null value(s); every row needs both "
                             f"{keys[0]!r} and {keys[1]!r}")
        out.append(np.asarray(col.combine_chunks().to_numpy(zero_copy_only=False)).astype(np.int64))
    return out


class _Keep(np.ndarray):
    """ndarray view that keeps the owning pinned torch tensor alive."""
    def __new__(cls, arr, owner):
        obj = np.asarray(arr).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def pack_groups(table: pa.Table, keys=("series_id", "dim_id"), ds_col="ds", y_col="y", pin: bool = True) -> PackedGroups:
    if table.num_rows == 0:
        z = np.zeros(0, np.int32)
        return PackedGroups(z, z, np.zeros(1, np.int64), np.zeros(0, np.int64), z, np.zeros(0, np.int64), np.zeros(0, np.int64))
    ds_arr = table[ds_col]
    if pa.types.is_timestamp(ds_arr.type):
        ds_arr = pc.cast(ds_arr, pa.timestamp("ns"))
        if ds_arr.null_count:
            raise ValueError("Found NaN in column ds.")       # fbprophet setup_dataframe
        ds_np = pc.cast(ds_arr, pa.int64()).to_numpy() if isinstance(ds_arr, pa.Array) else \
            pc.cast(ds_arr, pa.int64()).combine_chunks().to_numpy()
    else:
        ds_np = ds_arr.to_numpy() if isinstance(ds_arr, pa.Array) else ds_arr.combine_chunks().to_numpy()
        ds_np = np.asarray(ds_np, dtype=np.int64)
    k0, k1 = _group_keys(table, keys)
    ycol = table[y_col].combine_chunks()
    y_null = np.asarray(ycol.is_null().to_numpy(zero_copy_only=False)) if ycol.null_count else None
    if pa.types.is_integer(ycol.type):
        y_np = np.asarray(ycol.fill_null(0).to_numpy(zero_copy_only=False)).astype(np.int32)
    else:
        y_np = np.asarray(ycol.to_numpy(zero_copy_only=False)).astype(np.float64)
        nanmask = np.isnan(y_np)
        if nanmask.any():
            y_null = nanmask if y_null is None else (y_null | nanmask)
    order = np.lexsort((ds_np, k1, k0))            # stable: by series_id, dim_id, ds
    k0, k1, ds_np, y_np = k0[order], k1[order], ds_np[order], y_np[order]
    if y_null is not None:
        y_null = y_null[order]
    new_grp = np.empty(k0.size, dtype=bool)
    new_grp[0] = True
    np.logical_or(k0[1:] != k0[:-1], k1[1:] != k1[:-1], out=new_grp[1:])
    starts = np.flatnonzero(new_grp)
    ends = np.append(starts[1:], k0.size)
    last_ds = ds_np[ends - 1]
    n_rows_in = (ends - starts).astype(np.int64)
    sid, did = k0[starts].astype(np.int32), k1[starts].astype(np.int32)
    if y_null is not None and y_null.any():
        keep = ~y_null
        grp_id = np.cumsum(new_grp) - 1
        counts = np.bincount(grp_id[keep], minlength=starts.size).astype(np.int64)
        ds_np, y_np = ds_np[keep], y_np[keep]
    else:
        counts = n_rows_in
    offsets = np.zeros(starts.size + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    ds_np, y_np = np.ascontiguousarray(ds_np), np.ascontiguousarray(y_np)
    if pin:
        ds_np, y_np = _pinned_like(ds_np), _pinned_like(y_np)
    return PackedGroups(sid, did, offsets, ds_np, y_np, last_ds.astype(np.int64), n_rows_in)


def pack_groups_cuda(table: pa.Table, device=None, keys=("series_id", "dim_id"), ds_col="ds", y_col="y"):
    """GPU version of :func:`pack_groups`: the (series_id, dim_id, ds) sort of the whole
    frame runs on the B200 as two stable radix sorts (``torch.sort`` -- plumbing, not a hand-written
    kernel) and ``ds`` / ``y`` stay resident in HBM for ``pb200_fit_device``.  Returns a PackedGroups
    whose ``ds`` / ``y`` are CUDA tensors; ids, offsets and last_ds are host numpy arrays.
    Null ``y`` rows are dropped exactly as on the host path."""
    import torch
    if table.num_rows == 0:
        return pack_groups(table, keys, ds_col, y_col, pin=False)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ds_arr = table[ds_col]
    if pa.types.is_timestamp(ds_arr.type):
        if ds_arr.null_count:
            raise ValueError("Found NaN in column ds.")
        ds_arr = pc.cast(pc.cast(ds_arr, pa.timestamp("ns")), pa.int64())
    ds_np = np.asarray(ds_arr.combine_chunks().to_numpy(zero_copy_only=False), dtype=np.int64)
    k0, k1 = _group_keys(table, keys)
    ycol = table[y_col].combine_chunks()
    integral = pa.types.is_integer(ycol.type)
    if integral:
        y_null = np.asarray(ycol.is_null().to_numpy(zero_copy_only=False)) if ycol.null_count else None
        y_np = np.asarray(ycol.fill_null(0).to_numpy(zero_copy_only=False)).astype(np.int32)
    else:
        y_np = np.asarray(ycol.to_numpy(zero_copy_only=False)).astype(np.float64)
        y_null = np.isnan(y_np) if np.isnan(y_np).any() else None
    # one signed 64-bit sort key: series_id in the high word, dim_id biased by 2^31 in the low word, so that the
    # signed order of the key is the (series_id, dim_id) order of the host path's lexsort for negative ids too
    key = torch.from_numpy(((k0 << 32) | ((k1 + (1 << 31)) & 0xFFFFFFFF)).copy()).to(dev)
    ds_t = torch.from_numpy(np.array(ds_np, copy=True)).to(dev)      # (Arrow buffers are read-only: copy before wrapping)
    y_t = torch.from_numpy(np.array(y_np, copy=True)).to(dev)
    i1 = torch.argsort(ds_t, stable=True)
    i2 = torch.argsort(key[i1], stable=True)
    order = i1[i2]
    key, ds_t, y_t = key[order], ds_t[order], y_t[order]
    new_grp = torch.ones(key.numel(), dtype=torch.bool, device=dev)
    new_grp[1:] = key[1:] != key[:-1]
    starts = torch.nonzero(new_grp).flatten()
    ends = torch.cat((starts[1:], torch.tensor([key.numel()], device=dev)))
    last_ds = ds_t[ends - 1].cpu().numpy()
    n_rows_in = (ends - starts).cpu().numpy().astype(np.int64)
    gkeys = key[starts].cpu().numpy()
    sid, did = (gkeys >> 32).astype(np.int32), ((gkeys & 0xFFFFFFFF) - (1 << 31)).astype(np.int32)
    if y_null is not None and y_null.any():
        keep = ~torch.from_numpy(y_null).to(dev)[order]
        grp_id = torch.cumsum(new_grp.to(torch.int64), 0) - 1
        counts = torch.bincount(grp_id[keep], minlength=starts.numel()).cpu().numpy().astype(np.int64)
        ds_t, y_t = ds_t[keep].contiguous(), y_t[keep].contiguous()
    else:
        counts = n_rows_in
    offsets = np.zeros(sid.size + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return PackedGroups(sid, did, offsets, ds_t.contiguous(), y_t.contiguous(), last_ds.astype(np.int64), n_rows_in)

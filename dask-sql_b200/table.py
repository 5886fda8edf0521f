"""Device tables: the registered relations behind Context.create_table.

A table is a list of partitions; a partition maps column name -> DeviceColumn (HBM-resident,
`persist=True`) or HostColumn (pinned host memory, streamed to the GPU per query, the
counterpart of dask's lazy partitions when `persist=False`, input_utils/convert.py:70-71).
In a multi-rank job (one process per GPU) every rank holds its own shard of a `sharded`
table; `replicated` tables hold the same rows on every rank; `root` tables hold rows on rank 0
only and are broadcast over NCCL when a join needs them as its build side.
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from .device import DeviceColumn, Stats, column_from_host, I64, F64, U8, _pack_valid


class HostColumn:
    """Pinned host copy of a column; .to_device() issues the (async) H2D copy."""

    __slots__ = ("data", "valid", "dtype", "logical", "n", "stats")

    def __init__(self, data: torch.Tensor, valid: Optional[torch.Tensor], dtype, logical):
        self.data, self.valid, self.dtype, self.logical = data, valid, dtype, logical
        self.n = int(data.shape[0])
        self.stats = None

    def to_device(self, device) -> DeviceColumn:
        d = self.data.to(device, non_blocking=True)
        v = self.valid.to(device, non_blocking=True) if self.valid is not None else None
        col = DeviceColumn(d, v, self.dtype, self.logical, self.stats)
        return col

    def nbytes(self):
        return self.data.numel() * self.data.element_size() + (0 if self.valid is None else self.valid.numel() * 4)


def _host_column(values, pin=True) -> HostColumn:
    import pandas as pd

    logical = str(getattr(values, "dtype", "float64"))
    mask = None
    if isinstance(values, pd.Series):
        arr = values.array
        if isinstance(arr, pd.arrays.BooleanArray):
            mask, vals = np.asarray(arr._mask), np.asarray(arr._data).astype(np.uint8)
        elif isinstance(arr, (pd.arrays.IntegerArray, pd.arrays.FloatingArray)):
            mask, vals = np.asarray(arr._mask), np.asarray(arr._data)
        else:
            vals = values.to_numpy()
    elif isinstance(values, torch.Tensor):
        vals = values
    else:
        vals = np.asarray(values)
    if isinstance(vals, torch.Tensor):
        t = vals
        dt = {torch.int64: I64, torch.float64: F64, torch.uint8: U8, torch.bool: U8}[t.dtype]
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        logical = {I64: "int64", F64: "float64", U8: "bool"}[dt]
    else:
        kind = vals.dtype.kind
        if kind == "b":
            vals, dt = vals.astype(np.uint8), U8
        elif kind in "iu" and not (kind == "u" and vals.dtype.itemsize == 8):
            vals, dt = vals.astype(np.int64, copy=False), I64
        elif kind == "f":
            vals, dt = vals.astype(np.float64, copy=False), F64
        else:
            raise NotImplementedError(
                f"column dtype {logical} is outside the int64/float64/bool hot path of the B200 layer")
        import warnings
        with warnings.catch_warnings():
            # pandas hands out read-only views; the tensor is only ever read (H2D source), so the
            # zero-copy view is what we want
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(np.ascontiguousarray(vals))
    if pin and torch.cuda.is_available() and not t.is_pinned():
        t = t.pin_memory()
    v = None
    if mask is not None and mask.any():
        v = torch.from_numpy(_pack_valid(mask))
        if pin and torch.cuda.is_available():
            v = v.pin_memory()
    return HostColumn(t, v, dt, logical)


class DeviceTable:
    def __init__(self, partitions: List[Dict[str, object]], distribution="local", name=None):
        self.partitions = partitions
        self.distribution = distribution
        self.name = name
        self._schema = None

    def schema(self):
        if self._schema is None:
            p0 = self.partitions[0]
            self._schema = [(n, c.dtype, c.logical) for n, c in p0.items()]
        return self._schema

    @property
    def columns(self):
        return [n for n, _, _ in self.schema()]

    @property
    def nrows(self):
        return sum(next(iter(p.values())).n if p else 0 for p in self.partitions)

    def nbytes(self):
        return sum(c.nbytes() for p in self.partitions for c in p.values())

    def is_resident(self):
        return all(isinstance(c, DeviceColumn) for p in self.partitions for c in p.values())

    def column_stats(self, name) -> Stats:
        """Table-level statistics of one column (min/max/nulls), combining partitions.
        Computed on the GPU the first time a plan needs them, then cached."""
        mn = mx = None
        nulls = 0
        for p in self.partitions:
            c = p[name]
            if c.stats is None:
                if isinstance(c, HostColumn):
                    dev = torch.device("cuda", torch.cuda.current_device())
                    c.stats = c.to_device(dev).ensure_stats()
                else:
                    c.ensure_stats()
            st = c.stats
            nulls += st.nulls
            if st.vmin is not None:
                mn = st.vmin if mn is None else min(mn, st.vmin)
                mx = st.vmax if mx is None else max(mx, st.vmax)
        return Stats(mn, mx, nulls)

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_columns(cls, columns: Dict[str, object], npartitions=1, device=None, persist=True,
                     distribution="local", name=None):
        """columns: name -> pandas Series / numpy array / torch tensor (host or device)."""
        names = list(columns)
        n = len(next(iter(columns.values()))) if names else 0
        npartitions = max(1, min(int(npartitions), max(1, n)))
        # partition boundaries on multiples of 32 rows so validity bitmaps split on word boundaries
        step = -(-n // npartitions)
        step = max(32, (step + 31) // 32 * 32)
        bounds = [(lo, min(n, lo + step)) for lo in range(0, max(n, 1), step)]
        if n == 0:
            bounds = [(0, 0)]
        parts = []
        for lo, hi in bounds:
            part = {}
            for nm in names:
                v = columns[nm]
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    dt = {torch.int64: I64, torch.float64: F64, torch.uint8: U8}[v.dtype]
                    part[nm] = DeviceColumn(v[lo:hi], None, dt)
                    continue
                piece = v.iloc[lo:hi] if hasattr(v, "iloc") else v[lo:hi]
                hc = _host_column(piece, pin=not persist)
                part[nm] = hc.to_device(device) if persist else hc
            parts.append(part)
        return cls(parts, distribution, name)

    @classmethod
    def from_pandas(cls, df, npartitions=1, device=None, persist=True, distribution="local", name=None):
        return cls.from_columns({str(c): df[c] for c in df.columns}, npartitions, device, persist,
                                distribution, name)

"""Return-type policy of the drop-in API.

The reference returns Python lists / dicts / strs.  Up to MATERIALIZE_MAX
elements we hand back genuine ``list`` / ``dict`` objects so results compare
``==`` with the reference's.  Above it, lazy views over the device tensors that
implement ``__len__`` / ``__getitem__`` / ``__iter__`` / ``__eq__`` / ``tolist()``:
a 200 MB text has 2*10^8 suffix-array entries and its dense Occ table would be
n * sigma integers (utils/utils.py:28-31), neither of which can exist as Python
objects.
"""
from __future__ import annotations

import os
from collections.abc import Mapping, Sequence

import numpy as np
import torch

MATERIALIZE_MAX = int(os.environ.get("HKCSA_MATERIALIZE_MAX", str(1 << 22)))
_CHUNK = 1 << 22


class DeviceSequence(Sequence):
    """Read-only integer sequence backed by a device tensor (e.g. the suffix array)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.numel()

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.tensor[i].cpu().tolist()
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        return int(self.tensor[i].item())

    def __iter__(self):
        for s in range(0, len(self), _CHUNK):
            yield from self.tensor[s:s + _CHUNK].cpu().tolist()

    def tolist(self):
        return self.tensor.cpu().tolist()

    def numpy(self):
        from . import engine
        return engine.to_host(self.tensor)

    def __eq__(self, other):
        if isinstance(other, DeviceSequence):
            return len(self) == len(other) and bool(torch.equal(self.tensor.to(torch.int64),
                                                                other.tensor.to(torch.int64)))
        if isinstance(other, (list, tuple, np.ndarray)):
            if len(other) != len(self):
                return False
            for s in range(0, len(self), _CHUNK):
                a = self.tensor[s:s + _CHUNK].cpu().numpy().astype(np.int64)
                b = np.asarray(other[s:s + _CHUNK], dtype=np.int64)
                if not np.array_equal(a, b):
                    return False
            return True
        return NotImplemented

    def __repr__(self):
        head = self.tensor[:8].cpu().tolist()
        return f"DeviceSequence(len={len(self)}, head={head})"


def int_sequence(tensor: torch.Tensor):
    """list[int] when small enough, otherwise a DeviceSequence."""
    if tensor.numel() <= MATERIALIZE_MAX:
        return tensor.cpu().tolist()
    return DeviceSequence(tensor)


def as_device_i32(seq, device) -> torch.Tensor:
    """list / ndarray / DeviceSequence / tensor of suffix-array entries -> int32 device tensor."""
    if isinstance(seq, DeviceSequence):
        return seq.tensor.to(device=device, dtype=torch.int32)
    if isinstance(seq, torch.Tensor):
        return seq.to(device=device, dtype=torch.int32).contiguous()
    arr = np.asarray(seq, dtype=np.int64)
    if arr.ndim != 1:
        raise TypeError("suffix array must be one-dimensional")
    return torch.from_numpy(arr).to(device).to(torch.int32)


class OccColumn(Sequence):
    """occ[c] of build_occ (utils/utils.py:26-32): occ[c][i] = #c in bwt[0:i], i in [0, n]."""

    def __init__(self, wt, byte: int):
        self._wt = wt
        self._byte = byte

    def __len__(self):
        return self._wt.n + 1

    def __getitem__(self, i):
        n1 = len(self)
        if isinstance(i, slice):
            idx = np.arange(*i.indices(n1), dtype=np.int64)
            if idx.size == 0:
                return []
            sym = np.full(idx.size, self._byte, dtype=np.uint8)
            return self._wt.rank(sym, idx).cpu().tolist()
        if i < 0:
            i += n1
        if not 0 <= i < n1:
            raise IndexError("list index out of range")
        return int(self._wt.rank(np.array([self._byte], dtype=np.uint8), np.array([i], dtype=np.int64)).item())

    def __iter__(self):
        n1 = len(self)
        for s in range(0, n1, _CHUNK):
            yield from self[s:min(n1, s + _CHUNK)]

    def tolist(self):
        return list(self)

    def __eq__(self, other):
        if isinstance(other, (list, tuple, OccColumn)):
            return len(other) == len(self) and all(a == b for a, b in zip(self, other))
        return NotImplemented


class OccView(Mapping):
    """The dict build_occ returns, answered lazily by rank queries on the wavelet tree."""

    def __init__(self, wt):
        self._wt = wt
        self._keys = [chr(b) for b in wt.alphabet]

    def __getitem__(self, ch):
        if ch not in self._keys:
            raise KeyError(ch)
        return OccColumn(self._wt, ord(ch))

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def __contains__(self, ch):
        return ch in self._keys


def occ_mapping(wt):
    """dict[str, list[int]] when n * sigma is small, otherwise the lazy OccView."""
    view = OccView(wt)
    if (wt.n + 1) * max(1, wt.sigma) <= MATERIALIZE_MAX:
        return {k: view[k].tolist() for k in view}
    return view


class LazyList(Sequence):
    """A list computed on first use (Golomb code lists, next_text of big wavelet levels)."""

    def __init__(self, fetch):
        self._fetch = fetch
        self._data = None

    def _get(self):
        if self._data is None:
            self._data = self._fetch()
            self._fetch = None
        return self._data

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def tolist(self):
        return list(self._get())

    def __eq__(self, other):
        if isinstance(other, (list, tuple, LazyList)):
            return list(self._get()) == list(other)
        return NotImplemented

    def __repr__(self):
        return "LazyList(<pending>)" if self._data is None else f"LazyList(len={len(self._data)})"


def maybe_lazy(size_hint: int, fetch):
    """Materialise now when small, otherwise defer until first use."""
    return fetch() if size_hint <= MATERIALIZE_MAX else LazyList(fetch)

"""Typed data handles the generators accept (duck-typed like the reference's
lattice/preset.py:44-204: an object with `.load(key)` that returns something indexable,
plus `.Ne` for eigenvectors).

Only what the elemental path touches is here: in-memory handles for synthetic / already
loaded data, and numpy-memmap readers/writers for the raw-binary gauge field, the `.npy`
eigenvector file and the `.npy` elemental file ([Nop, Nmom, Lt, Ne, Ne] complex128,
tests/test_elemental.py:47 and lattice/data.py:26 of the reference), and readers of the two
big-endian production formats (ILDG gauge fields, QDP timeslice eigenvectors) that keep the file's
byte order so the device does the conversion.  The reference's own handle objects
(GaugeFieldIldg, EigenvectorTimeSlice, ...) work unchanged as inputs too.
"""
from __future__ import annotations

from time import perf_counter
from typing import List, Sequence

import numpy as np


class FileMetaData:
    def __init__(self, shape: Sequence[int], dtype: str = "<c16", extra=None):
        self.shape = list(shape)
        self.dtype = dtype
        self.extra = extra


class ArrayData:
    """Indexable view with the bookkeeping attributes of the reference's FileData
    (lattice/filedata/abstract.py:11-20): `.file`, `.shape`, `.dtype`, I/O counters."""

    def __init__(self, array, file: str = "<memory>"):
        self._a = array
        self.file = file
        self.shape = list(array.shape)
        self.dtype = array.dtype.str
        self.time_in_sec = 0.0
        self.size_in_byte = 0

    def __getitem__(self, key):
        s = perf_counter()
        ret = np.ascontiguousarray(self._a[key])
        self.time_in_sec += perf_counter() - s
        self.size_in_byte += ret.nbytes
        return ret


class GaugeField:
    def __init__(self, elem: FileMetaData) -> None:
        self.elem = elem


class Eigenvector:
    def __init__(self, elem: FileMetaData, eigenNum: int) -> None:
        self.elem = elem
        self.Ne = eigenNum


class Elemental:
    def __init__(self, elem: FileMetaData, eigenNum: int) -> None:
        self.elem = elem
        self.Ne = eigenNum


# ---------------------------------------------------------------------------------------------
# in-memory handles
# ---------------------------------------------------------------------------------------------
class GaugeFieldHostmem(GaugeField):
    """A whole configuration already in host memory: [Lt, Lz, Ly, Lx, Nd, Nc, Nc] complex128."""

    def __init__(self, host_ndarray: np.ndarray) -> None:
        if host_ndarray.ndim != 7 or host_ndarray.shape[-3:] != (4, 3, 3):
            raise ValueError(f"gauge field must be [Lt, Lz, Ly, Lx, 4, 3, 3], got {host_ndarray.shape}")
        super().__init__(FileMetaData(host_ndarray.shape, host_ndarray.dtype.str))
        self.data = ArrayData(host_ndarray)

    def load(self, key: str = None):
        return self.data


class EigenvectorHostmem(Eigenvector):
    """Eigenvectors already in host memory: [Lt, Ne, Lz, Ly, Lx, Nc], complex64 or complex128
    (cf. lattice/preset.py:77-89)."""

    def __init__(self, host_ndarray: np.ndarray, totNe: int = None) -> None:
        if host_ndarray.ndim != 6 or host_ndarray.shape[-1] != 3:
            raise ValueError(f"eigenvectors must be [Lt, Ne, Lz, Ly, Lx, 3], got {host_ndarray.shape}")
        Ne = host_ndarray.shape[1] if totNe is None else totNe
        if Ne > host_ndarray.shape[1]:
            raise ValueError("totNe exceeds the number of stored eigenvectors")
        super().__init__(FileMetaData(host_ndarray.shape, host_ndarray.dtype.str), Ne)
        self.data = ArrayData(host_ndarray)

    def load(self, key: str = None):
        return self.data


# ---------------------------------------------------------------------------------------------
# file handles (numpy memmap; one open per load, not one per eigenvector)
# ---------------------------------------------------------------------------------------------
class _KeyedFile:
    def __init__(self, prefix: str, suffix: str):
        self.prefix = prefix
        self.suffix = suffix
        self.file = None
        self.data = None

    def _name(self, key: str) -> str:
        return f"{self.prefix}{key}{self.suffix}"


class GaugeFieldBinary(_KeyedFile, GaugeField):
    """Raw binary [Lt, Lz, Ly, Lx, Nd, Nc, Nc] (lattice/preset.py:162-170)."""

    def __init__(self, prefix: str, suffix: str, shape: List[int], dtype: str = "<c16") -> None:
        _KeyedFile.__init__(self, prefix, ".dat" if suffix is None else suffix)
        GaugeField.__init__(self, FileMetaData(shape, dtype, 0))

    def load(self, key: str):
        name = self._name(key)
        if self.file != name:
            mm = np.memmap(name, dtype=self.elem.dtype, mode="r", shape=tuple(self.elem.shape))
            self.file, self.data = name, ArrayData(mm, name)
        return self.data


class GaugeFieldIldg(_KeyedFile, GaugeField):
    """ILDG / LIME configuration (lattice/preset.py:140-148).  `load(key)[:]` is the binary payload as
    stored, big-endian [Lt, Lz, Ly, Lx, Nd, Nc, Nc] (or `shape` with the same element count, e.g.
    the reference's flattened default): one memory map per file, no host-side byte swap."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = None) -> None:
        _KeyedFile.__init__(self, prefix, ".lime" if suffix is None else suffix)
        GaugeField.__init__(self, FileMetaData(shape or [], ">c16", 0))

    def load(self, key: str):
        from .fileio import ildg_layout, ildg_memmap

        name = self._name(key)
        if self.file != name:
            data = ArrayData(ildg_memmap(name, self.elem.shape or None), name)
            data.latt_size = ildg_layout(name)[2]
            self.file, self.data = name, data
        return self.data


class EigenvectorTimeSlice(_KeyedFile, Eigenvector):
    """QDP LazyDiskMapObj file with one big-endian complex64 record per (t, e)
    (lattice/preset.py:55-63).  `shape` = [Lt, Ne, ..., Nc]; `load(key)[t]` returns all Ne records of a
    timeslice from one memory map (the reference re-opens the file for every eigenvector)."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = [128, 70, 16**3, 3], totNe: int = 70) -> None:
        _KeyedFile.__init__(self, prefix, ".stout.n20.f0.12.laplace_eigs.3d.mod" if suffix is None else suffix)
        Eigenvector.__init__(self, FileMetaData(shape, ">c8", 2), totNe)

    def load(self, key: str):
        from .fileio import TimesliceRecords

        name = self._name(key)
        if self.file != name:
            self.file, self.data = name, TimesliceRecords(name, self.elem.shape, self.elem.dtype, self.elem.extra)
        return self.data


class GaugeFieldNpy(_KeyedFile, GaugeField):
    def __init__(self, prefix: str, suffix: str, shape: List[int] = None) -> None:
        _KeyedFile.__init__(self, prefix, ".npy" if suffix is None else suffix)
        GaugeField.__init__(self, FileMetaData(shape or [], "<c16", 0))

    def load(self, key: str):
        name = self._name(key)
        if self.file != name:
            self.file, self.data = name, ArrayData(np.load(name, mmap_mode="r"), name)
        return self.data


class EigenvectorNpy(_KeyedFile, Eigenvector):
    """`.npy` [Lt, Ne, Lz, Ly, Lx, Nc] (lattice/preset.py:66-74; shape/dtype come from the header)."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = None, totNe: int = 70) -> None:
        _KeyedFile.__init__(self, prefix, ".lime.npy" if suffix is None else suffix)
        Eigenvector.__init__(self, FileMetaData(shape or [], "<c16", 2), totNe)

    def load(self, key: str):
        name = self._name(key)
        if self.file != name:
            self.file, self.data = name, ArrayData(np.load(name, mmap_mode="r"), name)
        return self.data


class ElementalNpy(_KeyedFile, Elemental):
    """`.npy` [Nop, Nmom, Lt, Ne, Ne] (lattice/preset.py:173-181)."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = None, totNe: int = 70) -> None:
        _KeyedFile.__init__(self, prefix, ".stout.n20.f0.12.nev70.meson.npy" if suffix is None else suffix)
        Elemental.__init__(self, FileMetaData(shape or [], "<c16", 0), totNe)

    def load(self, key: str):
        name = self._name(key)
        if self.file != name:
            self.file, self.data = name, ArrayData(np.load(name, mmap_mode="r"), name)
        return self.data

    def create(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        """Pre-sized writable file so every timeslice (or rank) can drop its slab in place."""
        name = self._name(key)
        self.file = self.data = None
        return np.lib.format.open_memmap(name, mode="w+", dtype=dtype, shape=tuple(int(s) for s in shape))

    def open_rw(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        """Re-open a file made by `create` for in-place writes (another rank's timeslice slab)."""
        mm = np.lib.format.open_memmap(self._name(key), mode="r+")
        if tuple(mm.shape) != tuple(int(s) for s in shape) or mm.dtype != np.dtype(dtype):
            raise ValueError(f"{self._name(key)} has shape {mm.shape} / dtype {mm.dtype}, expected {tuple(shape)} / {dtype}")
        self.file = self.data = None
        return mm


class ElementalBinary(_KeyedFile, Elemental):
    """Raw little-endian binary [Nop, Nmom, Lt, Ne, Ne] without a header (lattice/preset.py:129-137,
    lattice/filedata/binary.py:16-62): shape and dtype come from the constructor, as in the reference."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = [40, 27, 128, 70, 70], totNe: int = 70, dtype: str = "<c16") -> None:
        _KeyedFile.__init__(self, prefix, ".stout.n20.f0.12.nev70.meson" if suffix is None else suffix)
        Elemental.__init__(self, FileMetaData(shape, dtype, 0), totNe)

    def load(self, key: str):
        name = self._name(key)
        if self.file != name:
            mm = np.memmap(name, dtype=self.elem.dtype, mode="r", shape=tuple(self.elem.shape))
            self.file, self.data = name, ArrayData(mm, name)
        return self.data

    def _check(self, shape, dtype):
        if [int(v) for v in shape] != [int(v) for v in self.elem.shape] or np.dtype(dtype) != np.dtype(self.elem.dtype):
            raise ValueError(f"ElementalBinary declared {self.elem.shape} / {self.elem.dtype}, asked to hold {list(shape)} / {dtype}: "
                             "a headerless file can only be read back with the shape and dtype it was written with")

    def create(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        self._check(shape, dtype)
        self.file = self.data = None
        return np.memmap(self._name(key), dtype=dtype, mode="w+", shape=tuple(int(v) for v in shape))

    def open_rw(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        import os

        self._check(shape, dtype)
        name = self._name(key)
        want = int(np.prod([int(v) for v in shape])) * np.dtype(dtype).itemsize
        if os.path.getsize(name) != want:
            raise ValueError(f"{name} holds {os.path.getsize(name)} bytes, expected {want}")
        self.file = self.data = None
        return np.memmap(name, dtype=dtype, mode="r+", shape=tuple(int(v) for v in shape))

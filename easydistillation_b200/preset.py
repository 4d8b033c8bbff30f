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
    """Eigenvectors already in host memory, indexed [t, e] like every eigenvector handle: [Lt, Ne, Lz, Ly, Lx, Nc] or
    the flattened [Lt, Ne, Lz*Ly*Lx, Nc], complex64 or complex128.

    Signature of the reference, `EigenvectorHostmem(host_ndarray, shape, totNe)` (lattice/preset.py:77-89): `shape`, when
    given, must equal the array's shape (ValueError otherwise, as there; a list or a tuple is accepted - the reference
    only ever matches a tuple), `totNe` defaults to the number of stored eigenvectors.  The array is referenced, not
    copied.  An int in the second position is taken as totNe (this package's earlier two-argument form)."""

    def __init__(self, host_ndarray: np.ndarray, shape: Sequence[int] = None, totNe: int = None) -> None:
        if isinstance(shape, (int, np.integer)) and totNe is None:
            shape, totNe = None, int(shape)
        if host_ndarray.ndim not in (4, 6) or host_ndarray.shape[-1] != 3:
            raise ValueError(f"eigenvectors must be [Lt, Ne, Lz, Ly, Lx, 3] or [Lt, Ne, Lz*Ly*Lx, 3], got {host_ndarray.shape}")
        if shape is not None and tuple(host_ndarray.shape) != tuple(int(v) for v in shape):
            raise ValueError(f"Please check that host_ndarray shape {host_ndarray.shape} does not match expected shape "
                             f"(Lt, totNe, Lz, Ly, Lx, Nc) = {list(shape)}")
        Ne = host_ndarray.shape[1] if totNe is None else int(totNe)
        if Ne > host_ndarray.shape[1]:
            raise ValueError("totNe exceeds the number of stored eigenvectors")
        super().__init__(FileMetaData(host_ndarray.shape, host_ndarray.dtype.str), Ne)
        self.data = ArrayData(host_ndarray)

    def load(self, key: str = None):
        return self.data


# ---------------------------------------------------------------------------------------------
# device-resident handles: inputs that already sit in HBM ("additionally accept torch CUDA tensors")
# ---------------------------------------------------------------------------------------------
class DeviceTimeslices:
    """Per-timeslice CUDA tensors behind the indexing the generators use: `data[t]` is the tensor of timeslice t,
    `data[t, e]` one eigenvector.  `cyclic=True` repeats the given timeslices (benchmarks: two resident input sets
    alternated over many timeslices)."""

    device_resident = True

    def __init__(self, tensors, cyclic: bool = False, file: str = "<device>"):
        self._t = list(tensors) if isinstance(tensors, (list, tuple)) else [tensors[i] for i in range(tensors.shape[0])]
        if not self._t:
            raise ValueError("no timeslices")
        self.cyclic = bool(cyclic)
        self.file = file

    def __len__(self):
        return len(self._t)

    def __getitem__(self, key):
        if isinstance(key, slice):  # `data[:]` of the gauge-field protocol: the indexable itself
            if key != slice(None):
                raise IndexError("only [:] and integer timeslices are supported on device-resident data")
            return self
        t, rest = (key[0], key[1:]) if isinstance(key, tuple) else (key, ())
        n = len(self._t)
        if self.cyclic:
            t = t % n
        elif not 0 <= t < n:
            raise IndexError(f"timeslice {t} out of range [0, {n})")
        x = self._t[t]
        return x[rest] if rest else x


class GaugeFieldDevice(GaugeField):
    """Gauge links in HBM: a CUDA tensor [Lt, Lz, Ly, Lx, Nd, Nc, Nc] complex128 or a list of per-timeslice tensors."""

    def __init__(self, tensors, cyclic: bool = False) -> None:
        self.data = DeviceTimeslices(tensors, cyclic)
        first = self.data[0]
        super().__init__(FileMetaData([len(self.data)] + list(first.shape), "<c16"))

    def load(self, key: str = None):
        return self.data


class EigenvectorDevice(Eigenvector):
    """Eigenvectors in HBM: a CUDA tensor [Lt, Ne, Lz, Ly, Lx, Nc] (complex64 or complex128) or per-timeslice tensors."""

    def __init__(self, tensors, totNe: int = None, cyclic: bool = False) -> None:
        self.data = DeviceTimeslices(tensors, cyclic)
        first = self.data[0]
        Ne = int(first.shape[0]) if totNe is None else int(totNe)
        if Ne > int(first.shape[0]):
            raise ValueError("totNe exceeds the number of stored eigenvectors")
        super().__init__(FileMetaData([len(self.data)] + list(first.shape), "<c8" if first.element_size() == 8 else "<c16"), Ne)

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
    """Raw binary [Lt, Lz, Ly, Lx, Nd, Nc, Nc] (lattice/preset.py:162-170, same defaults: the flattened
    [128, 16^3, 4, 3, 3] shape and '<f8' - a gauge field needs dtype='<c16', which is what the reference's own
    scripts pass; the generators refuse a real dtype)."""

    def __init__(self, prefix: str, suffix: str, shape: List[int] = [128, 16**3, 4, 3, 3], dtype: str = "<f8") -> None:
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


def _ensure_file(name, reopen, make, shape, dtype, replace) -> bool:
    """See ElementalNpy.ensure.  `reopen()` maps the existing file (raising if it cannot hold shape / dtype),
    `make(tmp)` writes a fully sized new file under a temporary name."""
    import os

    if os.path.exists(name):
        ok = False
        try:
            mm = reopen()
            ok = tuple(mm.shape) == shape and mm.dtype == dtype
            del mm
        except Exception:
            ok = False
        if ok:
            return False
        if not replace:
            raise ValueError(f"{name} exists with another shape / dtype than {shape} / {dtype}; remove it, or let the "
                             "coordinating process create the file (calc_to_file under torch.distributed, or without t_range)")
    tmp = f"{name}.tmp.{os.getpid()}"
    mm = make(tmp)
    mm.flush()
    del mm
    try:
        if replace:
            os.replace(tmp, name)
            return True
        try:
            os.link(tmp, name)  # fails if another process created the file in the meantime: theirs is used
            return True
        except FileExistsError:
            return False
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


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

    def ensure(self, key: str, shape: Sequence[int], dtype: str = "<c16", replace: bool = False) -> bool:
        """Make sure a pre-sized file of this shape / dtype exists WITHOUT ever truncating one that already does.
        A missing file is written under a temporary name and hard-linked into place (atomic: if another process won
        the race, its file is the one used).  A file of another shape is replaced only with `replace=True` (the single
        coordinating process of a run), else ValueError.  Returns True if this call created the file."""
        return _ensure_file(self._name(key), lambda: np.lib.format.open_memmap(self._name(key), mode="r+"),
                            lambda tmp: np.lib.format.open_memmap(tmp, mode="w+", dtype=dtype, shape=tuple(int(s) for s in shape)),
                            tuple(int(s) for s in shape), np.dtype(dtype), replace)

    def open_rw(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        """Re-open a file made by `create` / `ensure` for in-place writes (a rank's timeslice slab)."""
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

    def ensure(self, key: str, shape: Sequence[int], dtype: str = "<c16", replace: bool = False) -> bool:
        """As ElementalNpy.ensure; a headerless file is recognised by its size."""
        self._check(shape, dtype)
        name = self._name(key)
        shp = tuple(int(v) for v in shape)

        def reopen():
            import os

            want = int(np.prod(shp)) * np.dtype(dtype).itemsize
            if os.path.getsize(name) != want:
                raise ValueError(f"{name} holds {os.path.getsize(name)} bytes, expected {want}")
            return np.memmap(name, dtype=dtype, mode="r+", shape=shp)

        return _ensure_file(name, reopen, lambda tmp: np.memmap(tmp, dtype=dtype, mode="w+", shape=shp), shp, np.dtype(dtype), replace)

    def open_rw(self, key: str, shape: Sequence[int], dtype: str = "<c16") -> np.memmap:
        import os

        self._check(shape, dtype)
        name = self._name(key)
        want = int(np.prod([int(v) for v in shape])) * np.dtype(dtype).itemsize
        if os.path.getsize(name) != want:
            raise ValueError(f"{name} holds {os.path.getsize(name)} bytes, expected {want}")
        self.file = self.data = None
        return np.memmap(name, dtype=dtype, mode="r+", shape=tuple(int(v) for v in shape))

"""Containers of the two big-endian input formats of the elemental path, read ONCE per file and
handed to the device as raw payload (the kernels that consume them swap the bytes):

  * ILDG / LIME gauge configurations   (reference reader: lattice/filedata/ildg.py:49-105)
  * QDP "LazyDiskMapObj" timeslice files of eigenvectors, one record per (t, e)
                                        (reference reader: lattice/filedata/timeslice.py:13-122)

The reference opens + mmaps the file again for every `__getitem__` (once per eigenvector and
timeslice) and converts to little-endian on the host; here one memory map per file serves every
access and no host-side conversion takes place.  The writers exist for tests and synthetic data.
"""
from __future__ import annotations

import re
import struct
import xml.etree.ElementTree as ET
from time import perf_counter
from typing import Dict, List, Sequence, Tuple

import numpy as np

LIME_MAGIC = 0x456789AB
LIME_HEADER_BYTES = 144  # magic u32, version u16, flags u16, payload length u64, type char[128]
QDP_MAGIC = "XXXXQDPLazyDiskMapObjFileXXXX"


# -------------------------------------------------------------------------------------------------
# LIME / ILDG
# -------------------------------------------------------------------------------------------------
def lime_records(path: str) -> List[Tuple[str, int, int]]:
    """[(record type, payload offset, payload length)] of a LIME file."""
    out = []
    with open(path, "rb") as f:
        pos = 0
        while True:
            f.seek(pos)
            head = f.read(LIME_HEADER_BYTES)
            if len(head) < LIME_HEADER_BYTES:  # end of file (some writers leave a trailing newline)
                break
            magic, _version, _flags, length = struct.unpack(">IHHQ", head[:16])
            if magic != LIME_MAGIC:
                raise ValueError(f"{path}: no LIME record header at byte {pos}")
            rtype = head[16:].split(b"\0", 1)[0].decode("utf-8")
            out.append((rtype, pos + LIME_HEADER_BYTES, length))
            pos += LIME_HEADER_BYTES + (length + 7) // 8 * 8
    return out


def _xml_fields(text: str) -> Dict[str, str]:
    root = ET.fromstring(text)
    return {re.sub(r"^\{.*\}", "", child.tag): (child.text or "").strip() for child in root}


def ildg_layout(path: str):
    """(payload offset, payload length, [lx, ly, lz, lt], precision in bits) of an ILDG file."""
    recs = {rtype: (off, length) for rtype, off, length in lime_records(path)}
    if "ildg-format" not in recs or "ildg-binary-data" not in recs:
        raise ValueError(f"{path}: not an ILDG file (records: {sorted(recs)})")
    off, length = recs["ildg-format"]
    with open(path, "rb") as f:
        f.seek(off)
        fields = _xml_fields(f.read(length).strip(b"\0").decode("utf-8"))
    latt = [int(fields[k]) for k in ("lx", "ly", "lz", "lt")]
    return recs["ildg-binary-data"][0], recs["ildg-binary-data"][1], latt, int(fields["precision"])


def ildg_memmap(path: str, shape: Sequence[int] = None) -> np.memmap:
    """Read-only big-endian view of the binary payload, [Lt, Lz, Ly, Lx, 4, 3, 3] unless `shape`
    (same number of elements) says otherwise.  dtype `>c16` (precision 64) or `>c8` (32)."""
    off, length, (lx, ly, lz, lt), precision = ildg_layout(path)
    if precision not in (32, 64):
        raise ValueError(f"{path}: unsupported ILDG precision {precision}")
    dtype = np.dtype(">c16" if precision == 64 else ">c8")
    full = (lt, lz, ly, lx, 4, 3, 3)
    shape = full if shape is None else tuple(int(s) for s in shape)
    if int(np.prod(shape)) != int(np.prod(full)) or int(np.prod(full)) * dtype.itemsize != length:
        raise ValueError(f"{path}: payload of {length} bytes / lattice {(lx, ly, lz, lt)} does not match shape {shape}")
    return np.memmap(path, dtype=dtype, mode="r", offset=off, shape=shape)


def _lime_record(rtype: str, payload: bytes, first: bool, last: bool) -> bytes:
    flags = (0x8000 if first else 0) | (0x4000 if last else 0)  # message begin / message end bits
    head = struct.pack(">IHHQ", LIME_MAGIC, 1, flags, len(payload)) + rtype.encode("utf-8").ljust(128, b"\0")
    return head + payload + b"\0" * (-len(payload) % 8)


def write_ildg(path: str, U: np.ndarray, precision: int = 64) -> None:
    """U [Lt, Lz, Ly, Lx, 4, 3, 3] -> ILDG file (ildg-format, ildg-binary-data, ildg-data-lfn records)."""
    if U.ndim != 7 or U.shape[4:] != (4, 3, 3):
        raise ValueError(f"gauge field must be [Lt, Lz, Ly, Lx, 4, 3, 3], got {U.shape}")
    lt, lz, ly, lx = U.shape[:4]
    xml = ('<?xml version="1.0" encoding="UTF-8"?><ildgFormat xmlns="http://www.lqcd.org/ildg" '
           'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
           'xsi:schemaLocation="http://www.lqcd.org/ildg http://www.lqcd.org/ildg/filefmt.xsd">'
           f"<version>1.0</version><field>su3gauge</field><precision>{precision}</precision>"
           f"<lx>{lx}</lx><ly>{ly}</ly><lz>{lz}</lz><lt>{lt}</lt></ildgFormat>").encode("utf-8")
    data = np.ascontiguousarray(U, dtype=">c16" if precision == 64 else ">c8").tobytes()
    with open(path, "wb") as f:
        f.write(_lime_record("ildg-format", xml, True, False))
        f.write(_lime_record("ildg-binary-data", data, False, False))
        f.write(_lime_record("ildg-data-lfn", b"lfn://easydistillation_b200/synthetic", False, True))


# -------------------------------------------------------------------------------------------------
# QDP LazyDiskMapObj timeslice files
# -------------------------------------------------------------------------------------------------
def _qdp_str(buf, pos: int) -> Tuple[str, int]:
    (n,) = struct.unpack_from(">i", buf, pos)
    return bytes(buf[pos + 4 : pos + 4 + n]).decode("utf-8"), pos + 4 + n


def qdp_index(path: str):
    """({key tuple: payload offset}, metadata fields, version) of a LazyDiskMapObj file: magic string,
    version, XML metadata, position of the key table; the table lists (key, position) pairs."""
    with open(path, "rb") as f:
        head = f.read(1 << 16)
        magic, pos = _qdp_str(head, 0)
        if magic != QDP_MAGIC:
            raise ValueError(f"{path}: not a QDP LazyDiskMapObj file")
        (version,) = struct.unpack_from(">i", head, pos)
        pos += 4
        (nxml,) = struct.unpack_from(">i", head, pos)
        if pos + 4 + nxml + 16 > len(head):
            f.seek(0)
            head = f.read(pos + 4 + nxml + 16)
        xml, pos = _qdp_str(head, pos)
        _, table_pos = struct.unpack_from(">qq", head, pos)
        f.seek(table_pos)
        table = f.read()
    meta = _xml_fields(xml)
    (nrec,) = struct.unpack_from(">I", table, 0)
    pos = 4
    offsets: Dict[Tuple[int, ...], int] = {}
    for _ in range(nrec):
        (nbytes,) = struct.unpack_from(">i", table, pos)
        key = struct.unpack_from(">" + "i" * (nbytes // 4), table, pos + 4)
        pos += 4 + nbytes
        _, where = struct.unpack_from(">qq", table, pos)
        pos += 16
        offsets[tuple(key)] = where
    return offsets, meta, version


class TimesliceRecords:
    """Indexable view of a (t, e)-keyed record file with the reference's FileData conventions
    (timeslice.py:65-99): `data[t, e]` is one record, `data[t]` all records of a timeslice, missing
    keys raise IndexError.  Records keep the file's byte order (`.dtype`, normally `>c8`)."""

    def __init__(self, path: str, shape: Sequence[int], dtype: str, nkey: int = 2):
        self.file = path
        self.offsets, self.meta, self.version = qdp_index(path)
        if int(self.meta.get("decay_dir", 3)) != 3:
            raise ValueError(f"{path}: decay_dir must be 3")
        self.latt_size = [int(v) for v in self.meta.get("lattSize", "").split()]
        self.extra = nkey
        self.extra_shape = list(shape[:nkey])
        self.shape = list(shape[nkey:])
        self.dtype = np.dtype(dtype)
        self._count = int(np.prod(self.shape))
        self._bytes = np.memmap(path, dtype=np.uint8, mode="r")
        self.time_in_sec = 0.0
        self.size_in_byte = 0

    def _record(self, key: Tuple[int, ...]) -> np.ndarray:
        if key not in self.offsets:
            raise IndexError(f"index {key} is out of bounds for axes")
        off = self.offsets[key]
        return self._bytes[off : off + self._count * self.dtype.itemsize].view(self.dtype).reshape(self.shape)

    def __getitem__(self, key):
        s = perf_counter()
        if isinstance(key, (int, np.integer)):
            key = (int(key),)
        key = tuple(int(k) for k in key)
        if len(key) >= self.extra:
            ret = np.array(self._record(key[: self.extra])[key[self.extra :]])
        else:  # all records below a key prefix, e.g. data[t] -> [Ne, ...]
            n = self.extra_shape[len(key)]
            if len(key) != self.extra - 1:
                raise IndexError("only one key axis may be left open")
            rec = self._count * self.dtype.itemsize
            offs = [self.offsets.get(key + (e,)) for e in range(n)]
            if None in offs:
                raise IndexError(f"index {key} is out of bounds for axes")
            if all(offs[e] == offs[0] + e * rec for e in range(n)):  # back-to-back records: no gather needed
                ret = self._bytes[offs[0] : offs[0] + n * rec].view(self.dtype).reshape([n] + self.shape)
            else:
                ret = np.empty([n] + self.shape, self.dtype)
                for e in range(n):
                    ret[e] = self._record(key + (e,))
        self.time_in_sec += perf_counter() - s
        self.size_in_byte += ret.nbytes
        return ret


def write_qdp_timeslices(path: str, V: np.ndarray, latt_size: Sequence[int], dtype: str = ">c8") -> None:
    """V [Lt, Ne, ...] -> LazyDiskMapObj file with one record per (t, e), keys in (t, e) order."""
    Lt, Ne = V.shape[:2]
    xml = ("<MODMetaData><id>eigenVecsTimeSlice</id>"
           f"<lattSize>{' '.join(str(int(v)) for v in latt_size)}</lattSize><decay_dir>3</decay_dir>"
           f"<num_vecs>{Ne}</num_vecs></MODMetaData>").encode("utf-8")
    magic = QDP_MAGIC.encode("utf-8")
    head = struct.pack(">i", len(magic)) + magic + struct.pack(">i", 1) + struct.pack(">i", len(xml)) + xml
    data_pos = len(head) + 16
    rec = int(np.prod(V.shape[2:])) * np.dtype(dtype).itemsize
    table_pos = data_pos + Lt * Ne * rec
    with open(path, "wb") as f:
        f.write(head + struct.pack(">qq", 0, table_pos))
        f.write(np.ascontiguousarray(V, dtype=dtype).tobytes())
        f.write(struct.pack(">I", Lt * Ne))
        for t in range(Lt):
            for e in range(Ne):
                f.write(struct.pack(">iii", 8, t, e) + struct.pack(">qq", 0, data_pos + (t * Ne + e) * rec))

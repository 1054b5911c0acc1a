"""GGUF v2/v3 container: reader (mmap) and writer.

Wire format the hot path consumes.  Mirrors the reference parser
``tensor/GGUF.java:43-92`` (header, key/values, tensor infos, alignment padding,
tensor-data offset) and ``tensor/GGUF.java:105-137`` (one mapping of the tensor-data
section, per-tensor slices), plus ``tensor/GGMLType.java:5-20`` (block sizes:
Q8_0 = 34 bytes / 32 elements).  The writer has no reference counterpart: it exists
because no real checkpoints are available offline and every BASELINE config runs on
seeded synthetic GGUF files.
"""
from __future__ import annotations

import mmap
import struct
from dataclasses import dataclass
from typing import Any, BinaryIO

import numpy as np

GGUF_MAGIC = 0x46554747  # "GGUF" little-endian
DEFAULT_ALIGNMENT = 32


class GGMLType:
    F32 = 0
    F16 = 1
    Q8_0 = 8
    Q4_K = 12  # K-quants: accepted for Q8_0 plans, re-quantised on the device at upload (ModelLoader.java:163, csrc/kquant.cuh)
    Q5_K = 13
    Q6_K = 14
    NAMES = {0: "F32", 1: "F16", 8: "Q8_0", 12: "Q4_K", 13: "Q5_K", 14: "Q6_K"}
    # (type_size_bytes, block_size_elems), GGMLType.java:5-20
    SIZES = {0: (4, 1), 1: (2, 1), 8: (34, 32), 12: (144, 256), 13: (176, 256), 14: (210, 256)}
    K_QUANTS = (12, 13, 14)

    @staticmethod
    def byte_size_for(ggml_type: int, n_elements: int) -> int:
        ts, bs = GGMLType.SIZES[ggml_type]
        assert n_elements % bs == 0
        return n_elements // bs * ts


# metadata value types, MetadataValueType.java
_U8, _I8, _U16, _I16, _U32, _I32, _F32, _BOOL, _STR, _ARR, _U64, _I64, _F64 = range(13)
_SCALAR_FMT = {_U8: "<B", _I8: "<b", _U16: "<H", _I16: "<h", _U32: "<I", _I32: "<i", _F32: "<f",
               _BOOL: "<?", _U64: "<Q", _I64: "<q", _F64: "<d"}


@dataclass
class TensorInfo:
    name: str
    dims: tuple[int, ...]  # ne[0] is the innermost (column) dimension
    ggml_type: int
    offset: int  # relative to the tensor-data section

    @property
    def n_elements(self) -> int:
        n = 1
        for d in self.dims:
            n *= d
        return n

    @property
    def n_bytes(self) -> int:
        return GGMLType.byte_size_for(self.ggml_type, self.n_elements)


class _Reader:
    def __init__(self, buf: memoryview):
        self.buf = buf
        self.pos = 0

    def scalar(self, vt: int):
        fmt = _SCALAR_FMT[vt]
        (v,) = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += struct.calcsize(fmt)
        return v

    def string(self) -> str:
        n = self.scalar(_U64)
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8", errors="replace")
        self.pos += n
        return s

    def value(self, vt: int):
        if vt == _STR:
            return self.string()
        if vt == _ARR:
            et = self.scalar(_U32)
            n = self.scalar(_U64)
            return [self.value(et) for _ in range(n)]
        return self.scalar(vt)


class GGUFFile:
    """Parsed GGUF file with the tensor-data section memory-mapped read-only."""

    def __init__(self, path: str):
        self.path = str(path)
        self._f = open(self.path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        r = _Reader(memoryview(self._mm))
        magic = r.scalar(_U32)
        if magic != GGUF_MAGIC:
            raise ValueError(f"unsupported magic.number {magic:#x}")
        self.version = r.scalar(_U32)
        if self.version not in (2, 3):
            raise ValueError(f"unsupported version {self.version}")
        tensor_count = r.scalar(_U64)
        kv_count = r.scalar(_U64)
        self.metadata: dict[str, Any] = {}
        for _ in range(kv_count):
            key = r.string()
            vt = r.scalar(_U32)
            self.metadata[key] = r.value(vt)
        self.tensor_infos: dict[str, TensorInfo] = {}
        for _ in range(tensor_count):
            name = r.string()
            nd = r.scalar(_U32)
            dims = tuple(r.scalar(_U64) for _ in range(nd))
            tt = r.scalar(_U32)
            off = r.scalar(_U64)
            self.tensor_infos[name] = TensorInfo(name, dims, tt, off)
        self.alignment = int(self.metadata.get("general.alignment", DEFAULT_ALIGNMENT))
        pad = (self.alignment - (r.pos % self.alignment)) % self.alignment
        self.tensor_data_offset = r.pos + pad

    def tensor_bytes(self, name: str) -> np.ndarray:
        """uint8 view (zero-copy) of one tensor's raw GGUF bytes."""
        ti = self.tensor_infos[name]
        start = self.tensor_data_offset + ti.offset
        return np.frombuffer(self._mm, dtype=np.uint8, count=ti.n_bytes, offset=start)

    def close(self):
        try:
            self._mm.close()
        except BufferError:
            pass  # numpy views still alive; the mapping goes with them
        self._f.close()


def _write_string(f: BinaryIO, s: str):
    b = s.encode("utf-8")
    f.write(struct.pack("<Q", len(b)))
    f.write(b)


def _write_value(f: BinaryIO, v: Any):
    """Python value -> (type, payload).  ints -> UINT32 (the reference casts to int),
    floats -> FLOAT32, list[str] -> ARRAY of STRING, list[int] -> ARRAY of INT32."""
    if isinstance(v, bool):
        f.write(struct.pack("<I?", _BOOL, v))
    elif isinstance(v, int):
        f.write(struct.pack("<II", _U32, v))
    elif isinstance(v, float):
        f.write(struct.pack("<If", _F32, v))
    elif isinstance(v, str):
        f.write(struct.pack("<I", _STR))
        _write_string(f, v)
    elif isinstance(v, (list, tuple)):
        f.write(struct.pack("<I", _ARR))
        if len(v) and isinstance(v[0], str):
            f.write(struct.pack("<IQ", _STR, len(v)))
            for s in v:
                _write_string(f, s)
        elif len(v) and isinstance(v[0], float):
            f.write(struct.pack("<IQ", _F32, len(v)))
            f.write(np.asarray(v, dtype="<f4").tobytes())
        else:
            f.write(struct.pack("<IQ", _I32, len(v)))
            f.write(np.asarray(v, dtype="<i4").tobytes())
    else:
        raise TypeError(f"unsupported metadata value {type(v)}")


def write_gguf(path: str, metadata: dict[str, Any], tensors: list[tuple[str, int, tuple[int, ...], np.ndarray]],
               alignment: int = DEFAULT_ALIGNMENT):
    """Write a GGUF v3 file.  ``tensors`` = [(name, ggml_type, dims(ne0 innermost), raw uint8 bytes)]."""
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQQ", GGUF_MAGIC, 3, len(tensors), len(metadata)))
        for k, v in metadata.items():
            _write_string(f, k)
            _write_value(f, v)
        off = 0
        offsets = []
        for name, tt, dims, raw in tensors:
            nbytes = GGMLType.byte_size_for(tt, int(np.prod(dims)))
            assert raw.dtype == np.uint8 and raw.size == nbytes, (name, raw.size, nbytes)
            _write_string(f, name)
            f.write(struct.pack("<I", len(dims)))
            for d in dims:
                f.write(struct.pack("<Q", d))
            f.write(struct.pack("<IQ", tt, off))
            offsets.append(off)
            off += (nbytes + alignment - 1) // alignment * alignment
        pad = (alignment - (f.tell() % alignment)) % alignment
        f.write(b"\0" * pad)
        base = f.tell()
        for (name, tt, dims, raw), o in zip(tensors, offsets):
            f.seek(base + o)
            f.write(raw.tobytes() if not raw.flags.c_contiguous else memoryview(raw))
        end = base + off
        f.seek(0, 2)
        if f.tell() < end:
            f.write(b"\0" * (end - f.tell()))

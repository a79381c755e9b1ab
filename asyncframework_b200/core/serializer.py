"""Serializers and block compression codecs.

Reference: closures are always Java-serialised (core/SparkEnv.scala:289), data goes through ``spark.serializer``
(Java default / Kryo, core/serializer/KryoSerializer.scala:55); broadcast pieces, shuffle blocks and task results
are compressed with ``spark.io.compression.codec`` = lz4 (default) / lzf / snappy / zstd
(core/io/CompressionCodec.scala:58,99,113,156,235 — JNI libraries lz4-java / snappy-java / zstd-jni).

Here: ``pickle`` (data; protocol 5, out-of-band-free) and ``cloudpickle`` (closures: lambdas, local functions,
captured state — what ``ClosureCleaner`` + Java serialisation do for Scala closures).  Codecs: ``lz4`` and ``zstd``
bind the system ``liblz4.so.1`` / ``libzstd.so.1`` through ctypes (block API; the native libraries the reference
reaches through JNI), ``zlib`` / ``lzma`` / ``bz2`` come from the standard library, ``none`` is the identity.
Every compressed block is self-describing: 1 codec byte + 8-byte little-endian raw length + CRC32 of the raw bytes, so a reader never
needs out-of-band knowledge of which codec the writer was configured with.
"""
from __future__ import annotations

import bz2
import ctypes
import ctypes.util
import lzma
import pickle
import struct
import zlib
from typing import Any, Callable, Dict, Optional

import cloudpickle


# ------------------------------------------------------------------------------------------------ serializers
class Serializer:
    name = "abstract"

    def dumps(self, obj: Any) -> bytes:
        raise NotImplementedError

    def loads(self, data) -> Any:
        return pickle.loads(data)

    @staticmethod
    def create(name: str) -> "Serializer":
        try:
            return _SERIALIZERS[name.lower()]()
        except KeyError:
            raise ValueError(f"unknown serializer {name!r} (have: {sorted(_SERIALIZERS)})") from None


class PickleSerializer(Serializer):
    """Data serializer (``JavaSerializer``'s role).  Falls back to cloudpickle for objects plain pickle refuses
    (records that hold lambdas / locally defined classes)."""
    name = "pickle"

    def dumps(self, obj: Any) -> bytes:
        try:
            return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        except (pickle.PicklingError, AttributeError, TypeError):
            return cloudpickle.dumps(obj)


class CloudPickleSerializer(Serializer):
    """Closure serializer: by-value functions and classes."""
    name = "cloudpickle"

    def dumps(self, obj: Any) -> bytes:
        return cloudpickle.dumps(obj)


_SERIALIZERS: Dict[str, Callable[[], Serializer]] = {"pickle": PickleSerializer, "java": PickleSerializer,
                                                    "cloudpickle": CloudPickleSerializer, "kryo": PickleSerializer}


# ------------------------------------------------------------------------------------------------ codecs
_CODEC_IDS = {"none": 0, "zlib": 1, "lz4": 2, "zstd": 3, "lzma": 4, "bz2": 5}
_ID_CODECS = {v: k for k, v in _CODEC_IDS.items()}
_HDR = struct.Struct("<BQI")      # codec id, raw length, CRC32 of the raw bytes (LZ4BlockOutputStream carries a checksum too)


def _load(*names: str) -> Optional[ctypes.CDLL]:
    for n in names:
        try:
            return ctypes.CDLL(n)
        except OSError:
            continue
    return None


class _Lz4:
    lib = None

    @classmethod
    def get(cls):
        if cls.lib is None:
            lib = _load("liblz4.so.1", ctypes.util.find_library("lz4") or "liblz4.so")
            if lib is None:
                raise RuntimeError("liblz4 not found")
            lib.LZ4_compressBound.restype = ctypes.c_int
            lib.LZ4_compressBound.argtypes = [ctypes.c_int]
            lib.LZ4_compress_default.restype = ctypes.c_int
            lib.LZ4_compress_default.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
            lib.LZ4_decompress_safe.restype = ctypes.c_int
            lib.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
            cls.lib = lib
        return cls.lib

    CHUNK = 1 << 30      # LZ4's block API takes int sizes: larger inputs are framed as <u32 clen><chunk> records

    @classmethod
    def compress(cls, data: bytes) -> bytes:
        lib = cls.get()
        out = []
        for off in range(0, max(len(data), 1), cls.CHUNK):
            chunk = data[off:off + cls.CHUNK]
            bound = lib.LZ4_compressBound(len(chunk))
            buf = ctypes.create_string_buffer(bound)
            n = lib.LZ4_compress_default(chunk, buf, len(chunk), bound)
            if n <= 0 and len(chunk):
                raise RuntimeError("LZ4_compress_default failed")
            out.append(struct.pack("<II", n, len(chunk)))
            out.append(buf.raw[:n])
        return b"".join(out)

    @classmethod
    def decompress(cls, data: bytes, raw_len: int) -> bytes:
        lib = cls.get()
        out, off = [], 0
        while off < len(data):
            n, rlen = struct.unpack_from("<II", data, off)
            off += 8
            buf = ctypes.create_string_buffer(max(rlen, 1))
            got = lib.LZ4_decompress_safe(data[off:off + n], buf, n, rlen)
            if got != rlen:
                raise ValueError(f"corrupt lz4 block: expected {rlen} bytes, got {got}")
            out.append(buf.raw[:rlen])
            off += n
        res = b"".join(out)
        if len(res) != raw_len:
            raise ValueError("corrupt lz4 stream: length mismatch")
        return res


class _Zstd:
    lib = None

    @classmethod
    def get(cls):
        if cls.lib is None:
            lib = _load("libzstd.so.1", ctypes.util.find_library("zstd") or "libzstd.so")
            if lib is None:
                raise RuntimeError("libzstd not found")
            lib.ZSTD_compressBound.restype = ctypes.c_size_t
            lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
            lib.ZSTD_compress.restype = ctypes.c_size_t
            lib.ZSTD_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
            lib.ZSTD_decompress.restype = ctypes.c_size_t
            lib.ZSTD_decompress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
            lib.ZSTD_isError.restype = ctypes.c_uint
            lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
            cls.lib = lib
        return cls.lib

    @classmethod
    def compress(cls, data: bytes, level: int = 1) -> bytes:
        lib = cls.get()
        bound = lib.ZSTD_compressBound(len(data))
        buf = ctypes.create_string_buffer(bound)
        n = lib.ZSTD_compress(buf, bound, data, len(data), level)
        if lib.ZSTD_isError(n):
            raise RuntimeError("ZSTD_compress failed")
        return buf.raw[:n]

    @classmethod
    def decompress(cls, data: bytes, raw_len: int) -> bytes:
        lib = cls.get()
        buf = ctypes.create_string_buffer(max(raw_len, 1))
        n = lib.ZSTD_decompress(buf, raw_len, data, len(data))
        if lib.ZSTD_isError(n) or n != raw_len:
            raise ValueError("corrupt zstd block")
        return buf.raw[:raw_len]


class CompressionCodec:
    def __init__(self, name: str):
        self.name = name
        self.id = _CODEC_IDS[name]
        if name == "lz4":
            _Lz4.get()
        elif name == "zstd":
            _Zstd.get()

    @staticmethod
    def available() -> list:
        out = []
        for n in _CODEC_IDS:
            try:
                CompressionCodec(n)
                out.append(n)
            except RuntimeError:
                pass
        return out

    @staticmethod
    def create(name: str) -> "CompressionCodec":
        name = name.lower()
        if name not in _CODEC_IDS:
            raise ValueError(f"unknown compression codec {name!r} (have: {sorted(_CODEC_IDS)})")
        try:
            return CompressionCodec(name)
        except RuntimeError:                 # native library missing on this host: degrade to zlib, never fail a job
            return CompressionCodec("zlib")

    def compress(self, data) -> bytes:
        data = bytes(data)
        if self.name == "none":
            body = data
        elif self.name == "zlib":
            body = zlib.compress(data, 1)
        elif self.name == "lz4":
            body = _Lz4.compress(data)
        elif self.name == "zstd":
            body = _Zstd.compress(data)
        elif self.name == "lzma":
            body = lzma.compress(data, preset=0)
        else:
            body = bz2.compress(data, 1)
        return _HDR.pack(self.id, len(data), zlib.crc32(data)) + body

    @staticmethod
    def decompress(block) -> bytes:
        block = bytes(block)
        cid, raw_len, crc = _HDR.unpack_from(block)
        body = block[_HDR.size:]
        name = _ID_CODECS.get(cid)
        if name is None:
            raise ValueError(f"unknown codec id {cid}")
        if name == "none":
            out = body
        elif name == "zlib":
            out = zlib.decompress(body)
        elif name == "lz4":
            out = _Lz4.decompress(body, raw_len)
        elif name == "zstd":
            out = _Zstd.decompress(body, raw_len)
        elif name == "lzma":
            out = lzma.decompress(body)
        else:
            out = bz2.decompress(body)
        if len(out) != raw_len:
            raise ValueError(f"corrupt {name} block: {len(out)} bytes, header says {raw_len}")
        if zlib.crc32(out) != crc:
            raise ValueError(f"corrupt {name} block: CRC mismatch")
        return out

"""Minimal CBOR (RFC 8949) codec for the `PlayGamesResult` wire format.

The reference pickles `PlayGamesResult` as the bytes `serde_cbor 0.11.2` produces
(rust/src/pybridge.rs:73-92): structs are maps keyed by field name in declaration order, integers
use the shortest head, and an f32 is written as a half float whenever that is lossless, else as a
single.  Only the data model those structs need is implemented (unsigned/negative ints, floats,
text, byte strings, arrays, maps, bool/null).
"""

from __future__ import annotations

import struct
from typing import Any, Tuple

import numpy as np


def _head(major: int, n: int) -> bytes:
    if n < 24:
        return bytes([(major << 5) | n])
    if n < 1 << 8:
        return bytes([(major << 5) | 24, n])
    if n < 1 << 16:
        return bytes([(major << 5) | 25]) + struct.pack(">H", n)
    if n < 1 << 32:
        return bytes([(major << 5) | 26]) + struct.pack(">I", n)
    return bytes([(major << 5) | 27]) + struct.pack(">Q", n)


def encode_f32(v) -> bytes:
    f = np.float32(v)
    if np.isnan(f):
        return b"\xf9\x7e\x00"
    with np.errstate(over="ignore"):
        h = np.float16(f)
    if np.float32(h) == f:  # lossless as a half (includes +-inf and +-0)
        return b"\xf9" + struct.pack(">e", float(h))
    return b"\xfa" + struct.pack(">f", float(f))


class F32(float):
    """Marks a float that must be written with f32 (serde `serialize_f32`) width rules."""


def encode(obj: Any, out: bytearray) -> None:
    if isinstance(obj, bool):
        out += b"\xf5" if obj else b"\xf4"
    elif obj is None:
        out += b"\xf6"
    elif isinstance(obj, (int, np.integer)):
        n = int(obj)
        out += _head(0, n) if n >= 0 else _head(1, -1 - n)
    elif isinstance(obj, (F32, np.float32)):
        out += encode_f32(obj)
    elif isinstance(obj, float):
        out += b"\xfb" + struct.pack(">d", obj)
    elif isinstance(obj, str):
        b = obj.encode("utf-8")
        out += _head(3, len(b)) + b
    elif isinstance(obj, (bytes, bytearray)):
        out += _head(2, len(obj)) + bytes(obj)
    elif isinstance(obj, (list, tuple)):
        out += _head(4, len(obj))
        for x in obj:
            encode(x, out)
    elif isinstance(obj, dict):
        out += _head(5, len(obj))
        for k, v in obj.items():
            encode(k, out)
            encode(v, out)
    else:
        raise ValueError(f"cannot CBOR-encode {type(obj).__name__}")


def dumps(obj: Any) -> bytes:
    out = bytearray()
    encode(obj, out)
    return bytes(out)


def _decode(b: bytes, i: int) -> Tuple[Any, int]:
    if i >= len(b):
        raise ValueError("truncated CBOR")
    ib = b[i]
    major, info = ib >> 5, ib & 31
    i += 1
    if major == 7:
        if info == 20:
            return False, i
        if info == 21:
            return True, i
        if info in (22, 23):
            return None, i
        if info == 25:
            return struct.unpack(">e", b[i : i + 2])[0], i + 2
        if info == 26:
            return struct.unpack(">f", b[i : i + 4])[0], i + 4
        if info == 27:
            return struct.unpack(">d", b[i : i + 8])[0], i + 8
        raise ValueError(f"unsupported CBOR simple value {info}")
    if info < 24:
        n = info
    elif info == 24:
        n, i = b[i], i + 1
    elif info == 25:
        n, i = struct.unpack(">H", b[i : i + 2])[0], i + 2
    elif info == 26:
        n, i = struct.unpack(">I", b[i : i + 4])[0], i + 4
    elif info == 27:
        n, i = struct.unpack(">Q", b[i : i + 8])[0], i + 8
    else:
        raise ValueError("indefinite-length CBOR items are not supported")
    if major == 0:
        return n, i
    if major == 1:
        return -1 - n, i
    if major == 2:
        if i + n > len(b):
            raise ValueError("truncated CBOR")
        return b[i : i + n], i + n
    if major == 3:
        if i + n > len(b):
            raise ValueError("truncated CBOR")
        return b[i : i + n].decode("utf-8"), i + n
    if major == 4:
        out = []
        for _ in range(n):
            v, i = _decode(b, i)
            out.append(v)
        return out, i
    if major == 5:
        d = {}
        for _ in range(n):
            k, i = _decode(b, i)
            v, i = _decode(b, i)
            d[k] = v
        return d, i
    if major == 6:  # tag: ignore the tag number, return the content
        return _decode(b, i)
    raise ValueError("bad CBOR major type")


def loads(b: bytes) -> Any:
    try:
        v, i = _decode(bytes(b), 0)
    except (struct.error, IndexError) as exc:
        raise ValueError(f"malformed CBOR: {exc}") from exc
    if i != len(b):
        raise ValueError("trailing bytes after CBOR item")
    return v

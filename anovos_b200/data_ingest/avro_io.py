"""Avro Object Container Files <-> pyarrow Tables for `read_dataset` / `write_dataset` (reference
data_ingest.py:23-51 lists "avro" among the file types; Spark needs the external spark-avro package for it).
No avro library exists in this image, so the container format (Apache Avro 1.x specification: header `Obj\\x01`,
metadata map with avro.schema / avro.codec, 16-byte sync marker, blocks of <row count, byte size, rows, sync>) and
the binary encoding of flat records are implemented here: null, boolean, int, long (zig-zag varints), float, double,
bytes, string, and [null, T] unions - what spark-avro writes for a DataFrame of primitive columns.  Codecs: null,
deflate, snappy (through pyarrow's codec).  Nested records, arrays, maps, enums and fixed are rejected."""
from __future__ import annotations

import json
import os
import struct
import zlib

import numpy as np

MAGIC = b"Obj\x01"
_PRIMITIVES = ("null", "boolean", "int", "long", "float", "double", "bytes", "string")


class _Reader:
    def __init__(self, buf: bytes, pos: int = 0):
        self.b, self.p = buf, pos

    def long(self) -> int:
        shift = result = 0
        while True:
            byte = self.b[self.p]
            self.p += 1
            result |= (byte & 0x7F) << shift
            if not byte & 0x80:
                break
            shift += 7
        return (result >> 1) ^ -(result & 1)

    def raw(self, n: int) -> bytes:
        out = self.b[self.p:self.p + n]
        if len(out) != n:
            raise ValueError("avro: truncated file")
        self.p += n
        return out


def _field_plan(schema):
    """-> [(name, base type, null branch index | None)] of a flat record schema."""
    if not isinstance(schema, dict) or schema.get("type") != "record":
        raise ValueError("avro: the top-level schema must be a record")
    plan = []
    for f in schema["fields"]:
        t, null_at = f["type"], None
        if isinstance(t, list):
            if len(t) != 2 or "null" not in t:
                raise NotImplementedError("avro: only [null, T] unions are supported (field %r)" % f["name"])
            null_at = t.index("null")
            t = t[1 - null_at]
        if isinstance(t, dict):          # logical types ride on a primitive (date -> int, timestamp-* -> long)
            t = t.get("type")
        if t not in _PRIMITIVES or t == "null":
            raise NotImplementedError("avro: field %r has type %r; only flat primitive columns are supported" % (f["name"], t))
        plan.append((f["name"], t, null_at))
    return plan


def _decompress(codec: str, data: bytes) -> bytes:
    if codec == "null":
        return data
    if codec == "deflate":
        return zlib.decompress(data, -15)
    if codec == "snappy":              # block + 4-byte big-endian CRC32 of the uncompressed bytes
        import pyarrow as pa
        body = data[:-4]
        r = _Reader(body)              # raw snappy starts with the uncompressed length as an unsigned varint
        size = shift = 0
        while True:
            byte = r.b[r.p]
            r.p += 1
            size |= (byte & 0x7F) << shift
            if not byte & 0x80:
                break
            shift += 7
        return pa.decompress(body, decompressed_size=size, codec="snappy", asbytes=True)
    raise NotImplementedError("avro codec %r" % codec)


def read_avro(path):
    """One .avro file -> pyarrow Table (int -> int32, long -> int64, float, double, string, boolean, bytes -> binary)."""
    import pyarrow as pa
    buf = open(path, "rb").read()
    if buf[:4] != MAGIC:
        raise ValueError("%s is not an Avro object container file" % path)
    r = _Reader(buf, 4)
    meta = {}
    while True:
        n = r.long()
        if n == 0:
            break
        if n < 0:
            n = -n
            r.long()
        for _ in range(n):
            k = r.raw(r.long()).decode("utf-8")
            meta[k] = r.raw(r.long())
    sync = r.raw(16)
    plan = _field_plan(json.loads(meta["avro.schema"].decode("utf-8")))
    codec = meta.get("avro.codec", b"null").decode("utf-8")
    cols = [[] for _ in plan]
    while r.p < len(buf):
        count, size = r.long(), r.long()
        blk = _Reader(_decompress(codec, r.raw(size)))
        if r.raw(16) != sync:
            raise ValueError("avro: sync marker mismatch in %s" % path)
        for _ in range(count):
            for out, (_, t, null_at) in zip(cols, plan):
                if null_at is not None and blk.long() == null_at:
                    out.append(None)
                elif t in ("int", "long"):
                    out.append(blk.long())
                elif t == "double":
                    out.append(struct.unpack_from("<d", blk.b, blk.p)[0])
                    blk.p += 8
                elif t == "float":
                    out.append(struct.unpack_from("<f", blk.b, blk.p)[0])
                    blk.p += 4
                elif t == "boolean":
                    out.append(blk.raw(1) != b"\x00")
                elif t == "string":
                    out.append(blk.raw(blk.long()).decode("utf-8"))
                else:
                    out.append(blk.raw(blk.long()))
    types = {"int": pa.int32(), "long": pa.int64(), "float": pa.float32(), "double": pa.float64(), "string": pa.string(),
             "boolean": pa.bool_(), "bytes": pa.binary()}
    return pa.table([pa.array(c, type=types[t]) for c, (_, t, _) in zip(cols, plan)], names=[n for n, _, _ in plan])


def _zigzag(n: int) -> bytes:
    n = (n << 1) ^ (n >> 63)
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def write_avro(table, path, codec="null", rows_per_block=4096):
    """pyarrow Table of primitive columns -> one Avro container file, every field a [T, null] union like spark-avro."""
    import pyarrow as pa
    names, kinds, cols = [], [], []
    for f in table.schema:
        t = f.type
        if pa.types.is_dictionary(t):
            t = t.value_type
        if pa.types.is_int32(t) or pa.types.is_int16(t) or pa.types.is_int8(t):
            k = "int"
        elif pa.types.is_integer(t):
            k = "long"
        elif pa.types.is_float32(t):
            k = "float"
        elif pa.types.is_floating(t):
            k = "double"
        elif pa.types.is_string(t) or pa.types.is_large_string(t):
            k = "string"
        elif pa.types.is_boolean(t):
            k = "boolean"
        else:
            raise NotImplementedError("avro writer: column %r of type %s" % (f.name, t))
        names.append(f.name)
        kinds.append(k)
        cols.append(table.column(f.name).to_pylist())
    schema = {"type": "record", "name": "topLevelRecord",
              "fields": [{"name": n, "type": [k, "null"]} for n, k in zip(names, kinds)]}
    sync = os.urandom(16)
    meta = {"avro.schema": json.dumps(schema).encode("utf-8"), "avro.codec": codec.encode("utf-8")}
    with open(path, "wb") as fh:
        fh.write(MAGIC + _zigzag(len(meta)))
        for k, v in meta.items():
            kb = k.encode("utf-8")
            fh.write(_zigzag(len(kb)) + kb + _zigzag(len(v)) + v)
        fh.write(_zigzag(0) + sync)
        for r0 in range(0, table.num_rows, rows_per_block):
            r1 = min(r0 + rows_per_block, table.num_rows)
            body = bytearray()
            for i in range(r0, r1):
                for k, col in zip(kinds, cols):
                    v = col[i]
                    if v is None:
                        body += _zigzag(1)
                        continue
                    body += _zigzag(0)
                    if k in ("int", "long"):
                        body += _zigzag(int(v))
                    elif k == "double":
                        body += struct.pack("<d", v)
                    elif k == "float":
                        body += struct.pack("<f", v)
                    elif k == "boolean":
                        body += b"\x01" if v else b"\x00"
                    else:
                        b = v.encode("utf-8")
                        body += _zigzag(len(b)) + b
            data = bytes(body)
            if codec == "deflate":
                comp = zlib.compressobj(6, zlib.DEFLATED, -15)
                data = comp.compress(data) + comp.flush()
            elif codec == "snappy":
                raw = data
                data = pa.compress(raw, codec="snappy", asbytes=True) + struct.pack(">I", zlib.crc32(raw) & 0xFFFFFFFF)
            elif codec != "null":
                raise NotImplementedError("avro codec %r" % codec)
            fh.write(_zigzag(r1 - r0) + _zigzag(len(data)) + data + sync)

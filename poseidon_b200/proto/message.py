"""Schema-driven proto2 messages: text format + binary wire format, no protoc needed.

Design notes
------------
* ``Message`` subclasses are generated at import time from :mod:`.schema`.
* Repeated ``float`` fields are stored as ``numpy.float32`` arrays; packed encoding is a
  single ``tobytes()`` / ``frombuffer`` so model weights never go through Python lists.
* The decoder accepts both packed and unpacked encodings for repeated scalars (proto2
  parsers must), and silently skips unknown fields.
* Text format: ``name: value``, ``name { ... }``, ``name < ... >``, ``name: [a, b]``,
  ``#`` comments, single/double-quoted strings with C escapes, adjacent string
  concatenation.

Replaces the reference's protobuf-generated classes and util/io.cpp
(reference: src/caffe/util/io.cpp:31-80 ReadProtoFromTextFile / ReadProtoFromBinaryFile /
WriteProtoToBinaryFile).
"""
from __future__ import annotations

import io
import math
import re
import struct
from typing import Dict, List, Optional

import numpy as np

from . import schema as _schema


class DecodeError(ValueError):
    """Malformed binary protobuf input."""


_VARINT_TYPES = {"int32", "int64", "uint32", "uint64", "bool", "enum", "sint32", "sint64"}
_WT_VARINT, _WT_64, _WT_LEN, _WT_32 = 0, 1, 2, 5


class FieldSpec:
    __slots__ = ("number", "name", "kind", "ref", "label", "default", "packed")

    def __init__(self, number, name, typ, label, default, packed=False):
        self.number, self.name, self.label = number, name, label
        self.default, self.packed = default, packed
        if typ.startswith("enum:"):
            self.kind, self.ref = "enum", typ[5:]
        elif typ.startswith("msg:"):
            self.kind, self.ref = "msg", typ[4:]
        else:
            self.kind, self.ref = typ, None

    @property
    def repeated(self):
        return self.label == "rep"


class RepeatedMessages(list):
    """List of sub-messages with the protobuf-style ``add()`` helper."""

    def __init__(self, cls):
        super().__init__()
        self._cls = cls

    def add(self, **kw):
        m = self._cls(**kw)
        self.append(m)
        return m


def _wire_type(f: FieldSpec) -> int:
    if f.kind in _VARINT_TYPES:
        return _WT_VARINT
    if f.kind in ("float", "fixed32", "sfixed32"):
        return _WT_32
    if f.kind in ("double", "fixed64", "sfixed64"):
        return _WT_64
    return _WT_LEN


def _enc_varint(out: bytearray, v: int) -> None:
    if v < 0:
        v += 1 << 64
    while v > 0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _dec_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _signed(v: int, bits: int) -> int:
    v &= (1 << 64) - 1
    if v >= 1 << 63:
        v -= 1 << 64
    if bits == 32:
        v = ((v + (1 << 31)) % (1 << 32)) - (1 << 31)
    return v


class Message:
    """Base class; concrete classes get ``_fields`` / ``_by_name`` / ``_by_number``."""

    _fields: List[FieldSpec] = []
    _by_name: Dict[str, FieldSpec] = {}
    _by_number: Dict[int, FieldSpec] = {}
    _name = "Message"

    def __init__(self, **kw):
        object.__setattr__(self, "_v", {})
        for k, v in kw.items():
            setattr(self, k, v)

    # ---- attribute protocol -------------------------------------------------------
    def __getattr__(self, name):
        by_name = type(self)._by_name
        if name not in by_name:
            raise AttributeError(f"{type(self)._name} has no field '{name}'")
        f = by_name[name]
        v = self._v
        if name in v:
            return v[name]
        if f.repeated:
            if f.kind == "msg":
                val = RepeatedMessages(_CLASSES[f.ref])
            elif f.kind in ("float", "double"):
                return np.zeros(0, dtype=np.float32 if f.kind == "float" else np.float64)
            else:
                val = []
            v[name] = val
            return val
        if f.kind == "msg":
            # reading an unset sub-message yields a default instance (not stored)
            return _CLASSES[f.ref]()
        return f.default

    def __setattr__(self, name, value):
        by_name = type(self)._by_name
        if name not in by_name:
            raise AttributeError(f"{type(self)._name} has no field '{name}'")
        f = by_name[name]
        self._v[name] = self._coerce(f, value)

    @staticmethod
    def _coerce_scalar(f: FieldSpec, value):
        if f.kind == "enum":
            if isinstance(value, str):
                return _schema.ENUMS[f.ref][value]
            return int(value)
        if f.kind in ("int32", "int64", "uint32", "uint64"):
            return int(value)
        if f.kind in ("float", "double"):
            return float(value)
        if f.kind == "bool":
            return bool(value)
        if f.kind == "string":
            return value.decode("utf-8") if isinstance(value, bytes) else str(value)
        if f.kind == "bytes":
            return value.encode("latin-1") if isinstance(value, str) else bytes(value)
        return value

    def _coerce(self, f: FieldSpec, value):
        if f.repeated:
            if f.kind == "msg":
                lst = RepeatedMessages(_CLASSES[f.ref])
                lst.extend(value)
                return lst
            if f.kind in ("float", "double"):
                return np.ascontiguousarray(
                    value, dtype=np.float32 if f.kind == "float" else np.float64).reshape(-1)
            return [self._coerce_scalar(f, x) for x in value]
        if f.kind == "msg":
            if not isinstance(value, Message):
                raise TypeError(f"field {f.name} expects a message")
            return value
        return self._coerce_scalar(f, value)

    def has(self, name: str) -> bool:
        v = self._v.get(name)
        if v is None:
            return False
        f = type(self)._by_name[name]
        if f.repeated:
            return len(v) > 0
        return True

    HasField = has

    def mutable(self, name: str):
        """Return the sub-message ``name``, creating (and storing) it if unset."""
        f = type(self)._by_name[name]
        assert f.kind == "msg" and not f.repeated
        if name not in self._v:
            self._v[name] = _CLASSES[f.ref]()
        return self._v[name]

    def clear(self, name: str):
        self._v.pop(name, None)

    ClearField = clear

    def append(self, name: str, value):
        """Append one scalar to a repeated field (handles numpy-backed floats)."""
        f = type(self)._by_name[name]
        assert f.repeated
        if f.kind in ("float", "double"):
            cur = getattr(self, name)
            self._v[name] = np.append(cur, np.asarray([value], dtype=cur.dtype))
        elif f.kind == "msg":
            getattr(self, name).append(value)
        else:
            getattr(self, name).append(self._coerce_scalar(f, value))

    def copy(self):
        return type(self).FromString(self.SerializeToString())

    def CopyFrom(self, other):
        object.__setattr__(self, "_v", other.copy()._v)

    def MergeFrom(self, other):
        for f in type(self)._fields:
            if not other.has(f.name):
                continue
            ov = other._v[f.name]
            if f.repeated:
                if f.kind in ("float", "double"):
                    self._v[f.name] = np.concatenate([getattr(self, f.name), ov])
                elif f.kind == "msg":
                    getattr(self, f.name).extend(x.copy() for x in ov)
                else:
                    getattr(self, f.name).extend(ov)
            elif f.kind == "msg":
                self.mutable(f.name).MergeFrom(ov)
            else:
                self._v[f.name] = ov

    def __eq__(self, other):
        return type(self) is type(other) and self.SerializeToString() == other.SerializeToString()

    def __repr__(self):
        return f"<{type(self)._name} {to_text(self, max_floats=8)!r}>"

    def enum_name(self, name: str) -> Optional[str]:
        f = type(self)._by_name[name]
        v = getattr(self, name)
        if v is None:
            return None
        for k, n in _schema.ENUMS[f.ref].items():
            if n == v:
                return k
        return str(v)

    # ---- binary wire format -----------------------------------------------------
    def SerializeToString(self) -> bytes:
        out = bytearray()
        self._encode(out)
        return bytes(out)

    def _encode(self, out: bytearray) -> None:
        for f in sorted(type(self)._fields, key=lambda x: x.number):
            if f.name not in self._v:
                continue
            val = self._v[f.name]
            wt = _wire_type(f)
            if f.repeated:
                if len(val) == 0:
                    continue
                if f.kind in ("float", "double"):
                    arr = np.ascontiguousarray(val, dtype="<f4" if f.kind == "float" else "<f8")
                    if f.packed:
                        _enc_varint(out, (f.number << 3) | _WT_LEN)
                        _enc_varint(out, arr.nbytes)
                        out += arr.tobytes()
                    else:
                        tag = bytearray()
                        _enc_varint(tag, (f.number << 3) | wt)
                        itemsize = arr.itemsize
                        raw = arr.tobytes()
                        # interleave tag + element with numpy (fast for SVProto payloads)
                        n = arr.size
                        rec = np.empty((n, len(tag) + itemsize), dtype=np.uint8)
                        rec[:, : len(tag)] = np.frombuffer(bytes(tag), dtype=np.uint8)
                        rec[:, len(tag):] = np.frombuffer(raw, dtype=np.uint8).reshape(n, itemsize)
                        out += rec.tobytes()
                    continue
                for x in val:
                    self._encode_one(out, f, wt, x)
            else:
                if val is None:
                    continue
                self._encode_one(out, f, wt, val)

    @staticmethod
    def _encode_one(out, f, wt, x):
        _enc_varint(out, (f.number << 3) | wt)
        if wt == _WT_VARINT:
            _enc_varint(out, int(x))
        elif wt == _WT_32:
            out += struct.pack("<f", x)
        elif wt == _WT_64:
            out += struct.pack("<d", x)
        else:
            if f.kind == "msg":
                sub = bytearray()
                x._encode(sub)
                _enc_varint(out, len(sub))
                out += sub
            else:
                b = x.encode("utf-8") if isinstance(x, str) else bytes(x)
                _enc_varint(out, len(b))
                out += b

    @classmethod
    def FromString(cls, data) -> "Message":
        m = cls()
        m._decode_checked(memoryview(data) if not isinstance(data, memoryview) else data)
        return m

    def ParseFromString(self, data):
        object.__setattr__(self, "_v", {})
        self._decode_checked(memoryview(data))
        return self

    def _decode_checked(self, buf) -> None:
        """Top-level entry: malformed wire data (truncated fields, runaway varints, wrong wire types, bad UTF-8) always
        surfaces as DecodeError, as protobuf's ParseFromString does."""
        try:
            self._decode(buf)
        except DecodeError:
            raise
        except (IndexError, ValueError, struct.error, AttributeError, TypeError, OverflowError, RecursionError) as e:
            raise DecodeError(f"{type(self).__name__}: malformed protobuf wire data ({type(e).__name__}: {e})") from e

    def _decode(self, buf) -> None:
        pos, end = 0, len(buf)
        by_number = type(self)._by_number
        v = self._v
        pending: Dict[str, list] = {}
        while pos < end:
            key, pos = _dec_varint(buf, pos)
            num, wt = key >> 3, key & 7
            f = by_number.get(num)
            if wt == _WT_VARINT:
                raw, pos = _dec_varint(buf, pos)
                if f is None:
                    continue
                val = self._from_varint(f, raw)
                if f.repeated:
                    v.setdefault(f.name, []).append(val)
                else:
                    v[f.name] = val
            elif wt == _WT_32:
                chunk = buf[pos:pos + 4]
                pos += 4
                if f is None:
                    continue
                val = struct.unpack("<f", chunk)[0]
                if f.repeated:
                    pending.setdefault(f.name, []).append(val)
                else:
                    v[f.name] = val
            elif wt == _WT_64:
                chunk = buf[pos:pos + 8]
                pos += 8
                if f is None:
                    continue
                val = struct.unpack("<d", chunk)[0]
                if f.repeated:
                    pending.setdefault(f.name, []).append(val)
                else:
                    v[f.name] = val
            elif wt == _WT_LEN:
                n, pos = _dec_varint(buf, pos)
                chunk = buf[pos:pos + n]
                pos += n
                if f is None:
                    continue
                if f.kind == "msg":
                    sub = _CLASSES[f.ref]()
                    sub._decode(chunk)
                    if f.repeated:
                        if f.name not in v:
                            v[f.name] = RepeatedMessages(_CLASSES[f.ref])
                        v[f.name].append(sub)
                    elif f.name in v:
                        v[f.name].MergeFrom(sub)
                    else:
                        v[f.name] = sub
                elif f.kind == "string":
                    s = bytes(chunk).decode("utf-8", errors="replace")
                    if f.repeated:
                        v.setdefault(f.name, []).append(s)
                    else:
                        v[f.name] = s
                elif f.kind == "bytes":
                    if f.repeated:
                        v.setdefault(f.name, []).append(bytes(chunk))
                    else:
                        v[f.name] = bytes(chunk)
                elif f.kind in ("float", "double"):  # packed
                    dt = "<f4" if f.kind == "float" else "<f8"
                    arr = np.frombuffer(chunk, dtype=dt)
                    if f.name in v and len(v[f.name]):
                        arr = np.concatenate([v[f.name], arr])
                    v[f.name] = arr
                else:  # packed varints
                    p2, lst = 0, v.setdefault(f.name, [])
                    while p2 < n:
                        raw, p2 = _dec_varint(chunk, p2)
                        lst.append(self._from_varint(f, raw))
            elif wt == 3 or wt == 4:
                raise ValueError("proto groups are not supported")
            else:
                raise ValueError(f"bad wire type {wt}")
        for name, lst in pending.items():
            f = type(self)._by_name[name]
            arr = np.asarray(lst, dtype=np.float32 if f.kind == "float" else np.float64)
            if name in v and len(v[name]):
                arr = np.concatenate([v[name], arr])
            v[name] = arr

    @staticmethod
    def _from_varint(f, raw):
        if f.kind == "bool":
            return bool(raw)
        if f.kind in ("int32", "enum"):
            return _signed(raw, 32)
        if f.kind == "int64":
            return _signed(raw, 64)
        return raw


_CLASSES: Dict[str, type] = {}


def _build_classes():
    for mname, rows in _schema.MESSAGES.items():
        fields = [FieldSpec(*r) for r in rows]
        cls = type(mname, (Message,), {
            "_fields": fields,
            "_by_name": {f.name: f for f in fields},
            "_by_number": {f.number: f for f in fields},
            "_name": mname,
            "__slots__": ("_v",),
        })
        _CLASSES[mname] = cls


_build_classes()


def get_class(name: str) -> type:
    return _CLASSES[name]


# =====================================================================================
# Text format
# =====================================================================================
_TOKEN_RE = re.compile(
    r"""\s*(?:
        (?P<comment>\#[^\n]*) |
        (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*') |
        (?P<punct>[{}<>\[\]:,;]) |
        (?P<atom>[^\s{}<>\[\]:,;"'\#]+)
    )""",
    re.X,
)

_ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", "'": "'", '"': '"', "0": "\0",
        "a": "\a", "b": "\b", "f": "\f", "v": "\v"}


def _unescape(s: str) -> str:
    body = s[1:-1]
    if "\\" not in body:
        return body
    out, i = [], 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out.append(c)
            i += 1
            continue
        i += 1
        c = body[i]
        if c in _ESC:
            out.append(_ESC[c])
            i += 1
        elif c == "x":
            j = i + 1
            while j < len(body) and j < i + 3 and body[j] in "0123456789abcdefABCDEF":
                j += 1
            out.append(chr(int(body[i + 1:j], 16)))
            i = j
        elif c in "01234567":
            j = i
            while j < len(body) and j < i + 3 and body[j] in "01234567":
                j += 1
            out.append(chr(int(body[i:j], 8)))
            i = j
        else:
            out.append(c)
            i += 1
    return "".join(out)


def _tokenize(text: str):
    pos, n = 0, len(text)
    toks = []
    while pos < n:
        m = _TOKEN_RE.match(text, pos)
        if m is None:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"prototxt: cannot tokenize at offset {pos}: {text[pos:pos+30]!r}")
        pos = m.end()
        if m.lastgroup == "comment":
            continue
        toks.append((m.lastgroup, m.group(m.lastgroup)))
    return toks


class _TextParser:
    def __init__(self, text: str):
        self.toks = _tokenize(text)
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def parse_message(self, msg: Message, closer: Optional[str]):
        cls = type(msg)
        while True:
            kind, tok = self.peek()
            if kind is None:
                if closer is not None:
                    raise ValueError("prototxt: unexpected EOF, missing '%s'" % closer)
                return
            if kind == "punct" and tok == closer:
                self.next()
                return
            if kind == "punct" and tok in ",;":
                self.next()
                continue
            if kind != "atom":
                raise ValueError(f"prototxt: expected field name, got {tok!r}")
            self.next()
            f = cls._by_name.get(tok)
            if f is None:
                raise ValueError(f"prototxt: message {cls._name} has no field '{tok}'")
            kind2, tok2 = self.peek()
            if kind2 == "punct" and tok2 == ":":
                self.next()
                kind2, tok2 = self.peek()
            if f.kind == "msg":
                if not (kind2 == "punct" and tok2 in "{<"):
                    raise ValueError(f"prototxt: expected '{{' after {f.name}")
                self.next()
                sub = _CLASSES[f.ref]()
                self.parse_message(sub, "}" if tok2 == "{" else ">")
                if f.repeated:
                    getattr(msg, f.name).append(sub)
                elif f.name in msg._v:
                    msg._v[f.name].MergeFrom(sub)
                else:
                    msg._v[f.name] = sub
                continue
            if kind2 == "punct" and tok2 == "[":
                self.next()
                while True:
                    k3, t3 = self.peek()
                    if k3 == "punct" and t3 == "]":
                        self.next()
                        break
                    if k3 == "punct" and t3 == ",":
                        self.next()
                        continue
                    self._store(msg, f, self.parse_scalar(f))
                continue
            self._store(msg, f, self.parse_scalar(f))

    @staticmethod
    def _store(msg, f, val):
        if f.repeated:
            if f.kind in ("float", "double"):
                msg._v.setdefault("__pend_" + f.name, []).append(val)
            else:
                getattr(msg, f.name).append(val)
        else:
            msg._v[f.name] = val

    def parse_scalar(self, f: FieldSpec):
        kind, tok = self.next()
        if f.kind in ("string", "bytes"):
            if kind != "str":
                raise ValueError(f"prototxt: field {f.name} expects a quoted string, got {tok!r}")
            s = _unescape(tok)
            while self.peek()[0] == "str":
                s += _unescape(self.next()[1])
            return s if f.kind == "string" else s.encode("latin-1")
        if kind == "str":
            tok = _unescape(tok)
        if f.kind == "enum":
            table = _schema.ENUMS[f.ref]
            if tok in table:
                return table[tok]
            try:
                return int(tok)
            except ValueError:
                raise ValueError(f"prototxt: bad enum value {tok!r} for {f.name} ({f.ref})")
        if f.kind == "bool":
            if tok in ("true", "True", "t", "1"):
                return True
            if tok in ("false", "False", "f", "0"):
                return False
            raise ValueError(f"prototxt: bad bool {tok!r}")
        if f.kind in ("float", "double"):
            t = tok.rstrip("fF") if tok[-1:] in "fF" and tok.lower() not in ("inf", "-inf") else tok
            return float(t)
        return int(tok, 0)


def _finalize_pending(msg: Message):
    for k in [k for k in msg._v if k.startswith("__pend_")]:
        name = k[7:]
        f = type(msg)._by_name[name]
        arr = np.asarray(msg._v.pop(k), dtype=np.float32 if f.kind == "float" else np.float64)
        if name in msg._v and len(msg._v[name]):
            arr = np.concatenate([msg._v[name], arr])
        msg._v[name] = arr
    for f in type(msg)._fields:
        if f.kind == "msg" and f.name in msg._v:
            val = msg._v[f.name]
            if f.repeated:
                for s in val:
                    _finalize_pending(s)
            else:
                _finalize_pending(val)


class ParseError(ValueError):
    """Malformed protobuf text (.prototxt) input."""


def parse_text(text: str, msg_or_cls):
    msg = msg_or_cls() if isinstance(msg_or_cls, type) else msg_or_cls
    try:
        _TextParser(text).parse_message(msg, None)
        _finalize_pending(msg)
    except ParseError:
        raise
    except (ValueError, TypeError, IndexError, KeyError, AttributeError, OverflowError, RecursionError) as e:
        raise ParseError(f"{type(msg).__name__}: {e}") from e
    return msg


def _fmt_float(x: float) -> str:
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    if math.isnan(x):
        return "nan"
    r = repr(float(np.float32(x)))
    s = "%.9g" % x
    # shortest representation that round-trips through float32
    for prec in range(1, 10):
        c = "%.*g" % (prec, x)
        if np.float32(float(c)) == np.float32(x):
            s = c
            break
    del r
    return s


def _escape(s) -> str:
    if isinstance(s, bytes):
        s = s.decode("latin-1")
    out = []
    for ch in s:
        o = ord(ch)
        if ch == "\\":
            out.append("\\\\")
        elif ch == '"':
            out.append('\\"')
        elif ch == "\n":
            out.append("\\n")
        elif o < 32 or o > 126:
            out.append("\\%03o" % o)
        else:
            out.append(ch)
    return "".join(out)


def to_text(msg: Message, indent: int = 0, max_floats: Optional[int] = None) -> str:
    buf = io.StringIO()
    _write_text(msg, buf, indent, max_floats)
    return buf.getvalue()


def _write_text(msg: Message, buf, indent: int, max_floats):
    pad = "  " * indent
    for f in type(msg)._fields:
        if f.name not in msg._v:
            continue
        val = msg._v[f.name]
        items = val if f.repeated else [val]
        if f.repeated and max_floats is not None and f.kind in ("float", "double"):
            items = items[:max_floats]
        for x in items:
            if x is None:
                continue
            if f.kind == "msg":
                buf.write(f"{pad}{f.name} {{\n")
                _write_text(x, buf, indent + 1, max_floats)
                buf.write(f"{pad}}}\n")
            elif f.kind in ("string", "bytes"):
                buf.write(f'{pad}{f.name}: "{_escape(x)}"\n')
            elif f.kind == "enum":
                name = next((k for k, n in _schema.ENUMS[f.ref].items() if n == x), str(x))
                buf.write(f"{pad}{f.name}: {name}\n")
            elif f.kind == "bool":
                buf.write(f"{pad}{f.name}: {'true' if x else 'false'}\n")
            elif f.kind in ("float", "double"):
                buf.write(f"{pad}{f.name}: {_fmt_float(float(x))}\n")
            else:
                buf.write(f"{pad}{f.name}: {int(x)}\n")

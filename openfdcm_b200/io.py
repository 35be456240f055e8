"""`.scene` / `.tmpl` line files: openfdcm.read / openfdcm.write (reference modules/python/src/core.cpp:41-42,
modules/core/include/openfdcm/core/serialization.h:42-132) in pure Python.

Container (packio v0.2.x): 16-byte signature "OPENFDCM" (zero padded), packio version u16 x 3, u8 compressed flag,
u64 uncompressed size, u64 stored size, then the (zlib-compressed) body.  Body = packed 45-byte `LinesSerialHeader`
followed by lineRecordNum x [x1,y1,x2,y2] little-endian float32 (the column-major memory of `LineArray`)."""
import os
import struct
import time
import zlib

import numpy as np

_SIG = b"OPENFDCM".ljust(16, b"\x00")
_PACKIO_VERSION = (0, 2, 0)
_HDR = struct.Struct("<HIHH8s3HHHHIBHQ")   # LinesSerialHeader, packed (serialization.h:42-57)
_VERSION = (0, 10, 0)
assert _HDR.size == 45


def read(filepath):
    """LineArray as a (4, N) float32 array (one column per line)."""
    if not os.path.exists(filepath):
        raise RuntimeError(f"File '{filepath}' does not exist")
    with open(filepath, "rb") as f:
        blob = f.read()
    if len(blob) < 39 or blob[:16] != _SIG:
        raise RuntimeError(f"Cannot open file '{filepath}': bad signature")
    compressed = blob[22]
    usize, ssize = struct.unpack_from("<2Q", blob, 23)
    body = blob[39:39 + ssize]
    if compressed:
        body = zlib.decompress(body)
    if len(body) != usize or len(body) < _HDR.size:
        raise RuntimeError(f"Cannot open file '{filepath}': truncated body")
    hdr = _HDR.unpack_from(body, 0)
    header_size, offset, fmt, rec_len, n = hdr[10], hdr[11], hdr[12], hdr[13], hdr[14]
    if fmt != 0:
        raise RuntimeError(f"Line data format not recognized, found <{rec_len}>")
    data = np.frombuffer(body, "<f4", count=4 * n, offset=offset).reshape(n, 4)
    return np.ascontiguousarray(data.T)


def write(filepath, linearray):
    a = np.asarray(linearray, dtype=np.float32)
    if a.size and (a.ndim != 2 or a.shape[0] != 4):
        raise ValueError(f"expected a (4,N) line array, got shape {a.shape}")
    rec = np.ascontiguousarray(a.T if a.size else np.zeros((0, 4), np.float32), dtype="<f4")
    t = time.gmtime()
    hdr = _HDR.pack(0, 0, 0, 0, b"\x00" * 8, _VERSION[0], _VERSION[1], _VERSION[2], t.tm_yday - 1, t.tm_year - 1900,
                    _HDR.size, _HDR.size, 0, 16, rec.shape[0])
    body = hdr + rec.tobytes()
    comp = zlib.compress(body)
    if os.path.exists(filepath):
        os.remove(filepath)
    with open(filepath, "wb") as f:
        f.write(_SIG + struct.pack("<3H", *_PACKIO_VERSION) + b"\x01" + struct.pack("<2Q", len(body), len(comp)) + comp)

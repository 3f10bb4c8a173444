"""Oracle for the device-side PNG writer (test infrastructure — never imported by product code).

Plain-Python restatement of the file `dl_png_stored` produces: PNG (ISO/IEC 15948) with 8-bit RGB, filter
type 0 on every scanline and a zlib stream (RFC 1950) of STORED deflate blocks (RFC 1951 section 3.2.4); CRC-32 and
Adler-32 from the standard library's zlib.  It stands where the reference calls PIL:
`img.save(buf, format="PNG")`, `backends/cuda_worker.py:234-239` — the job contract is "PNG bytes that decode
to the generated pixels, same seed => same bytes" (`tests/test_sdxl_worker.py:139-198`)."""
import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def png_stored(img: np.ndarray) -> bytes:
    """img uint8 [h, w, 3] -> PNG file bytes."""
    h, w, c = img.shape
    assert c == 3 and img.dtype == np.uint8
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))
    z = bytearray(b"\x78\x01")
    nblk = (len(raw) + 65534) // 65535
    for k in range(nblk):
        blk = raw[k * 65535:(k + 1) * 65535]
        z += bytes([1 if k == nblk - 1 else 0]) + struct.pack("<HH", len(blk), len(blk) ^ 0xFFFF) + blk
    z += struct.pack(">I", zlib.adler32(raw) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            _chunk(b"IDAT", bytes(z)) + _chunk(b"IEND", b""))

"""Test infrastructure: a minimal C3D WRITER (Intel byte order; float32 or scaled int16 point data) used to build
fixtures for the reader tests.  Layout per the public C3D specification (c3d.org): 512-byte header, parameter section
(groups POINT / TRIAL), frame-major point records of 4 words (x, y, z, residual; a negative residual = no data)."""
import struct

import numpy as np


def _group(gid: int, name: str, desc: str = "") -> bytes:
    body = struct.pack("<bb", len(name), -gid) + name.encode()
    rest = struct.pack("<B", len(desc)) + desc.encode()
    return body + struct.pack("<h", 2 + len(rest)) + rest


def _param(gid: int, name: str, dtype: int, dims, data: bytes, desc: str = "") -> bytes:
    body = struct.pack("<bb", len(name), gid) + name.encode()
    rest = struct.pack("<bB", dtype, len(dims)) + bytes(dims) + data + struct.pack("<B", len(desc)) + desc.encode()
    return body + struct.pack("<h", 2 + len(rest)) + rest


def write_c3d(path: str, xyz: np.ndarray, valid: np.ndarray, labels, rate: float = 120.0, as_int16: bool = False,
              scale: float = 0.001, units: str = "m", analog_per_frame: int = 0) -> None:
    """xyz (frames, points, 3) float, valid (frames, points) bool."""
    frames, points = xyz.shape[:2]
    width = max(len(s) for s in labels)
    lab = b"".join(s.ljust(width).encode() for s in labels)
    prm = b""
    prm += _group(1, "POINT", "3-D point parameters")
    prm += _group(2, "TRIAL")
    prm += _param(1, "USED", 2, [], struct.pack("<h", points))
    prm += _param(1, "FRAMES", 2, [], struct.pack("<H", min(frames, 65535)))
    prm += _param(1, "SCALE", 4, [], struct.pack("<f", scale if as_int16 else -abs(scale)))
    prm += _param(1, "RATE", 4, [], struct.pack("<f", rate))
    prm += _param(1, "UNITS", -1, [len(units)], units.encode())
    prm += _param(1, "LABELS", -1, [width, points], lab)
    prm += _param(2, "ACTUAL_START_FIELD", 2, [2], struct.pack("<HH", 1, 0))
    prm += b"\x00\x00"  # end of the parameter records
    nblocks = (4 + len(prm) + 511) // 512
    section = (struct.pack("<BBBB", 1, 0x50, nblocks, 84) + prm).ljust(nblocks * 512, b"\x00")
    data_block = 2 + nblocks
    header = bytearray(512)
    header[0] = 2  # parameter section starts at block 2
    header[1] = 0x50
    struct.pack_into("<HHHHH", header, 2, points, analog_per_frame, 1, min(frames, 65535), 10)
    struct.pack_into("<f", header, 12, scale if as_int16 else -abs(scale))
    struct.pack_into("<HH", header, 16, data_block, 1 if analog_per_frame else 0)
    struct.pack_into("<f", header, 20, rate)
    rec = np.zeros((frames, points * 4 + analog_per_frame), dtype=np.int16 if as_int16 else np.float32)
    pts = rec[:, :points * 4].reshape(frames, points, 4)
    if as_int16:
        pts[..., :3] = np.round(xyz / scale).astype(np.int16)
        pts[..., 3] = np.where(valid, 1, -1)
    else:
        pts[..., :3] = xyz
        pts[..., 3] = np.where(valid, 0.0, -1.0)
    with open(path, "wb") as f:
        f.write(bytes(header))
        f.write(section)
        f.write(rec.tobytes())

"""Scan I/O: convert the reference's data/train_lidar*.mat (MAT v5, loaded by MathWorks libmat in
src/lidar.cpp:17-49) into the engine's packed scan format, without MATLAB libraries.

Format `.scans.u16`: little-endian header  b"PFSCAN1\\0", uint32 n_frames, uint32 n_beams,
then n_frames*n_beams uint16 ranges in millimetres.  The reference's ranges are float32(mm/1000)
exactly (verified for all five datasets), so  float32(u16 / 1000.0)  reproduces them bit-for-bit;
the sensor's invalid-return sentinel 4294967.0 (= UINT32_MAX mm) is stored as 65535.

Usage: python tools/mat2scans.py /root/reference/data/train_lidar0.mat data/_cache/train_lidar0.scans.u16 [max_frames]
"""
import struct
import sys

import numpy as np

MAGIC = b"PFSCAN1\0"
SENTINEL_F32 = np.float32(4294967.0)


def mat_to_f32(path):
    import scipy.io
    m = scipy.io.loadmat(path)
    lid = m["lidar"]
    return np.stack([lid[0, i]["scan"][0, 0].reshape(-1) for i in range(lid.shape[1])]).astype(np.float32)


def encode(scans_f32):
    s = np.asarray(scans_f32, dtype=np.float32)
    mm = np.round(s.astype(np.float64) * 1000.0)
    sent = s == SENTINEL_F32
    if not (mm[~sent] < 65535).all():
        raise ValueError("range above 65.534 m cannot be packed")
    u = np.where(sent, 65535, mm).astype(np.uint16)
    if not np.array_equal(decode(u), s):
        raise ValueError("scan data is not float32(mm/1000): u16 packing would not be exact")
    return u


def decode(u16):
    u = np.asarray(u16, dtype=np.uint16)
    f = (u.astype(np.float64) / 1000.0).astype(np.float32)
    f[u == 65535] = SENTINEL_F32
    return f


def save(path, u16):
    with open(path, "wb") as fh:
        fh.write(MAGIC + struct.pack("<II", u16.shape[0], u16.shape[1]))
        fh.write(np.ascontiguousarray(u16, dtype="<u2").tobytes())


def load(path):
    with open(path, "rb") as fh:
        head = fh.read(16)
        if head[:8] != MAGIC:
            raise ValueError("not a PFSCAN1 file: %s" % path)
        nf, nb = struct.unpack("<II", head[8:])
        u = np.frombuffer(fh.read(), dtype="<u2").reshape(nf, nb)
    return decode(u)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    s = mat_to_f32(src)
    if len(sys.argv) > 3:
        s = s[: int(sys.argv[3])]
    save(dst, encode(s))
    print("wrote %s: %d frames x %d beams" % (dst, s.shape[0], s.shape[1]))

"""Synthetic LiDAR workload of BASELINE.json configs[4] / SURVEY 8(d) "Config 5": a robot driving a
rectangular loop corridor, 1081 beams at -135..+135 deg in 0.25 deg steps (LIDAR_ANGLE, src/kernel.cu:42),
40 Hz.

World: corridor between an outer 30 m x 18 m and an inner 22 m x 10 m rectangle, both centred on the
origin (inside the 40 m map); the robot runs counter-clockwise on the centre line (26 m x 14 m) at
0.5 m/s (0.0125 m per frame), heading tangent, and starts at the reference's initial pose convention:
the trajectory is expressed in the frame of its first pose, so frame 0 is at (0, 0, 0) like the reference's
particles (src/kernel.cu:126-130).  Range = exact ray/segment hit + N(0, 0.01 m), quantised to 1 mm,
clipped to [0.02, 30]; 0.1 % of the returns replaced by the datasets' invalid-return sentinel 4294967.0.
Generator: numpy.random.Generator(PCG64(565)), so the scans are reproducible bit for bit.
"""
import numpy as np

N_BEAMS = 1081
SENTINEL = np.float32(4294967.0)


def _rect_segments(w, h):
    x, y = w / 2.0, h / 2.0
    c = np.array([[-x, -y], [x, -y], [x, y], [-x, y]], np.float64)
    return np.stack([c, np.roll(c, -1, axis=0)], axis=1)            # [4, 2 (a, b), 2 (x, y)]


def centre_line_pose(s, w=26.0, h=14.0):
    """pose (x, y, heading) at arc length s on the counter-clockwise rectangle w x h starting at the
    middle of the bottom side"""
    per = 2 * (w + h)
    s = np.mod(s, per)
    legs = [(w / 2, (0.0, -h / 2), 0.0), (h, (w / 2, -h / 2), np.pi / 2), (w, (w / 2, h / 2), np.pi),
            (h, (-w / 2, h / 2), -np.pi / 2), (w / 2, (-w / 2, -h / 2), 0.0)]
    for length, (x0, y0), hd in legs:
        if s <= length:
            return x0 + s * np.cos(hd), y0 + s * np.sin(hd), hd
        s -= length
    return 0.0, -h / 2, 0.0


def generate(n_frames=2000, speed=0.5, rate_hz=40.0, seed=565, noise=0.01, sentinel_frac=0.001):
    """float32 [n_frames, 1081] ranges and the ground-truth poses [n_frames, 3] in the first pose's frame"""
    rng = np.random.Generator(np.random.PCG64(seed))
    segs = np.concatenate([_rect_segments(30.0, 18.0), _rect_segments(22.0, 10.0)], axis=0)   # 8 walls
    a, b = segs[:, 0], segs[:, 1]
    e = b - a
    ang = np.deg2rad(-135.0 + 0.25 * np.arange(N_BEAMS))
    scans = np.empty((n_frames, N_BEAMS), np.float32)
    poses = np.empty((n_frames, 3), np.float64)
    step = speed / rate_hz
    x0, y0, h0 = centre_line_pose(0.0)
    for f in range(n_frames):
        px, py, hd = centre_line_pose(f * step)
        d = np.stack([np.cos(ang + hd), np.sin(ang + hd)], axis=1)                  # [B, 2]
        # ray p + t d against segment a + u e: t = cross(a - p, e) / cross(d, e), u = cross(a - p, d) / cross(d, e)
        ap = a[None, :, :] - np.array([px, py])[None, None, :]                       # [1, S, 2]
        den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]          # [B, S]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (ap[..., 0] * e[None, :, 1] - ap[..., 1] * e[None, :, 0]) / den
            u = (ap[..., 0] * d[:, None, 1] - ap[..., 1] * d[:, None, 0]) / den
        ok = (np.abs(den) > 1e-12) & (t > 0) & (u >= 0) & (u <= 1)
        r = np.where(ok, t, np.inf).min(axis=1)
        r = r + rng.normal(0.0, noise, N_BEAMS)
        r = np.clip(np.round(r * 1000.0) / 1000.0, 0.02, 30.0)
        out = r.astype(np.float32)
        out[rng.random(N_BEAMS) < sentinel_frac] = SENTINEL
        scans[f] = out
        c, s_ = np.cos(-h0), np.sin(-h0)
        poses[f] = [c * (px - x0) - s_ * (py - y0), s_ * (px - x0) + c * (py - y0), hd - h0]
    return scans, poses


if __name__ == "__main__":
    import sys
    from . import scans as S
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    sc, _ = generate(n)
    S.save(sys.argv[1], S.encode(sc))
    print("wrote %s: %d frames x %d beams" % (sys.argv[1], sc.shape[0], sc.shape[1]))

"""Seeded synthetic spinning-LiDAR scans for the BASELINE.json configs.

Stands in for the reference's synthetic acquisition (src/SlamCore/experimental/synthetic.cxx:293-380:
sensor moving along a trajectory, per-point timestamps, points expressed in the sensor frame at their own
acquisition instant) with an analytic urban scene that is ray-cast exactly, so scans have the structure of a
64-beam KITTI sweep (rings, range distribution, ~120k returns). numpy only; used by tests and bench.py.

Coordinates are rounded to float32 like real sensor drivers emit them (KITTI .bin, PointCloud2 FLOAT32).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class SensorModel:
    name: str
    n_rings: int
    elev_max_deg: float
    elev_min_deg: float
    n_azimuth: int
    period: float        # seconds per revolution
    min_range: float
    max_range: float
    height: float        # sensor height above ground
    elev_table_deg: tuple = None   # explicit ring elevations (top to bottom); None = uniform between max and min


HDL64 = SensorModel("HDL-64", 64, 2.0, -24.8, 2083, 0.1, 3.0, 100.0, 1.73)       # KITTI-shape, config 2/3
HDL32 = SensorModel("HDL-32", 32, 10.67, -30.67, 2170, 0.1, 1.0, 100.0, 1.0)     # NCLT-shape, config 4
DENSE128 = SensorModel("DENSE-128", 128, 15.0, -25.0, 2400, 0.1, 2.0, 120.0, 1.8)  # config 5
# the real HDL-64E: two laser blocks, 32 rings at 1/3 degree over +2 .. -8.33 and 32 rings at 1/2 degree below (the uniform
# HDL64 model above puts only 24 rings in the upper block, i.e. a quarter fewer far returns)
HDL64E = SensorModel("HDL-64E", 64, 2.0, -24.33, 2083, 0.1, 3.0, 100.0, 1.73,
                     tuple(np.concatenate([np.linspace(2.0, -8.33, 32), np.linspace(-8.83, -24.33, 32)]).tolist()))
SMALL16 = SensorModel("SMALL-16", 16, 15.0, -15.0, 900, 0.1, 1.0, 60.0, 1.5)      # ~10k pts, config 1 stand-in


_RAYCAST = None


def _raycast_lib():
    """tools/libraycast.so (built on demand with the system gcc; data generation only)."""
    global _RAYCAST
    if _RAYCAST is None:
        import ctypes
        import os
        import subprocess
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        src = os.path.join(root, "tools", "raycast.c")
        lib = os.path.join(root, "tools", "libraycast.so")
        if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
            base = [cc, "-O2", "-fPIC", "-shared", src, "-o", lib, "-lm"]
            if subprocess.run(base[:4] + ["-fopenmp"] + base[4:], capture_output=True).returncode != 0:
                subprocess.run(base, check=True, capture_output=True)
        L = ctypes.CDLL(lib)
        L.raycast_scene.restype = None
        L.raycast_scene.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double,
                                    ctypes.c_uint64, ctypes.c_void_p]
        _RAYCAST = L
    return _RAYCAST


class UrbanScene:
    """Ground plane z=0, axis-aligned boxes along a street (buildings, cars, fences = solid; bushes and tree
    crowns = porous, returning at a random depth), vertical cylinders (poles / trunks)."""

    def __init__(self, seed=1234, length=2000.0, profile="street"):
        """profile "street": a narrow street canyon (the round-1 scene: ~6.7k occupied 0.5 m voxels per HDL-64 sweep);
        "suburb": open lots, set-back buildings in two rows, parked cars, dense vegetation — the sweep then fills the
        25-40k 0.5 m voxels / 3-6k 1.5 m keypoint voxels SURVEY.md §8 quotes for real KITTI sweeps."""
        rng = np.random.default_rng(seed)
        boxes = []   # x0,y0,z0,x1,y1,z1,mean_free_path
        cyl = []     # cx,cy,radius,height
        self.profile = profile
        if profile == "suburb":
            self._build_suburb(rng, length, boxes, cyl)
            self._finish(boxes, cyl, seed)
            return
        if profile != "street":
            raise ValueError("unknown scene profile %r" % profile)
        x = -120.0
        while x < length:
            for side in (-1.0, 1.0):
                if rng.random() < 0.85:
                    w = rng.uniform(8.0, 25.0)
                    d = rng.uniform(6.0, 15.0)
                    h = rng.uniform(5.0, 20.0)
                    setback = rng.uniform(9.0, 16.0)
                    y0 = side * setback if side > 0 else side * setback - d
                    boxes.append([x, y0, 0.0, x + w, y0 + d, h, 0.0])
            x += rng.uniform(14.0, 30.0)
        xc = -100.0
        while xc < length:   # parked cars
            side = rng.choice([-1.0, 1.0])
            y0 = side * rng.uniform(4.5, 6.5)
            boxes.append([xc, y0 - 0.9, 0.0, xc + 4.2, y0 + 0.9, 1.5, 0.0])
            xc += rng.uniform(8.0, 25.0)
        xv = -110.0
        while xv < length:   # vegetation / clutter
            side = rng.choice([-1.0, 1.0])
            kind = rng.random()
            y = side * rng.uniform(6.5, 30.0)
            if kind < 0.45:      # bush (porous)
                sx, sy, sz = rng.uniform(0.8, 2.5, size=3)
                boxes.append([xv, y, 0.0, xv + sx, y + sy, sz, 0.5])
            elif kind < 0.75:    # tree: porous crown + trunk
                sx, sy = rng.uniform(2.0, 5.0, size=2)
                z0 = rng.uniform(2.5, 4.0)
                boxes.append([xv, y, z0, xv + sx, y + sy, z0 + rng.uniform(2.0, 5.0), 0.9])
                cyl.append([xv + sx / 2, y + sy / 2, rng.uniform(0.15, 0.4), z0 + 0.1])
            elif rng.random() < 0.5:   # fence along the street
                boxes.append([xv, y, 0.0, xv + rng.uniform(3.0, 9.0), y + 0.2, rng.uniform(0.8, 2.0), 0.0])
            else:                      # wall segment across the street direction
                boxes.append([xv, y, 0.0, xv + 0.2, y + rng.uniform(3.0, 9.0), rng.uniform(0.8, 2.0), 0.0])
            xv += rng.uniform(0.5, 1.5)
        xp = -110.0
        while xp < length:   # poles
            for side in (-1.0, 1.0):
                cyl.append([xp + rng.uniform(-2, 2), side * rng.uniform(7.0, 8.5), rng.uniform(0.12, 0.35),
                            rng.uniform(4.0, 9.0)])
            xp += rng.uniform(10.0, 20.0)
        self._finish(boxes, cyl, seed)

    def _finish(self, boxes, cyl, seed):
        b = np.asarray(boxes, dtype=np.float64)
        self.boxes = np.ascontiguousarray(b[np.argsort(b[:, 0], kind="stable")])
        self.cyl = np.ascontiguousarray(np.asarray(cyl, dtype=np.float64))
        self.seed = seed

    @staticmethod
    def _build_suburb(rng, length, boxes, cyl):
        """Tuned (tools/scene_stats.py) for the number of occupied 0.5 m voxels per HDL-64E sweep: what counts is how far
        the upper beams travel before they hit something — dense clutter next to the road occludes everything behind it
        and LOWERS the count — so: two rows of set-back buildings, parked cars, vegetation spread over the lots, lawns as
        low porous slabs (rough ground)."""
        x = -150.0
        while x < length:   # two rows of buildings per side
            for side in (-1.0, 1.0):
                for (s0, s1, prob) in ((25.0, 40.0, 0.9), (55.0, 85.0, 0.9)):
                    if rng.random() < prob:
                        w = rng.uniform(10.0, 26.0)
                        d = rng.uniform(8.0, 18.0)
                        h = rng.uniform(4.0, 14.0)
                        setback = rng.uniform(s0, s1)
                        y0 = side * setback if side > 0 else side * setback - d
                        boxes.append([x + rng.uniform(-3, 3), y0, 0.0, x + w, y0 + d, h, 0.0])
            x += rng.uniform(14.0, 24.0)
        xc = -120.0
        while xc < length:   # parked cars, both sides
            for side in (-1.0, 1.0):
                if rng.random() < 0.7:
                    y0 = side * rng.uniform(4.5, 6.5)
                    boxes.append([xc, y0 - 0.9, 0.0, xc + rng.uniform(3.8, 4.8), y0 + 0.9, rng.uniform(1.4, 1.9), 0.0])
            xc += rng.uniform(6.0, 12.0)
        xv = -140.0
        while xv < length:   # vegetation / clutter scattered over the lots
            side = rng.choice([-1.0, 1.0])
            kind = rng.random()
            y = side * rng.uniform(8.0, 70.0)
            if kind < 0.30:      # bush (porous)
                sx, sy, sz = rng.uniform(1.0, 3.0, size=3)
                boxes.append([xv, y, 0.0, xv + sx, y + sy, sz, 1.0])
            elif kind < 0.62:    # tree: porous crown + trunk
                sx, sy = rng.uniform(3.0, 8.0, size=2)
                z0 = rng.uniform(1.8, 3.0)
                boxes.append([xv, y, z0, xv + sx, y + sy, z0 + rng.uniform(2.5, 7.0), 2.0])
                cyl.append([xv + sx / 2, y + sy / 2, rng.uniform(0.15, 0.45), z0 + 0.1])
            elif kind < 0.80:    # hedge / tall grass strip (porous, low)
                boxes.append([xv, y, 0.0, xv + rng.uniform(2.0, 10.0), y + rng.uniform(0.6, 3.0), rng.uniform(0.4, 1.4), 0.6])
            elif kind < 0.90:    # fence along the street
                boxes.append([xv, y, 0.0, xv + rng.uniform(3.0, 12.0), y + 0.2, rng.uniform(0.8, 2.0), 0.0])
            else:                # wall segment across the street direction
                boxes.append([xv, y, 0.0, xv + 0.2, y + rng.uniform(3.0, 12.0), rng.uniform(0.8, 2.2), 0.0])
            xv += rng.uniform(0.3, 0.7)
        xg = -140.0
        while xg < length:   # lawns / verges: low porous slabs (returns scatter over 0..h above the plane)
            for side in (-1.0, 1.0):
                if rng.random() < 0.8:
                    w, d = rng.uniform(8.0, 22.0, size=2)
                    y0 = side * rng.uniform(3.8, 40.0)
                    y0 = y0 if side > 0 else y0 - d
                    boxes.append([xg, y0, 0.0, xg + w, y0 + d, rng.uniform(0.25, 0.7), rng.uniform(1.5, 5.0)])
            xg += rng.uniform(6.0, 14.0)
        xp = -140.0
        while xp < length:   # poles / sign posts
            for side in (-1.0, 1.0):
                cyl.append([xp + rng.uniform(-2, 2), side * rng.uniform(7.0, 8.5), rng.uniform(0.08, 0.3),
                            rng.uniform(3.0, 9.0)])
            xp += rng.uniform(8.0, 16.0)

    def raycast(self, o, d, max_range, ray_seed=0):
        """o, d: (n,3) origins / unit directions in the world. Returns range (n,), inf where nothing is hit."""
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        out = np.empty(len(o), dtype=np.float64)
        _raycast_lib().raycast_scene(o.ctypes.data, d.ctypes.data, len(o), self.boxes.ctypes.data, len(self.boxes),
                                     self.cyl.ctypes.data, len(self.cyl), float(max_range),
                                     (self.seed * 1000003 + ray_seed) & 0xFFFFFFFFFFFFFFFF, out.ctypes.data)
        return out


def _rotz(a):
    c, s = np.cos(a), np.sin(a)
    R = np.zeros(a.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    return R


def _roty(a):
    c, s = np.cos(a), np.sin(a)
    R = np.zeros(a.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 2], R[..., 2, 0], R[..., 2, 2], R[..., 1, 1] = c, s, -s, c, 1.0
    return R


class Trajectory:
    """Ground-truth sensor pose as a smooth function of time."""

    def __init__(self, speed=10.0, sway=2.0, sway_rate=0.03, pitch_amp=0.01, pitch_rate=1.1, height=1.73,
                 yaw_jerk=0.0, ramp=1.5):
        self.speed, self.sway, self.sway_rate, self.ramp = speed, sway, sway_rate, ramp
        self.pitch_amp, self.pitch_rate, self.height, self.yaw_jerk = pitch_amp, pitch_rate, height, yaw_jerk

    def _x(self, t):
        # starts at rest and accelerates to `speed` with time constant `ramp` (vehicles start from standstill)
        if self.ramp <= 0:
            return self.speed * t
        return self.speed * (t - self.ramp * (1.0 - np.exp(-t / self.ramp)))

    def _vx(self, t):
        if self.ramp <= 0:
            return np.full_like(t, self.speed)
        return self.speed * (1.0 - np.exp(-t / self.ramp))

    def position(self, t):
        t = np.asarray(t, dtype=np.float64)
        x = self._x(t)
        return np.stack([x, self.sway * np.sin(self.sway_rate * x), np.full_like(t, self.height)], -1)

    def rotation(self, t):
        t = np.asarray(t, dtype=np.float64)
        yaw = np.arctan(self.sway * self.sway_rate * np.cos(self.sway_rate * self._x(t)))   # heading follows the path
        if self.yaw_jerk:
            yaw = yaw + self.yaw_jerk * np.sin(2.3 * t) * np.sin(0.7 * t)
        pitch = self.pitch_amp * np.sin(self.pitch_rate * t)
        return _rotz(yaw) @ _roty(pitch)


def generate_scan(scene, traj, sensor, frame_idx, seed=1234, noise_sigma=0.02, relative_to=None):
    """One sweep. Returns dict(xyz (n,3) f64 [float32-rounded] in the sensor frame at acquisition time,
    t (n,) f64 timestamps, gt_begin/gt_end: 4x4 ground-truth poses at the sweep boundaries)."""
    rng = np.random.default_rng(seed + 7919 * frame_idx)
    t0 = frame_idx * sensor.period
    az_idx = np.arange(sensor.n_azimuth)
    elev = np.deg2rad(np.asarray(sensor.elev_table_deg) if sensor.elev_table_deg is not None
                      else np.linspace(sensor.elev_max_deg, sensor.elev_min_deg, sensor.n_rings))
    az = -2.0 * np.pi * az_idx / sensor.n_azimuth        # clockwise spin like a Velodyne
    frac = (az_idx + 0.5) / sensor.n_azimuth
    AZ, EL = np.meshgrid(az, elev, indexing="xy")          # (rings, az)
    FR = np.broadcast_to(frac[None, :], AZ.shape)
    d_s = np.stack([np.cos(EL) * np.cos(AZ), np.cos(EL) * np.sin(AZ), np.sin(EL)], -1).reshape(-1, 3)
    t = (t0 + FR * sensor.period).reshape(-1)
    R = traj.rotation(t)
    o = traj.position(t)
    d_w = np.einsum("nij,nj->ni", R, d_s)
    rng_ = scene.raycast(o, d_w, sensor.max_range, ray_seed=frame_idx)
    rng_ = rng_ + rng.normal(0.0, noise_sigma, size=rng_.shape)
    keep = np.isfinite(rng_) & (rng_ > sensor.min_range) & (rng_ < sensor.max_range)
    xyz = (d_s[keep] * rng_[keep, None]).astype(np.float32).astype(np.float64)
    ts = t[keep]

    def pose(tt):
        T = np.eye(4)
        T[:3, :3] = traj.rotation(np.asarray(tt))
        T[:3, 3] = traj.position(np.asarray(tt))
        return T

    return {"xyz": xyz, "t": ts, "gt_begin": pose(ts.min()), "gt_end": pose(ts.max()), "frame_idx": frame_idx}


def make_sequence(n_frames, sensor=HDL64, seed=1234, noise_sigma=0.02, traj=None, scene=None, start=0):
    scene = scene or UrbanScene(seed)
    traj = traj or Trajectory(height=sensor.height)
    return [generate_scan(scene, traj, sensor, start + i, seed, noise_sigma) for i in range(n_frames)]


def relative_pose(T_ref, T):
    """T_ref^-1 · T."""
    return np.linalg.inv(T_ref) @ T

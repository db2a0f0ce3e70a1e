"""Seeded synthetic inputs of the BASELINE.json shapes. Host-side numpy only: these are test and
benchmark INPUTS, generated identically for the CPU oracle and the GPU path; nothing here is on the product path."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

# BASELINE.json configs: sensors, rings x azimuth steps, crop passes (axis, lo, hi, negative), leaf, min_points
ROI_BOX = [(2, -0.5, 3.0, 0), (1, -5.0, 5.0, 0), (0, -15.0, 60.0, 0)]   # Parameter.h:31-35
WIDE_BOX = [(2, -60.0, 60.0, 0), (1, -60.0, 60.0, 0), (0, -60.0, 60.0, 0)]
CONFIGS: Dict[str, dict] = {
    "cfg1": dict(sensors=2, rings=64, azimuth=1024, passes=[(2, -0.5, 3.0, 0)], leaf=0.1, min_points=2,
                 what="2 x 64k XYZI, concat + PassThrough z + VoxelGrid 0.1 m"),
    "cfg2": dict(sensors=4, rings=128, azimuth=1024, passes=ROI_BOX, leaf=0.05, min_points=2,
                 what="4 sensors x 128k pts, transform+concat+box crop+VoxelGrid 0.05 m"),
    "cfg3": dict(sensors=8, rings=128, azimuth=2048, passes=ROI_BOX, leaf=0.05, min_points=2,
                 what="8 sensors x 256k pts per frame, frame-sharded"),
}


def _rng(seed: int, sensor: int, frame: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[(seed * 1000003 + sensor) & (2**63 - 1), frame & (2**63 - 1)]))


def lidar_cloud(seed: int, sensor: int, frame: int, rings: int, azimuth: int, nan_frac: float = 0.0) -> np.ndarray:
    """Spinning-lidar pattern in the sensor frame: rings x azimuth points, ground plane at z = -1.8 m, else
    log-uniform range in [2, 120] m, +-2 cm range noise, intensity uniform [0, 255]. Returns (N, 4) float32."""
    rng = _rng(seed, sensor, frame)
    n = rings * azimuth
    elev = np.deg2rad(np.linspace(-25.0, 15.0, rings))[:, None] + np.zeros((1, azimuth))
    az = (np.arange(azimuth)[None, :] + rng.random((rings, azimuth))) * (2.0 * math.pi / azimuth)
    dx, dy, dz = np.cos(elev) * np.cos(az), np.cos(elev) * np.sin(az), np.sin(elev)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dz < 0, -1.8 / dz, np.inf)
    hit = (t_ground >= 0.5) & (t_ground <= 120.0)
    free = np.exp(rng.uniform(math.log(2.0), math.log(120.0), size=(rings, azimuth)))
    obstacle = rng.random((rings, azimuth)) < 0.35      # some returns come from objects before the ground
    rng_m = np.where(hit & ~obstacle, t_ground, np.where(hit, np.minimum(free, t_ground), free))
    rng_m = rng_m + rng.normal(0.0, 0.02, size=(rings, azimuth))
    out = np.empty((n, 4), np.float32)
    out[:, 0] = (rng_m * dx).reshape(-1)
    out[:, 1] = (rng_m * dy).reshape(-1)
    out[:, 2] = (rng_m * dz).reshape(-1)
    out[:, 3] = rng.uniform(0.0, 255.0, size=n)
    if nan_frac > 0:
        bad = rng.random(n) < nan_frac
        out[bad, 0:3] = np.nan
    return out


def extrinsic(sensor: int, n_sensors: int) -> np.ndarray:
    """Sensor s at yaw s*360/S on a 1.2 m ring, 1.9 m high, 2 deg pitch and roll. 4x4 float32, row-major."""
    yaw = 2.0 * math.pi * sensor / n_sensors
    pitch = roll = math.radians(2.0)
    cz, sz, cy, sy, cx, sx = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    m = np.eye(4)
    m[:3, :3] = rz @ ry @ rx
    m[:3, 3] = [1.2 * cz, 1.2 * sz, 1.9]
    return m.astype(np.float32)


def pack_cloud(xyzi: np.ndarray, point_step: int, off_x: int, off_y: int, off_z: int, off_intensity: int,
               fill: int = 0xA5) -> np.ndarray:
    """(N, 4) float32 -> PointCloud2-style byte records (uint8, N * point_step); unused bytes get a filler pattern."""
    n = len(xyzi)
    rec = np.full((n, point_step), fill, np.uint8)
    src = np.ascontiguousarray(xyzi, np.float32).view(np.uint8).reshape(n, 16)
    for k, off in enumerate((off_x, off_y, off_z, off_intensity)):
        if off is not None and off >= 0:
            rec[:, off:off + 4] = src[:, 4 * k:4 * k + 4]
    return rec.reshape(-1)


def frame_clouds(cfg: str, seed: int, frame: int, nan_frac: float = 0.0) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """All sensor clouds + extrinsics of one frame of a named config."""
    c = CONFIGS[cfg]
    clouds = [lidar_cloud(seed, s, frame, c["rings"], c["azimuth"], nan_frac) for s in range(c["sensors"])]
    mats = [extrinsic(s, c["sensors"]) for s in range(c["sensors"])]
    return clouds, mats


def map_cloud(seed: int, n: int, extent=(400.0, 400.0, 20.0), n_boxes: int = 2000, chunk: int = 1 << 22) -> np.ndarray:
    """cfg 4 aggregated map: 70 % of the points on ground + axis-aligned box faces, 30 % uniform. (n, 4) float32."""
    rng = _rng(seed, 0, 0)
    ex, ey, ez = extent
    boxes_c = rng.uniform([-ex / 2, -ey / 2, 0], [ex / 2, ey / 2, 0], size=(n_boxes, 3))
    boxes_s = rng.uniform([1, 1, 1], [12, 12, ez * 0.6], size=(n_boxes, 3))
    out = np.empty((n, 4), np.float32)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        kind = rng.random(m)
        p = rng.uniform([-ex / 2, -ey / 2, -ez / 2], [ex / 2, ey / 2, ez / 2], size=(m, 3))
        ground = kind < 0.35
        p[ground, 2] = -ez / 2 + rng.normal(0, 0.02, size=int(ground.sum()))
        onbox = (kind >= 0.35) & (kind < 0.70)
        k = int(onbox.sum())
        b = rng.integers(0, n_boxes, size=k)
        q = boxes_c[b] + (rng.random((k, 3)) - 0.5) * boxes_s[b]
        face = rng.integers(0, 3, size=k)
        side = rng.integers(0, 2, size=k) * 2 - 1
        q[np.arange(k), face] = boxes_c[b, face] + side * boxes_s[b, face] / 2
        q[:, 2] = q[:, 2] - ez / 2 + boxes_s[b, 2] / 2
        p[onbox] = q
        out[lo:lo + m, 0:3] = p
        out[lo:lo + m, 3] = rng.uniform(0, 255, size=m)
    return out


def uniform_cloud(seed: int, n: int, extent=(200.0, 200.0, 10.0)) -> np.ndarray:
    """cfg 5 leaf-sweep cloud: uniform in a box centred on the origin. (n, 4) float32."""
    rng = _rng(seed, 0, 0)
    out = np.empty((n, 4), np.float32)
    for k in range(3):
        out[:, k] = rng.uniform(-extent[k] / 2, extent[k] / 2, size=n)
    out[:, 3] = rng.uniform(0, 255, size=n)
    return out

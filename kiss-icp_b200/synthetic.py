"""Seeded synthetic LiDAR streams (KITTI-shape 64x1024, Ouster-128-shape 128x1024).

There is no dataset access in this environment, so every workload in BASELINE.json is a
procedurally generated scan stream: a ray-cast spinning LiDAR driving a closed loop through a
box city. Conventions reused from the reference's loaders (for shape only, no code):
  * KITTI odometry scans are float32 on disk and come WITHOUT per-point stamps
    (python/kiss_icp/datasets/kitti.py:57,66)  -> ``stamps='none'``, coords rounded to fp32
  * MulRan/Ouster style per-column stamps floor(i/H)/W (python/kiss_icp/datasets/mulran.py:54-58)
    -> ``stamps='column'`` (points are then motion-distorted so that deskew matters)

The ray caster is written with torch ops so the same code runs on the CPU (small test scans)
and on the GPU (bench-size streams).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th**2 * (K @ K)


def _rpy(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


class SyntheticLidar:
    """One seeded sequence. ``scan(k)`` -> (points (N,3) float64, timestamps (N,) or (0,))."""

    def __init__(self, seed=0, beams=64, cols=1024, elev_deg=(-24.8, 2.0), max_ray=120.0, noise=0.02,
                 stamps="none", speed=10.0, hz=10.0, loop_radius=150.0, sensor_height=1.73, n_boxes=260,
                 device="cpu"):
        self.seed, self.beams, self.cols = int(seed), int(beams), int(cols)
        self.max_ray, self.noise, self.stamps = float(max_ray), float(noise), stamps
        self.device = torch.device(device)
        self.sensor_height = sensor_height
        rng = np.random.default_rng(self.seed)
        # trajectory: wobbly closed loop, constant arc speed
        self.R = loop_radius
        self.ds = speed / hz
        self.ecc = 0.25 + 0.1 * rng.random()
        self.wob_a = 12.0 * rng.random()
        self.wob_k = int(rng.integers(2, 5))
        self.phase = 2 * math.pi * rng.random()
        # boxes: jittered grid, removed where they touch the driving corridor
        g = int(math.ceil(math.sqrt(n_boxes)))
        half = self.R * 1.6
        cx, cy = np.meshgrid(np.linspace(-half, half, g), np.linspace(-half, half, g))
        ctr = np.stack([cx.ravel(), cy.ravel()], 1) + rng.uniform(-6, 6, size=(g * g, 2))
        size = rng.uniform(6.0, 24.0, size=(g * g, 2))
        height = rng.uniform(3.0, 25.0, size=(g * g,))
        # small clutter (cars / poles) near the road
        n_small = n_boxes // 2
        ang = rng.uniform(0, 2 * math.pi, n_small)
        road = self._xy(ang * self.R)  # param by arclength-ish
        off = rng.choice([-1.0, 1.0], n_small) * rng.uniform(4.0, 9.0, n_small)
        nrm = self._normal(ang * self.R)
        sctr = road + nrm * off[:, None]
        ssize = np.where(rng.random((n_small, 1)) < 0.5, rng.uniform(1.5, 4.5, (n_small, 2)), rng.uniform(0.2, 0.5, (n_small, 2)))
        sheight = np.where(ssize[:, 0] < 1.0, rng.uniform(3.0, 8.0, n_small), rng.uniform(1.2, 2.2, n_small))
        ctr = np.concatenate([ctr, sctr])
        size = np.concatenate([size, ssize])
        height = np.concatenate([height, sheight])
        lo = np.concatenate([ctr - size / 2, np.zeros((len(ctr), 1))], 1)
        hi = np.concatenate([ctr + size / 2, height[:, None]], 1)
        # corridor clearance
        s_samples = np.linspace(0, 2 * math.pi * self.R, 2000)
        path = self._xy(s_samples)
        keep = np.ones(len(lo), bool)
        for i in range(len(lo)):
            d = np.maximum(np.maximum(lo[i, :2] - path, path - hi[i, :2]), 0.0)
            if (np.hypot(d[:, 0], d[:, 1]) < 3.0).any():
                keep[i] = False
        self.box_lo = torch.tensor(lo[keep], dtype=torch.float64, device=self.device)
        self.box_hi = torch.tensor(hi[keep], dtype=torch.float64, device=self.device)
        # ray directions in the sensor frame, column-major point order (all beams of column 0, ...)
        el = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], self.beams))
        az = -2 * math.pi * np.arange(self.cols) / self.cols  # clockwise sweep
        ce, se = np.cos(el), np.sin(el)
        d = np.stack([np.outer(np.cos(az), ce), np.outer(np.sin(az), ce), np.outer(np.ones_like(az), se)], -1)
        self.dirs = torch.tensor(d.reshape(-1, 3), dtype=torch.float64, device=self.device)  # (cols*beams, 3)
        self.col_of_point = np.repeat(np.arange(self.cols), self.beams)

    # -- trajectory --------------------------------------------------------------------
    def _theta(self, s):
        return np.asarray(s) / self.R

    def _xy(self, s):
        th = self._theta(s)
        r = self.R * (1 + self.ecc * 0.3 * np.cos(2 * th)) + self.wob_a * np.sin(self.wob_k * th + self.phase)
        return np.stack([r * np.cos(th), r * np.sin(th)], -1)

    def _normal(self, s):
        e = 1e-3
        t = (self._xy(np.asarray(s) + e) - self._xy(np.asarray(s) - e)) / (2 * e)
        t /= np.linalg.norm(t, axis=-1, keepdims=True)
        return np.stack([-t[..., 1], t[..., 0]], -1)

    def pose(self, k: float) -> np.ndarray:
        """Ground-truth sensor pose (4x4, world<-sensor) at (fractional) scan index k."""
        # the vehicle starts from rest and approaches cruise speed exponentially (tau = 10 scans)
        tau = 10.0
        s = self.ds * (k + tau * (math.exp(-max(k, 0.0) / tau) - 1.0)) if k > 0 else 0.0
        e = 1e-3
        p = self._xy(s)
        t = (self._xy(s + e) - self._xy(s - e)) / (2 * e)
        yaw = math.atan2(t[1], t[0])
        roll = math.radians(0.8) * math.sin(0.05 * s + self.phase)
        pitch = math.radians(0.6) * math.sin(0.08 * s + 1.0)
        T = np.eye(4)
        T[:3, :3] = _rpy(roll, pitch, yaw)
        T[:3, 3] = [p[0], p[1], self.sensor_height + 0.03 * math.sin(0.11 * s)]
        return T

    # -- ray casting -------------------------------------------------------------------
    def _cast(self, origins, dirs):
        """origins/dirs (N,3) world frame -> range (N,), inf where nothing is hit."""
        big = float("inf")
        dz = dirs[:, 2]
        t = torch.where(dz < -1e-9, -origins[:, 2] / dz, torch.full_like(dz, big))
        # cull boxes outside the sensing disc
        c = origins.mean(0)
        near = ((self.box_lo[:, 0] - c[0]).clamp(min=0) + (c[0] - self.box_hi[:, 0]).clamp(min=0)) ** 2 + \
               ((self.box_lo[:, 1] - c[1]).clamp(min=0) + (c[1] - self.box_hi[:, 1]).clamp(min=0)) ** 2 < (self.max_ray + 5) ** 2
        lo, hi = self.box_lo[near], self.box_hi[near]
        inv = 1.0 / torch.where(dirs.abs() < 1e-12, torch.full_like(dirs, 1e-12), dirs)
        chunk = 16
        for b in range(0, lo.shape[0], chunk):
            l, h = lo[b:b + chunk], hi[b:b + chunk]
            t0 = (l[None] - origins[:, None]) * inv[:, None]
            t1 = (h[None] - origins[:, None]) * inv[:, None]
            tn = torch.minimum(t0, t1).amax(-1)
            tf = torch.maximum(t0, t1).amin(-1)
            hit = (tn <= tf) & (tn > 0.05)
            tb = torch.where(hit, tn, torch.full_like(tn, big)).amin(-1)
            t = torch.minimum(t, tb)
        return t

    def scan(self, k: int):
        """Scan k in the sensor frame. Returns (points (N,3) float64 numpy, timestamps numpy)."""
        pts, ts = self.scan_torch(k)
        return pts.cpu().numpy(), ts.cpu().numpy()

    def scan_torch(self, k: int):
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed * 1_000_003 + int(k))
        n = self.dirs.shape[0]
        if self.stamps == "column":
            # motion distortion: column c is captured at T(k-1+ (c+1)/cols)
            frac = (np.arange(self.cols) + 1) / self.cols
            Rs = np.empty((self.cols, 3, 3))
            ps = np.empty((self.cols, 3))
            T0, T1 = self.pose(k - 1), self.pose(k)
            D = np.linalg.inv(T0) @ T1
            # constant-velocity interpolation on SE(3) (rotation via so(3) log/exp, translation linear in body frame)
            ang = math.acos(max(-1.0, min(1.0, (np.trace(D[:3, :3]) - 1) / 2)))
            if ang < 1e-12:
                w = np.zeros(3)
            else:
                w = ang / (2 * math.sin(ang)) * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
            for c in range(self.cols):
                Tc = T0.copy()
                Tc[:3, :3] = T0[:3, :3] @ _so3_exp(frac[c] * w)
                Tc[:3, 3] = T0[:3, 3] + T0[:3, :3] @ (frac[c] * D[:3, 3])
                Rs[c], ps[c] = Tc[:3, :3], Tc[:3, 3]
            Rt = torch.tensor(Rs, dtype=torch.float64, device=self.device)[self.col_of_point]
            origins = torch.tensor(ps, dtype=torch.float64, device=self.device)[self.col_of_point]
            dirs_w = torch.einsum("nij,nj->ni", Rt, self.dirs)
        else:
            T = self.pose(k)
            Rt = torch.tensor(T[:3, :3], dtype=torch.float64, device=self.device)
            origins = torch.tensor(T[:3, 3], dtype=torch.float64, device=self.device).expand(n, 3)
            dirs_w = self.dirs @ Rt.T
        rng_ = self._cast(origins, dirs_w)
        noise = torch.randn(n, generator=g, dtype=torch.float64).to(self.device) * self.noise
        valid = rng_ < self.max_ray
        r = (rng_ + noise)[valid]
        pts = (self.dirs[valid] * r[:, None]).to(torch.float32).to(torch.float64)  # fp32-rounded like .bin files
        if self.stamps == "column":
            col = torch.tensor(self.col_of_point, device=self.device)[valid]
            ts = col.to(torch.float64) / self.cols
        else:
            ts = torch.empty(0, dtype=torch.float64, device=self.device)
        return pts, ts


def kitti_shape(seed=0, device="cpu", **kw):
    """BASELINE config 2: 64 beams x 1024 columns, -24.8..+2.0 deg, no stamps."""
    return SyntheticLidar(seed=seed, beams=64, cols=1024, elev_deg=(-24.8, 2.0), stamps="none", device=device, **kw)


def ouster128_shape(seed=0, device="cpu", **kw):
    """BASELINE config 3: 128 beams x 1024 columns, +-22.5 deg, per-column stamps."""
    return SyntheticLidar(seed=seed, beams=128, cols=1024, elev_deg=(-22.5, 22.5), stamps="column", device=device, **kw)


def small_shape(seed=0, beams=16, cols=256, stamps="none", device="cpu", **kw):
    """Reduced scan for CPU-speed tests."""
    return SyntheticLidar(seed=seed, beams=beams, cols=cols, elev_deg=(-24.8, 2.0), stamps=stamps, device=device, **kw)


def surface_cloud(n_raw: int, seed: int = 0, density: float = 14.0, wall_spacing: float = 25.0, wall_height: float = 6.0):
    """Dense map-like cloud for the NN-query sweep (BASELINE config 5): points on a ground sheet and on
    vertical wall sheets at ``density`` points per m^2, so that a 1 m voxel map built from it holds ~8-10
    points per voxel (real KISS-ICP maps: ~8) instead of the ~2 of uniform 3-D noise. Returns (n_raw, 3)."""
    rng = np.random.default_rng(seed)
    n_g = n_raw // 2
    n_w = n_raw - n_g
    side = math.sqrt(n_g / density)
    ground = np.stack([rng.uniform(-side / 2, side / 2, n_g), rng.uniform(-side / 2, side / 2, n_g),
                       rng.normal(0.0, 0.03, n_g)], 1)
    n_planes = max(1, int(side // wall_spacing))
    along = rng.uniform(-side / 2, side / 2, n_w)
    plane = (rng.integers(0, n_planes, n_w) - n_planes / 2 + 0.5) * wall_spacing + rng.normal(0.0, 0.03, n_w)
    z = rng.uniform(0.0, wall_height, n_w)
    flip = rng.random(n_w) < 0.5
    walls = np.stack([np.where(flip, plane, along), np.where(flip, along, plane), z], 1)
    pts = np.concatenate([ground, walls])
    rng.shuffle(pts)
    return pts

"""TEST INFRASTRUCTURE — a SECOND, independent restatement of the reference's step, in scalar numpy-float32
Python, written from the Rust sources (not from oracle/*.cpp):

    models/sfm.rs:48-89   spawn_pedestrians     models/sfm.rs:91-255  update_states
    neighbor_grid.rs:14-36                       util.rs:44-75,92-103   field.rs:235-258

Its only purpose is to cross-check the C++ oracle (tests/test_oracle_cross_check.py): the reference holds
no test for the step and cannot be compiled here, so two restatements written separately agreeing bit for
bit (cell table, order, despawns) and, for positions and velocities, up to the last bit of exp, is the
strongest pin available. Every arithmetic operation is a numpy float32 scalar operation, i.e. one IEEE
rounding per operation and no FMA, like rustc's f32 code. Pure-Python loops: small cases only.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
FMAX = f32(1e12)          # util.rs:45
COS_PHI = f32(-0.17364817766693036)   # sfm.rs:16


def _exp(x):
    """f32::exp: a correctly rounded stand-in (fp64 exp rounded once). glibc's expf, which the Rust code and
    the C++ oracle call, is a 0.502-ulp routine: the two differ in the last bit about once in 10^3 calls."""
    x = float(x)
    if math.isnan(x):
        return f32(np.nan)
    try:
        return f32(math.exp(x))
    except OverflowError:
        return f32(np.inf)


def _as_i32(v) -> int:
    """Rust `f32 as i32`: truncate toward zero, saturate, NaN -> 0."""
    v = float(v)
    if math.isnan(v):
        return 0
    return int(max(-2147483648.0, min(2147483647.0, math.trunc(v))))


def bilinear(grid: np.ndarray, px, py):
    """util.rs:44-58. `grid[(y, x)]`; a missing tap is 1e12."""
    ny, nx = grid.shape
    bx, by = f32(np.floor(px)), f32(np.floor(py))
    tx, ty = f32(px - bx), f32(py - by)
    sx, sy = f32(f32(1.0) - tx), f32(f32(1.0) - ty)
    ix, iy = _as_i32(bx), _as_i32(by)

    def tap(x, y):
        return f32(grid[y, x]) if (0 <= x < nx and 0 <= y < ny) else FMAX

    y = f32(0.0)
    y = f32(y + f32(f32(sy * sx) * tap(ix, iy)))
    y = f32(y + f32(f32(sy * tx) * tap(ix + 1, iy)))
    y = f32(y + f32(f32(ty * sx) * tap(ix, iy + 1)))
    y = f32(y + f32(f32(ty * tx) * tap(ix + 1, iy + 1)))
    return y


def sobel_filter(grid, px, py):
    """util.rs:61-75."""
    one = f32(1.0)
    u00 = bilinear(grid, f32(px - one), f32(py - one))
    u01 = bilinear(grid, px, f32(py - one))
    u02 = bilinear(grid, f32(px + one), f32(py - one))
    u10 = bilinear(grid, f32(px - one), py)
    u12 = bilinear(grid, f32(px + one), py)
    u20 = bilinear(grid, f32(px - one), f32(py + one))
    u21 = bilinear(grid, px, f32(py + one))
    u22 = bilinear(grid, f32(px + one), f32(py + one))
    gx = f32(f32(f32(f32(f32(f32(f32(u00 + u10) + u10) + u20) - u02) - u12) - u12) - u22)
    gy = f32(f32(f32(f32(f32(f32(f32(u00 + u01) + u01) + u02) - u20) - u21) - u21) - u22)
    return gx, gy


def _field_point(x, y, unit):
    """`position / unit - 0.5` (field.rs:236,243,250,256)."""
    return f32(f32(x / unit) - f32(0.5)), f32(f32(y / unit) - f32(0.5))


def _normalize(x, y):
    """glam Vec2::normalize: v * (1.0 / sqrt(x*x + y*y)), no zero guard."""
    with np.errstate(all="ignore"):
        rcp = f32(f32(1.0) / f32(np.sqrt(f32(f32(x * x) + f32(y * y)))))
        return f32(x * rcp), f32(y * rcp)


def distance_from_line(px, py, l0, l1):
    """util.rs:92-103."""
    ax, ay = f32(px - l0[0]), f32(py - l0[1])
    bx, by = f32(l1[0] - l0[0]), f32(l1[1] - l0[1])
    b2 = f32(f32(bx * bx) + f32(by * by))
    if b2 == 0:
        return f32(ax - l0[0]), f32(ay - l0[1])  # (sic) util.rs:98
    t = f32(f32(f32(ax * bx) + f32(ay * by)) / b2)
    t = f32(min(max(t, f32(0.0)), f32(1.0)))
    return f32(ax - f32(t * bx)), f32(ay - f32(t * by))


class PyModel:
    """sfm.rs `SocialForceModel` with the neighbor grid on."""

    def __init__(self, size, neighbor_unit, field_unit, distance_map, potential_maps, obstacles=(), use_distance_map=True):
        self.unit = f32(neighbor_unit)
        self.field_unit = f32(field_unit)
        self.nx = int(np.ceil(f32(f32(size[0]) / self.unit)))   # neighbor_grid.rs:15-16
        self.ny = int(np.ceil(f32(f32(size[1]) / self.unit)))
        self.dist = np.asarray(distance_map, np.float32)
        self.pots = np.asarray(potential_maps, np.float32)
        self.obstacles = [tuple(map(f32, o)) for o in obstacles]
        self.use_distance_map = use_distance_map
        self.p = []   # [x, y, dest, vx, vy, v0]
        self.indices = [0]

    def cell(self, x, y):
        return _as_i32(f32(x / self.unit)), _as_i32(f32(y / self.unit))

    def spawn_pedestrians(self, pos=(), dest=(), v0=()):
        """sfm.rs:48-89."""
        for (x, y), d, s in zip(pos, dest, v0):
            self.p.append([f32(x), f32(y), int(d), f32(0.0), f32(0.0), f32(s)])
        cells = [[] for _ in range(self.nx * self.ny)]
        for i, a in enumerate(self.p):                       # neighbor_grid.rs:22-36
            cx, cy = self.cell(a[0], a[1])
            if cx < 0 or cy < 0 or cx >= self.nx or cy >= self.ny:
                continue
            cells[cy * self.nx + cx].append(i)
        out, self.indices = [], [0]
        for members in cells:                                  # sfm.rs:66-75
            for i in members:
                a = self.p[i]
                qx, qy = _field_point(a[0], a[1], self.field_unit)
                if bilinear(self.pots[a[2]], qx, qy) > f32(0.25):
                    out.append(a)
            self.indices.append(len(out))
        self.p = out

    def _pair(self, a, e, b):
        """sfm.rs:129-155; returns the force on `a` from `b`, or None if beyond the cut-off."""
        dx, dy = f32(a[0] - b[0]), f32(a[1] - b[1])
        d2 = f32(f32(dx * dx) + f32(dy * dy))
        if d2 > f32(4.0):
            return None
        with np.errstate(all="ignore"):
            dist = f32(np.sqrt(d2))
            dirx, diry = _normalize(dx, dy)
            t1x, t1y = f32(dx - f32(b[3] * f32(0.1))), f32(dy - f32(b[4] * f32(0.1)))
            t1len = f32(np.sqrt(f32(f32(t1x * t1x) + f32(t1y * t1y))))
            t2 = f32(dist + t1len)
            vl = f32(f32(np.sqrt(f32(f32(b[3] * b[3]) + f32(b[4] * b[4])))) * f32(0.1))
            bb = f32(f32(np.sqrt(f32(f32(t2 * t2) - f32(vl * vl)))) * f32(0.5))
            sx, sy = f32(dirx + f32(t1x / t1len)), f32(diry + f32(t1y / t1len))
            fb = f32(f32(4.0) * bb)
            nbx, nby = f32(f32(t2 * sx) / fb), f32(f32(t2 * sy) / fb)
            coef = f32(f32(f32(2.1) / f32(0.3)) * _exp(f32(f32(-bb) / f32(0.3))))
            fx, fy = f32(coef * nbx), f32(coef * nby)
            lhs = f32(f32(e[0] * f32(-fx)) + f32(e[1] * f32(-fy)))
            flen = f32(np.sqrt(f32(f32(fx * fx) + f32(fy * fy))))
            if lhs < f32(flen * COS_PHI):
                fx, fy = f32(fx * f32(0.5)), f32(fy * f32(0.5))
        return fx, fy

    def update_states(self):
        """sfm.rs:91-255."""
        acc = []
        for i, a in enumerate(self.p):
            qx, qy = _field_point(a[0], a[1], self.field_unit)
            gx, gy = sobel_filter(self.pots[a[2]], qx, qy)
            e = _normalize(gx, gy)
            with np.errstate(all="ignore"):
                ax = f32(f32(f32(e[0] * a[5]) - a[3]) / f32(0.5))
                ay = f32(f32(f32(e[1] * a[5]) - a[4]) / f32(0.5))
                ax, ay = f32(f32(0.0) + ax), f32(f32(0.0) + ay)
            cx, cy = self.cell(a[0], a[1])
            y0, y1 = max(cy - 1, 0), min(cy + 1, self.ny - 1)
            x0, x1 = max(cx - 1, 0), min(cx + 1, self.nx - 1)
            for y in range(y0, y1 + 1):
                lo, hi = self.indices[y * self.nx + x0], self.indices[y * self.nx + x1 + 1]
                for j in range(lo, hi):
                    if j == i:
                        continue
                    f = self._pair(a, e, self.p[j])
                    if f is not None:
                        ax, ay = f32(ax + f[0]), f32(ay + f[1])
            with np.errstate(all="ignore"):
                if self.use_distance_map:                       # sfm.rs:188-192
                    d = bilinear(self.dist, qx, qy)
                    dgx, dgy = sobel_filter(self.dist, qx, qy)
                    nx_, ny_ = _normalize(dgx, dgy)
                    coef = f32(f32(f32(10.0) * f32(0.2)) * _exp(f32(f32(-d) / f32(0.2))))
                    ax, ay = f32(ax + f32(coef * f32(-nx_))), f32(ay + f32(coef * f32(-ny_)))
                else:                                           # sfm.rs:193-237
                    for (ox0, oy0, ox1, oy1, w) in self.obstacles:
                        ddx, ddy = f32(ox1 - ox0), f32(oy1 - oy0)
                        h = f32(np.sqrt(f32(f32(ddx * ddx) + f32(ddy * ddy))))
                        nnx, nny = ddy, f32(-ddx)
                        rcp = f32(f32(1.0) / f32(np.sqrt(f32(f32(nnx * nnx) + f32(nny * nny)))))
                        if np.isfinite(rcp) and rcp > 0:         # normalize_or_zero
                            nnx, nny = f32(nnx * rcp), f32(nny * rcp)
                        else:
                            nnx, nny = f32(0.0), f32(0.0)
                        nnx, nny = f32(f32(nnx * w) * f32(0.5)), f32(f32(nny * w) * f32(0.5))
                        v0p, v0m = (f32(ox0 + nnx), f32(oy0 + nny)), (f32(ox0 - nnx), f32(oy0 - nny))
                        v1p, v1m = (f32(ox1 + nnx), f32(oy1 + nny)), (f32(ox1 - nnx), f32(oy1 - nny))
                        lines = [(v0p, v0m), (v1p, v1m), (v0p, v1p), (v0m, v1m)]
                        diffs = [distance_from_line(a[0], a[1], l0, l1) for l0, l1 in lines]
                        dists = [f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))) for dx, dy in diffs]
                        if dists[0] < w and dists[1] < w and dists[2] < h and dists[3] < h:   # (sic) sfm.rs:211-216
                            continue
                        k = 0
                        for kk in range(1, 4):                   # min_by keeps the first of equal minima
                            if dists[k] > dists[kk]:
                                k = kk
                        ux, uy = _normalize(*diffs[k])
                        coef = f32(f32(f32(10.0) * f32(0.2)) * _exp(f32(f32(-dists[k]) / f32(0.2))))
                        ax, ay = f32(ax + f32(coef * ux)), f32(ay + f32(coef * uy))
            acc.append((ax, ay))
        for a, (ax, ay) in zip(self.p, acc):                    # sfm.rs:243-254
            with np.errstate(all="ignore"):
                vpx, vpy = a[3], a[4]
                vx, vy = f32(vpx + f32(ax * f32(0.1))), f32(vpy + f32(ay * f32(0.1)))
                vmax = f32(a[5] * f32(1.3))
                l2 = f32(f32(vx * vx) + f32(vy * vy))
                if l2 > f32(vmax * vmax):                        # glam clamp_length_max
                    ln = f32(np.sqrt(l2))
                    vx, vy = f32(vmax * f32(vx / ln)), f32(vmax * f32(vy / ln))
                a[3], a[4] = vx, vy
                a[0] = f32(a[0] + f32(f32(vx + vpx) * f32(0.05)))
                a[1] = f32(a[1] + f32(f32(vy + vpy) * f32(0.05)))

    def state(self):
        n = len(self.p)
        pos = np.array([[a[0], a[1]] for a in self.p], np.float32).reshape(n, 2)
        vel = np.array([[a[3], a[4]] for a in self.p], np.float32).reshape(n, 2)
        return pos, np.array([a[2] for a in self.p], np.uint32), vel, np.array([a[5] for a in self.p], np.float32)

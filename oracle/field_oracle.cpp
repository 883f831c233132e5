// TEST INFRASTRUCTURE — CPU oracle (see vec2.hpp header). Not product code.
//
// field_oracle.cpp: restatement of the reference's one-time field precompute
// (field.rs:16-192,220-232; util.rs:106-111), used to cross-check the product's host field builder.
//
// PARITY PIN STATUS: "parity unpinned".
//  * apply_fmm (field.rs:118-192) is restated line for line; its pop order is fully determined by
//    the heap's total order on (Reverse<NotNan<f32>>, Index{y,x}) so any correct priority queue
//    reproduces it.
//  * Outline rasterisation lives in the crates.io dependency geo-rasterize 0.1.2
//    (Cargo.lock:488-489), which is NOT vendored under /root/reference and cannot be compiled
//    here. Its LineString path burns each segment with the "all touched" line walk published in
//    GDAL (GDALdllImageLineAllTouched); that published algorithm is restated below from its
//    description. No reference test pins it (field.rs:272-324 assert nothing).
//  The hot path never depends on this: the SAME field arrays feed oracle and device.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <tuple>
#include <vector>

#include "vec2.hpp"

namespace oracle {

// util.rs:106-111
static void line_with_width(const Vec2 line[2], float width, Vec2 out[4]) {
    Vec2 a = normalize(line[1] - line[0]);
    Vec2 b = vec2(a.y, -a.x) * 0.5f * width;
    out[0] = line[0] - b;
    out[1] = line[0] + b;
    out[2] = line[1] + b;
    out[3] = line[1] - b;
}

// One segment of an outline, coordinates in pixel units (f32 promoted exactly to f64, as
// geo-rasterize does via `Into<f64>`). Calls burn(iy, ix) for every pixel the segment touches.
template <class Burn>
static void burn_segment_all_touched(double x, double y, double x_end, double y_end, int width, int height,
                                     Burn burn) {
    // Skip segments entirely off the raster.
    if ((y < 0.0 && y_end < 0.0) || (y > height && y_end > height) || (x < 0.0 && x_end < 0.0) ||
        (x > width && x_end > width))
        return;

    // Proceed left to right.
    if (x > x_end) {
        std::swap(x, x_end);
        std::swap(y, y_end);
    }

    // Vertical.
    if (std::floor(x) == std::floor(x_end) || std::fabs(x - x_end) < 0.01) {
        if (y_end < y) std::swap(y, y_end);
        int ix = static_cast<int>(std::floor(x_end));
        int iy = static_cast<int>(std::floor(y));
        int iy_end = static_cast<int>(std::floor(y_end));
        if (ix < 0 || ix >= width) return;
        if (iy < 0) iy = 0;
        if (iy_end >= height) iy_end = height - 1;
        for (; iy <= iy_end; ++iy) burn(iy, ix);
        return;
    }

    // Horizontal.
    if (std::floor(y) == std::floor(y_end) || std::fabs(y - y_end) < 0.01) {
        if (x_end < x) std::swap(x, x_end);
        int ix = static_cast<int>(std::floor(x));
        int iy = static_cast<int>(std::floor(y));
        int ix_end = static_cast<int>(std::floor(x_end));
        if (iy < 0 || iy >= height) return;
        if (ix < 0) ix = 0;
        if (ix_end >= width) ix_end = width - 1;
        for (; ix <= ix_end; ++ix) burn(iy, ix);
        return;
    }

    // General sloped case.
    const double slope = (y_end - y) / (x_end - x);

    if (x_end > width) {
        y_end -= (x_end - width) * slope;
        x_end = width;
    }
    if (x < 0.0) {
        y += (0.0 - x) * slope;
        x = 0.0;
    }
    if (y_end > y) {
        if (y < 0.0) {
            x += (0.0 - y) / slope;
            y = 0.0;
        }
        if (y_end >= height) x_end += (y_end - height) / slope;
    } else {
        if (y >= height) {
            x += (height - y) / slope;
            y = height;
        }
        if (y_end < 0.0) x_end -= (y_end - 0.0) / slope;
    }

    while (x >= 0.0 && x < x_end) {
        const int ix = static_cast<int>(std::floor(x));
        const int iy = static_cast<int>(std::floor(y));
        if (iy >= 0 && iy < height) burn(iy, ix);

        double step_x = std::floor(x + 1.0) - x;
        double step_y = step_x * slope;

        if (static_cast<int>(std::floor(y + step_y)) == iy) {
            x += step_x;
            y += step_y;
        } else if (slope < 0) {
            step_y = iy - y;
            if (step_y > -0.000000001) step_y = -0.000000001;
            step_x = step_y / slope;
            x += step_x;
            y += step_y;
        } else {
            step_y = (iy + 1) - y;
            if (step_y < 0.000000001) step_y = 0.000000001;
            step_x = step_y / slope;
            x += step_x;
            y += step_y;
        }
    }
}

// field.rs:44-61 / 68-85: closed LineString of the width-expanded rectangle, in units of cells.
template <class Burn>
static void burn_outline(const Vec2 line[2], float width, float unit, int fx, int fy, Burn burn) {
    Vec2 v[4];
    line_with_width(line, width, v);
    double px[5], py[5];
    for (int k = 0; k < 4; ++k) {
        Vec2 q = v[k] / unit;
        px[k] = q.x;
        py[k] = q.y;
    }
    px[4] = px[0];  // `shape.close()`
    py[4] = py[0];
    for (int k = 0; k < 4; ++k) burn_segment_all_touched(px[k], py[k], px[k + 1], py[k + 1], fx, fy, burn);
}

// field.rs:118-192
static void apply_fmm(std::vector<float>& potential, const std::vector<float>& f, int ny, int nx) {
    const float F32_MAX = 3.40282347e+38f;
    // BinaryHeap<(Reverse<NotNan<f32>>, Index)> is a max-heap: pops the smallest u first and, among
    // equal u, the greatest Index (derived Ord: y, then x).
    using Entry = std::tuple<float, int, int>;  // (u, y, x)
    auto lower_priority = [](const Entry& a, const Entry& b) {
        if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) > std::get<0>(b);
        if (std::get<1>(a) != std::get<1>(b)) return std::get<1>(a) < std::get<1>(b);
        return std::get<2>(a) < std::get<2>(b);
    };
    std::priority_queue<Entry, std::vector<Entry>, decltype(lower_priority)> queue(lower_priority);
    std::vector<uint8_t> accepted(static_cast<size_t>(ny) * nx, 0);
    auto at = [nx](int x, int y) { return static_cast<size_t>(y) * nx + x; };
    auto inb = [ny, nx](int x, int y) { return x >= 0 && y >= 0 && y < ny && x < nx; };
    const int dj[4] = {-1, 1, 0, 0}, di[4] = {0, 0, -1, 1};  // (j, i) pairs; `ix.add(i, j)` = (x+i, y+j)

    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) {
            if (potential[at(x, y)] == 0.0f) {
                accepted[at(x, y)] = 1;
                for (int k = 0; k < 4; ++k) {
                    int qx = x + di[k], qy = y + dj[k];
                    if (inb(qx, qy) && potential[at(qx, qy)] != 0.0f) {
                        float u = f[at(qx, qy)];
                        potential[at(qx, qy)] = u;
                        queue.emplace(u, qy, qx);
                    }
                }
            }
        }
    }

    while (!queue.empty()) {
        auto [u, y, x] = queue.top();
        queue.pop();
        if (accepted[at(x, y)]) continue;
        accepted[at(x, y)] = 1;

        for (int k = 0; k < 4; ++k) {
            int j = dj[k];
            int qx = x + di[k], qy = y + dj[k];
            if (!inb(qx, qy) || accepted[at(qx, qy)]) continue;

            float fq = f[at(qx, qy)];
            auto get = [&](int gx, int gy) { return inb(gx, gy) ? potential[at(gx, gy)] : F32_MAX; };
            float u1, u2;
            if (j == 0) {
                float u2a = get(qx, qy - 1), u2b = get(qx, qy + 1);
                u1 = u;
                u2 = std::fmin(u2a, u2b);
            } else {
                float u1a = get(qx - 1, qy), u1b = get(qx + 1, qy);
                u1 = std::fmin(u1a, u1b);
                u2 = u;
            }

            float un;
            if (u1 == F32_MAX) {
                un = u2 + fq;
            } else if (u2 == F32_MAX) {
                un = u1 + fq;
            } else {
                float diff = u1 - u2;
                float sq = 2.0f * fq * fq - diff * diff;
                if (sq >= 0.0f) {
                    un = (u1 + u2 + std::sqrt(sq)) / 2.0f;
                } else {
                    un = std::fmin(u1, u2) + fq;
                }
            }

            if (un < potential[at(qx, qy)]) {
                potential[at(qx, qy)] = un;
                queue.emplace(un, qy, qx);
            }
        }
    }
}

}  // namespace oracle

extern "C" {

// field.rs:24-26: shape = ceil(size / unit) as (fy, fx).
void oracle_field_shape(float size_x, float size_y, float unit, int* fy, int* fx) {
    using namespace oracle;
    Vec2 g = ceil(vec2(size_x, size_y) / unit);
    *fy = static_cast<int>(f32_as_usize(g.y));
    *fx = static_cast<int>(f32_as_usize(g.x));
}

// field.rs:220-232 `Field::from_scenario` + FieldBuilder::{new,add_obstacle,add_waypoint,build}.
// obstacles / waypoints: 5 floats each (x0, y0, x1, y1, width).
// Outputs (caller-allocated): obstacle_exist[fy*fx] (u8), distance_map[fy*fx], potential_maps[n_wp*fy*fx].
int oracle_field_build(float size_x, float size_y, float unit, int n_obstacles, const float* obstacles,
                       int n_waypoints, const float* waypoints, uint8_t* obstacle_exist, float* distance_map,
                       float* potential_maps) {
    using namespace oracle;
    int fy, fx;
    oracle_field_shape(size_x, size_y, unit, &fy, &fx);
    if (fy <= 0 || fx <= 0) return -1;
    const size_t cells = static_cast<size_t>(fy) * fx;
    std::vector<uint8_t> obs(cells, 0);

    // field.rs:29-32: outermost ring is obstacle.
    for (int x = 0; x < fx; ++x) obs[x] = obs[static_cast<size_t>(fy - 1) * fx + x] = 1;
    for (int y = 0; y < fy; ++y) obs[static_cast<size_t>(y) * fx] = obs[static_cast<size_t>(y) * fx + fx - 1] = 1;

    for (int k = 0; k < n_obstacles; ++k) {  // field.rs:42-64
        const float* o = obstacles + 5 * k;
        Vec2 line[2] = {vec2(o[0], o[1]), vec2(o[2], o[3])};
        burn_outline(line, o[4], unit, fx, fy, [&](int iy, int ix) { obs[static_cast<size_t>(iy) * fx + ix] = 1; });
    }

    const float F32_MAX = 3.40282347e+38f;
    std::vector<std::vector<float>> maps;
    for (int k = 0; k < n_waypoints; ++k) {  // field.rs:66-88
        const float* w = waypoints + 5 * k;
        Vec2 line[2] = {vec2(w[0], w[1]), vec2(w[2], w[3])};
        std::vector<float> grid(cells, F32_MAX);
        burn_outline(line, w[4], unit, fx, fy, [&](int iy, int ix) { grid[static_cast<size_t>(iy) * fx + ix] = 0.0f; });
        maps.push_back(std::move(grid));
    }

    // field.rs:98-99
    std::vector<float> dist(cells);
    for (size_t c = 0; c < cells; ++c) dist[c] = obs[c] ? 0.0f : 1e24f;
    std::vector<float> funit(cells, unit);
    apply_fmm(dist, funit, fy, fx);

    // field.rs:102-105
    std::vector<float> slowness(cells);
    for (size_t c = 0; c < cells; ++c) slowness[c] = unit * (obs[c] ? 1e6f : 1.0f);
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < n_waypoints; ++k) apply_fmm(maps[k], slowness, fy, fx);

    std::copy(obs.begin(), obs.end(), obstacle_exist);
    std::copy(dist.begin(), dist.end(), distance_map);
    for (int k = 0; k < n_waypoints; ++k) std::copy(maps[k].begin(), maps[k].end(), potential_maps + k * cells);
    return 0;
}

}  // extern "C"

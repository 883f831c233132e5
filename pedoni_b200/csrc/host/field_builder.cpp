// field_builder.cpp — host-side field precompute: the step BEFORE the hot
// path. Produces the arrays the reference's `Field::from_scenario` produces (field.rs:220-232): the
// obstacle mask, the obstacle distance map and one potential map per waypoint, so that the shipped
// scenario TOMLs can be run through libpedoni_cuda.so without the Rust side.
//
// Reference behaviour kept on purpose:
//   - obstacles and waypoints are rasterised as the OUTLINE of their width-expanded rectangle
//     (closed LineString, field.rs:44-53,68-77), not as filled polygons;
//   - the outermost ring of cells is always obstacle (field.rs:29-32);
//   - the marching is the reference's own variant (field.rs:118-192): a neighbour is updated when a cell
//     is accepted, from that cell's value on one axis and the smaller CURRENT value of the two cross-axis
//     neighbours on the other (tentative values included), obstacle cells cost 1e6 * unit
//     (field.rs:102); the result depends on the pop order, which the heap's total order on
//     (value, y, x) fixes.
// The outline walk restates the published "all touched" line rasterisation that geo-rasterize 0.1.2
// (Cargo.lock:488-489, not vendored) implements; nothing in the reference pins it (field.rs:272-324
// assert nothing). One map per waypoint is independent: maps are marched in parallel (rayon over maps at
// field.rs:103 -> OpenMP here).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <vector>

#include "../../../include/pedoni_cuda.h"

namespace {

constexpr float kF32Max = 3.40282347e+38f;

struct P2 {
    float x, y;
};

class FieldBuilder {
  public:
    FieldBuilder(float size_x, float size_y, float unit) : unit_(unit) {
        // field.rs:25-26: shape = ceil(size / unit) as (rows = y, cols = x)
        nx_ = to_count(std::ceil(size_x / unit));
        ny_ = to_count(std::ceil(size_y / unit));
        obstacle_.assign(cells(), 0);
        if (nx_ > 0 && ny_ > 0) {
            for (int x = 0; x < nx_; ++x) obstacle_[x] = obstacle_[at(x, ny_ - 1)] = 1;
            for (int y = 0; y < ny_; ++y) obstacle_[at(0, y)] = obstacle_[at(nx_ - 1, y)] = 1;
        }
    }
    int nx() const { return nx_; }
    int ny() const { return ny_; }
    size_t cells() const { return static_cast<size_t>(nx_) * static_cast<size_t>(ny_); }

    void add_obstacle(const float* o) {  // field.rs:42-64
        outline(o, [this](int x, int y) { obstacle_[at(x, y)] = 1; });
    }
    void add_waypoint(const float* w) {  // field.rs:66-88: LabelBuilder::background(f32::MAX), label 0.0
        std::vector<float> grid(cells(), kF32Max);
        outline(w, [&grid, this](int x, int y) { grid[at(x, y)] = 0.0f; });
        potentials_.push_back(std::move(grid));
    }

    // field.rs:90-114
    void build(uint8_t* obstacle_exist, float* distance_map, float* potential_maps) {
        std::vector<float> distance(cells()), speed_unit(cells(), unit_), slowness(cells());
        for (size_t c = 0; c < cells(); ++c) {
            distance[c] = obstacle_[c] ? 0.0f : 1e24f;
            slowness[c] = unit_ * (obstacle_[c] ? 1e6f : 1.0f);
        }
        march(distance, speed_unit);
        const int n_maps = static_cast<int>(potentials_.size());
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < n_maps; ++k) march(potentials_[k], slowness);
        std::copy(obstacle_.begin(), obstacle_.end(), obstacle_exist);
        std::copy(distance.begin(), distance.end(), distance_map);
        for (int k = 0; k < n_maps; ++k)
            std::copy(potentials_[k].begin(), potentials_[k].end(), potential_maps + static_cast<size_t>(k) * cells());
    }

    // For the device builder (field_device.cu): the rasterised inputs of the marching, nothing marched.
    const std::vector<uint8_t>& obstacle_mask() const { return obstacle_; }
    // cells of a waypoint's outline (potential 0), as flat indices y * nx + x, without a full-size grid
    void waypoint_outline_cells(const float* w, std::vector<uint32_t>& out) const {
        outline(w, [&out, this](int x, int y) { out.push_back(static_cast<uint32_t>(at(x, y))); });
    }

  private:
    float unit_;
    int nx_ = 0, ny_ = 0;
    std::vector<uint8_t> obstacle_;
    std::vector<std::vector<float>> potentials_;

    static int to_count(float v) { return (v != v || v <= 0.0f) ? 0 : (v > 1.0e9f ? 1000000000 : static_cast<int>(v)); }
    size_t at(int x, int y) const { return static_cast<size_t>(y) * nx_ + x; }
    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < nx_ && y < ny_; }

    // util.rs:106-111 `line_with_width`, then `/ unit`, closed (field.rs:44-53)
    template <class Mark>
    void outline(const float* seg, Mark mark) const {
        const P2 p0{seg[0], seg[1]}, p1{seg[2], seg[3]};
        const float width = seg[4];
        P2 a{p1.x - p0.x, p1.y - p0.y};
        const float rcp = 1.0f / std::sqrt((a.x * a.x) + (a.y * a.y));  // glam normalize: v * (1 / length)
        a = P2{a.x * rcp, a.y * rcp};
        const P2 b{a.y * 0.5f * width, -a.x * 0.5f * width};
        const P2 corner[4] = {{p0.x - b.x, p0.y - b.y}, {p0.x + b.x, p0.y + b.y}, {p1.x + b.x, p1.y + b.y},
                              {p1.x - b.x, p1.y - b.y}};
        double px[5], py[5];
        for (int k = 0; k < 4; ++k) {
            px[k] = static_cast<double>(corner[k].x / unit_);  // f32 divide, then widened exactly
            py[k] = static_cast<double>(corner[k].y / unit_);
        }
        px[4] = px[0];
        py[4] = py[0];
        for (int k = 0; k < 4; ++k) touch_segment(px[k], py[k], px[k + 1], py[k + 1], mark);
    }

    // Every cell a segment passes through ("all touched"), clipped to the raster.
    template <class Mark>
    void touch_segment(double x0, double y0, double x1, double y1, Mark mark) const {
        const double w = nx_, h = ny_;
        if ((y0 < 0 && y1 < 0) || (y0 > h && y1 > h) || (x0 < 0 && x1 < 0) || (x0 > w && x1 > w)) return;
        if (x0 > x1) {
            std::swap(x0, x1);
            std::swap(y0, y1);
        }
        if (std::floor(x0) == std::floor(x1) || std::fabs(x0 - x1) < 0.01) {  // one column
            if (y1 < y0) std::swap(y0, y1);
            const int col = static_cast<int>(std::floor(x1));
            if (col < 0 || col >= nx_) return;
            const int first = std::max(static_cast<int>(std::floor(y0)), 0);
            const int last = std::min(static_cast<int>(std::floor(y1)), ny_ - 1);
            for (int row = first; row <= last; ++row) mark(col, row);
            return;
        }
        if (std::floor(y0) == std::floor(y1) || std::fabs(y0 - y1) < 0.01) {  // one row
            const int row = static_cast<int>(std::floor(y0));
            if (row < 0 || row >= ny_) return;
            const int first = std::max(static_cast<int>(std::floor(x0)), 0);
            const int last = std::min(static_cast<int>(std::floor(x1)), nx_ - 1);
            for (int col = first; col <= last; ++col) mark(col, row);
            return;
        }
        const double slope = (y1 - y0) / (x1 - x0);
        if (x1 > w) {
            y1 -= (x1 - w) * slope;
            x1 = w;
        }
        if (x0 < 0) {
            y0 += (0.0 - x0) * slope;
            x0 = 0.0;
        }
        if (y1 > y0) {
            if (y0 < 0) {
                x0 += (0.0 - y0) / slope;
                y0 = 0.0;
            }
            if (y1 >= h) x1 += (y1 - h) / slope;
        } else {
            if (y0 >= h) {
                x0 += (h - y0) / slope;
                y0 = h;
            }
            if (y1 < 0) x1 -= (y1 - 0.0) / slope;
        }
        constexpr double kNudge = 0.000000001;
        while (x0 >= 0 && x0 < x1) {
            const int col = static_cast<int>(std::floor(x0)), row = static_cast<int>(std::floor(y0));
            if (row >= 0 && row < ny_) mark(col, row);
            double sx = std::floor(x0 + 1.0) - x0, sy = sx * slope;
            if (static_cast<int>(std::floor(y0 + sy)) != row) {  // leaves through the top / bottom edge first
                if (slope < 0) {
                    sy = std::min(row - y0, -kNudge);
                } else {
                    sy = std::max((row + 1) - y0, kNudge);
                }
                sx = sy / slope;
            }
            x0 += sx;
            y0 += sy;
        }
    }

    // field.rs:118-192 `apply_fmm`
    void march(std::vector<float>& u, const std::vector<float>& cost) const {
        struct Node {
            float value;
            int y, x;
        };
        // BinaryHeap<(Reverse<NotNan<f32>>, Index)>: smallest value first, ties -> greatest (y, x)
        auto later = [](const Node& a, const Node& b) {
            if (a.value != b.value) return a.value > b.value;
            if (a.y != b.y) return a.y < b.y;
            return a.x < b.x;
        };
        std::priority_queue<Node, std::vector<Node>, decltype(later)> heap(later);
        std::vector<uint8_t> done(cells(), 0);
        static const int step_y[4] = {-1, 1, 0, 0}, step_x[4] = {0, 0, -1, 1};  // field.rs:138,160: (j, i)

        for (int y = 0; y < ny_; ++y)
            for (int x = 0; x < nx_; ++x) {
                if (u[at(x, y)] != 0.0f) continue;
                done[at(x, y)] = 1;
                for (int k = 0; k < 4; ++k) {
                    const int qx = x + step_x[k], qy = y + step_y[k];
                    if (!inside(qx, qy) || u[at(qx, qy)] == 0.0f) continue;
                    u[at(qx, qy)] = cost[at(qx, qy)];
                    heap.push(Node{cost[at(qx, qy)], qy, qx});
                }
            }
        auto value_at = [&](int x, int y) { return inside(x, y) ? u[at(x, y)] : kF32Max; };
        while (!heap.empty()) {
            const Node top = heap.top();
            heap.pop();
            if (done[at(top.x, top.y)]) continue;
            done[at(top.x, top.y)] = 1;
            for (int k = 0; k < 4; ++k) {
                const int qx = top.x + step_x[k], qy = top.y + step_y[k];
                if (!inside(qx, qy) || done[at(qx, qy)]) continue;
                const float f = cost[at(qx, qy)];
                float along_x, along_y;  // (u1, u2) of field.rs:167-175
                if (step_y[k] == 0) {
                    along_x = top.value;
                    along_y = std::fmin(value_at(qx, qy - 1), value_at(qx, qy + 1));
                } else {
                    along_x = std::fmin(value_at(qx - 1, qy), value_at(qx + 1, qy));
                    along_y = top.value;
                }
                float cand;
                if (along_x == kF32Max) {
                    cand = along_y + f;
                } else if (along_y == kF32Max) {
                    cand = along_x + f;
                } else {
                    const float diff = along_x - along_y;
                    const float disc = 2.0f * f * f - diff * diff;
                    cand = disc >= 0.0f ? (along_x + along_y + std::sqrt(disc)) / 2.0f : std::fmin(along_x, along_y) + f;
                }
                if (cand < u[at(qx, qy)]) {
                    u[at(qx, qy)] = cand;
                    heap.push(Node{cand, qy, qx});
                }
            }
        }
    }
};

}  // namespace

namespace pedoni {
// Rasterisation only (field.rs:29-32,42-88): obstacle mask incl. the border ring, and per waypoint the cells of
// its outline. Shared with the device builder so that both march from exactly the same inputs.
int rasterize_scenario(float size_x, float size_y, float unit, int n_obstacles, const float* obstacles, int n_waypoints,
                       const float* waypoints, std::vector<uint8_t>& obstacle_mask,
                       std::vector<std::vector<uint32_t>>& waypoint_cells) {
    FieldBuilder b(size_x, size_y, unit);
    for (int k = 0; k < n_obstacles; ++k) b.add_obstacle(obstacles + 5 * k);
    obstacle_mask = b.obstacle_mask();
    waypoint_cells.clear();
    waypoint_cells.resize(static_cast<size_t>(n_waypoints));
    for (int k = 0; k < n_waypoints; ++k) b.waypoint_outline_cells(waypoints + 5 * k, waypoint_cells[k]);
    return 0;
}
}  // namespace pedoni

extern "C" {

int pedoni_field_shape(float size_x, float size_y, float unit, int32_t* field_ny, int32_t* field_nx) {
    if (!field_ny || !field_nx || !(unit > 0.0f)) return PEDONI_ERR_INVALID;
    FieldBuilder b(0.0f, 0.0f, unit);  // only for the rounding rule
    (void)b;
    const float gx = std::ceil(size_x / unit), gy = std::ceil(size_y / unit);
    if (!(gx >= 1.0f) || !(gy >= 1.0f) || gx > 1.0e6f || gy > 1.0e6f) return PEDONI_ERR_INVALID;
    *field_nx = static_cast<int32_t>(gx);
    *field_ny = static_cast<int32_t>(gy);
    return PEDONI_OK;
}

int pedoni_field_build(float size_x, float size_y, float unit, int32_t n_obstacles, const float* obstacles,
                       int32_t n_waypoints, const float* waypoints, uint8_t* obstacle_exist, float* distance_map,
                       float* potential_maps) {
    int32_t fy = 0, fx = 0;
    if (pedoni_field_shape(size_x, size_y, unit, &fy, &fx) != PEDONI_OK) return PEDONI_ERR_INVALID;
    if (n_obstacles < 0 || n_waypoints < 0 || (n_obstacles > 0 && !obstacles) || (n_waypoints > 0 && !waypoints) ||
        !obstacle_exist || !distance_map || (n_waypoints > 0 && !potential_maps))
        return PEDONI_ERR_INVALID;
    FieldBuilder b(size_x, size_y, unit);
    for (int k = 0; k < n_obstacles; ++k) b.add_obstacle(obstacles + 5 * k);
    for (int k = 0; k < n_waypoints; ++k) b.add_waypoint(waypoints + 5 * k);
    b.build(obstacle_exist, distance_map, potential_maps);
    return PEDONI_OK;
}

}  // extern "C"

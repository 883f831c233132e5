// TEST INFRASTRUCTURE — CPU oracle (see vec2.hpp header). Not product code.
//
// sfm_oracle.hpp: C++17 restatement of the reference's CPU social-force step, in the reference's
// operation order. Every function cites the reference file:line it follows
// (paths relative to /root/reference/pedoni-simulator/src/).
//
// PARITY PIN STATUS: the reference's own tests pin only `bilinear` (util.rs:156-163, 4 values) and
// `distance_from_line` (util.rs:148-154, 2 values); both are checked in tests/test_oracle_golden.py.
// Nothing in the reference pins spawn_pedestrians / update_states / NeighborGrid / FMM, and the
// Rust reference cannot be compiled here (no cargo/rustc, no vendored crates), so for the step
// itself this oracle is "parity unpinned": it was written line against line with sfm.rs.
#pragma once
#include <cstdint>
#include <vector>

#include "vec2.hpp"

namespace oracle {

// util.rs:8-41 `Index` + ndarray bounds rule: negative or >= dim -> None.
struct Grid2 {
    const float* data;
    int ny, nx;  // Array2 shape (rows=y, cols=x), row-major
    bool get(int x, int y, float* out) const {
        if (x < 0 || y < 0 || y >= ny || x >= nx) return false;
        *out = data[static_cast<size_t>(y) * nx + x];
        return true;
    }
};

float bilinear(const Grid2& grid, Vec2 pos);                 // util.rs:44-58
Vec2 sobel_filter(const Grid2& grid, Vec2 pos);              // util.rs:61-75
Vec2 distance_from_line(Vec2 point, Vec2 l0, Vec2 l1);       // util.rs:92-103

// field.rs:194-205 `Field` (hot-path input; arrays are borrowed).
struct Field {
    float unit;
    int fy, fx;
    int n_maps;
    const float* distance_map;    // fy*fx
    const float* potential_maps;  // n_maps*fy*fx
    Grid2 potential(size_t id) const { return {potential_maps + id * static_cast<size_t>(fy) * fx, fy, fx}; }
    Grid2 distance() const { return {distance_map, fy, fx}; }
    float get_potential(size_t waypoint_id, Vec2 position) const;       // field.rs:235-239
    float get_obstacle_distance(Vec2 position) const;                   // field.rs:242-245
    Vec2 get_potential_grad(size_t waypoint_id, Vec2 position) const;   // field.rs:248-252
    Vec2 get_obstacle_distance_grad(Vec2 position) const;               // field.rs:255-258
};

struct Obstacle {  // scenario.rs:23-27
    Vec2 line[2];
    float width;
};

// sfm.rs:26-33 `Pedestrian` SoA.
struct Pedestrians {
    std::vector<Vec2> position;
    std::vector<uint32_t> destination;
    std::vector<Vec2> velocity;
    std::vector<float> desired_speed;
    size_t len() const { return position.size(); }
    void push(Vec2 p, uint32_t d, Vec2 v, float s) {
        position.push_back(p);
        destination.push_back(d);
        velocity.push_back(v);
        desired_speed.push_back(s);
    }
    void reserve(size_t n) {
        position.reserve(n);
        destination.reserve(n);
        velocity.reserve(n);
        desired_speed.reserve(n);
    }
};

// neighbor_grid.rs:8-36
struct NeighborGrid {
    std::vector<std::vector<uint32_t>> data;  // row-major (ny, nx) cells of agent indices
    float unit;
    size_t ny, nx;  // `shape` = (ny, nx)
    NeighborGrid(Vec2 size, float unit);                  // neighbor_grid.rs:14-20
    void update(const std::vector<Vec2>& positions);      // neighbor_grid.rs:22-36
};

// sfm.rs:18-24,35-270 `SocialForceModel`
struct SocialForceModel {
    Pedestrians pedestrians;
    bool has_grid;
    NeighborGrid grid;
    std::vector<uint32_t> neighbor_grid_indices;
    bool use_distance_map;
    std::vector<Obstacle> obstacles;  // scenario.obstacles (borrowed per call in the reference)
    std::vector<Vec2> last_accelerations;  // kept for tests (sfm.rs:93 `accelerations`)

    SocialForceModel(Vec2 field_size, float neighbor_unit, bool use_neighbor_grid, bool use_distance_map,
                     std::vector<Obstacle> obstacles);
    // sfm.rs:48-89. The reference draws desired_speed from fastrand (unseeded, unreproducible);
    // here it is an input so that oracle and device see identical bits.
    void spawn_pedestrians(const Field& field, size_t n, const Vec2* pos, const uint32_t* dest,
                           const float* desired_speed);
    void update_states(const Field& field);  // sfm.rs:91-255
};

}  // namespace oracle

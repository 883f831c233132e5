// TEST INFRASTRUCTURE — CPU oracle (see vec2.hpp / sfm_oracle.hpp headers). Not product code.
// Restates /root/reference/pedoni-simulator/src/{models/sfm.rs,neighbor_grid.rs,util.rs,field.rs}
// in the reference's operation order. Build: -O2 -ffp-contract=off -fno-fast-math [-fopenmp].
#include "sfm_oracle.hpp"

#include <algorithm>
#include <utility>

namespace oracle {

// ---------------------------------------------------------------- util.rs:44-58
float bilinear(const Grid2& grid, Vec2 pos) {
    const float FMAX = 1e12f;

    Vec2 base = floor(pos);
    Vec2 t = pos - base;
    Vec2 s = vec2(1.0f, 1.0f) - t;
    int ix = f32_as_i32(base.x), iy = f32_as_i32(base.y);

    auto tap = [&](int dx, int dy) {
        float v;
        // `ix.add(1, 0)` is plain i32 addition; wrapping cannot occur for finite maps.
        return grid.get(ix + dx, iy + dy, &v) ? v : FMAX;
    };

    float y = 0.0f;
    y += s.y * s.x * tap(0, 0);
    y += s.y * t.x * tap(1, 0);
    y += t.y * s.x * tap(0, 1);
    y += t.y * t.x * tap(1, 1);
    return y;
}

// ---------------------------------------------------------------- util.rs:61-75
Vec2 sobel_filter(const Grid2& grid, Vec2 pos) {
    float u00 = bilinear(grid, pos + vec2(-1.0f, -1.0f));
    float u01 = bilinear(grid, pos + vec2(0.0f, -1.0f));
    float u02 = bilinear(grid, pos + vec2(1.0f, -1.0f));
    float u10 = bilinear(grid, pos + vec2(-1.0f, 0.0f));
    float u12 = bilinear(grid, pos + vec2(1.0f, 0.0f));
    float u20 = bilinear(grid, pos + vec2(-1.0f, 1.0f));
    float u21 = bilinear(grid, pos + vec2(0.0f, 1.0f));
    float u22 = bilinear(grid, pos + vec2(1.0f, 1.0f));

    return vec2(u00 + u10 + u10 + u20 - u02 - u12 - u12 - u22,
                u00 + u01 + u01 + u02 - u20 - u21 - u21 - u22);
}

// ---------------------------------------------------------------- util.rs:92-103
Vec2 distance_from_line(Vec2 point, Vec2 l0, Vec2 l1) {
    Vec2 a = point - l0;
    Vec2 b = l1 - l0;
    float b_len2 = length_squared(b);

    if (b_len2 == 0.0f) {
        return a - l0;  // (sic) util.rs:98
    }
    // f32::max / f32::min: NaN-ignoring, like fmaxf / fminf.
    float t = std::fmin(std::fmax(dot(a, b) / b_len2, 0.0f), 1.0f);
    return a - t * b;
}

// ---------------------------------------------------------------- field.rs:235-258
float Field::get_potential(size_t waypoint_id, Vec2 position) const {
    Vec2 p = position / unit - vec2(0.5f, 0.5f);
    return bilinear(potential(waypoint_id), p);
}
float Field::get_obstacle_distance(Vec2 position) const {
    Vec2 p = position / unit - vec2(0.5f, 0.5f);
    return bilinear(distance(), p);
}
Vec2 Field::get_potential_grad(size_t waypoint_id, Vec2 position) const {
    Vec2 p = position / unit - vec2(0.5f, 0.5f);
    return sobel_filter(potential(waypoint_id), p);
}
Vec2 Field::get_obstacle_distance_grad(Vec2 position) const {
    Vec2 p = position / unit - vec2(0.5f, 0.5f);
    return sobel_filter(distance(), p);
}

// ---------------------------------------------------------------- neighbor_grid.rs:14-20
NeighborGrid::NeighborGrid(Vec2 size, float unit_) : unit(unit_) {
    Vec2 shape = ceil(size / unit_);
    ny = static_cast<size_t>(f32_as_usize(shape.y));
    nx = static_cast<size_t>(f32_as_usize(shape.x));
    data.assign(ny * nx, {});
}

// ---------------------------------------------------------------- neighbor_grid.rs:22-36
void NeighborGrid::update(const std::vector<Vec2>& positions) {
    for (auto& cell : data) cell.clear();  // `fill(ThinVec::new())`

    for (size_t i = 0; i < positions.size(); ++i) {
        IVec2 ix = as_ivec2(positions[i] / unit);
        // util.rs:29-36: negative -> None; then ndarray's own (y, x) bounds check.
        if (ix.x < 0 || ix.y < 0) continue;
        if (static_cast<size_t>(ix.y) >= ny || static_cast<size_t>(ix.x) >= nx) continue;
        data[static_cast<size_t>(ix.y) * nx + ix.x].push_back(static_cast<uint32_t>(i));
    }
}

// ---------------------------------------------------------------- sfm.rs:36-46
SocialForceModel::SocialForceModel(Vec2 field_size, float neighbor_unit, bool use_neighbor_grid,
                                   bool use_distance_map_, std::vector<Obstacle> obstacles_)
    : has_grid(use_neighbor_grid),
      grid(use_neighbor_grid ? field_size : vec2(0.0f, 0.0f), neighbor_unit),
      use_distance_map(use_distance_map_),
      obstacles(std::move(obstacles_)) {}

// ---------------------------------------------------------------- sfm.rs:48-89
void SocialForceModel::spawn_pedestrians(const Field& field, size_t n, const Vec2* pos,
                                         const uint32_t* dest, const float* desired_speed) {
    for (size_t k = 0; k < n; ++k) {
        pedestrians.push(pos[k], dest[k], vec2(0.0f, 0.0f), desired_speed[k]);  // sfm.rs:49-56
    }

    if (has_grid) {
        grid.update(pedestrians.position);  // sfm.rs:59

        Pedestrians sorted;
        sorted.reserve(pedestrians.len());
        neighbor_grid_indices.clear();
        neighbor_grid_indices.reserve(grid.data.size() + 1);
        neighbor_grid_indices.push_back(0);
        size_t index = 0;

        for (const auto& cell : grid.data) {  // row-major: y outer, x inner
            for (size_t j = 0; j < cell.size(); ++j) {
                size_t src = cell[j];
                Vec2 p = pedestrians.position[src];
                uint32_t d = pedestrians.destination[src];
                if (field.get_potential(d, p) > 0.25f) {  // sfm.rs:69 (NaN -> false -> despawn)
                    sorted.push(p, d, pedestrians.velocity[src], pedestrians.desired_speed[src]);
                    index += 1;
                }
            }
            neighbor_grid_indices.push_back(static_cast<uint32_t>(index));
        }

        pedestrians = std::move(sorted);
    } else {
        Pedestrians kept;
        kept.reserve(pedestrians.len());
        for (size_t i = 0; i < pedestrians.len(); ++i) {
            if (field.get_potential(pedestrians.destination[i], pedestrians.position[i]) > 0.25f) {
                kept.push(pedestrians.position[i], pedestrians.destination[i], pedestrians.velocity[i],
                          pedestrians.desired_speed[i]);
            }
        }
        pedestrians = std::move(kept);
    }
}

namespace {

const float COS_PHI = -0.17364817766693036f;  // sfm.rs:16

// sfm.rs:129-155 (identical body at :159-183): force on `id` from pedestrian `i`.
inline void pair_force(Vec2 pos, Vec2 e, Vec2 pos_i, Vec2 vel_i, Vec2& acc) {
    Vec2 difference = pos - pos_i;
    float distance_squared = length_squared(difference);
    if (distance_squared > 4.0f) return;

    float distance = std::sqrt(distance_squared);
    Vec2 direction = normalize(difference);

    Vec2 t1 = difference - vel_i * 0.1f;
    float t1_length = length(t1);
    float t2 = distance + t1_length;
    float vl = length(vel_i) * 0.1f;
    float b = std::sqrt(t2 * t2 - vl * vl) * 0.5f;  // powi(2) == x*x

    Vec2 nabla_b = t2 * (direction + t1 / t1_length) / (4.0f * b);
    Vec2 force = ((2.1f / 0.3f) * std::exp(-b / 0.3f)) * nabla_b;

    if (dot(e, -force) < length(force) * COS_PHI) {
        force = force * 0.5f;
    }

    acc += force;
}

}  // namespace

// ---------------------------------------------------------------- sfm.rs:91-255
void SocialForceModel::update_states(const Field& field) {
    const Pedestrians& peds = pedestrians;
    const long n = static_cast<long>(peds.len());
    std::vector<Vec2>& accelerations = last_accelerations;
    accelerations.assign(peds.len(), vec2(0.0f, 0.0f));

    // rayon `into_par_iter` over agent ids (sfm.rs:93-95)
#pragma omp parallel for schedule(dynamic, 256)
    for (long id = 0; id < n; ++id) {
        Vec2 pos = peds.position[id];
        size_t destination = peds.destination[id];
        Vec2 vel = peds.velocity[id];
        float desired_speed = peds.desired_speed[id];

        Vec2 acc = vec2(0.0f, 0.0f);

        // Force from the destination (sfm.rs:106-109).
        Vec2 grad = field.get_potential_grad(destination, pos);
        Vec2 e = normalize(grad);
        acc += (e * desired_speed - vel) / 0.5f;

        // Force from other pedestrians.
        if (has_grid) {  // sfm.rs:112-156
            IVec2 ix = as_ivec2(pos / grid.unit);
            int32_t shape_x = static_cast<int32_t>(grid.nx), shape_y = static_cast<int32_t>(grid.ny);
            int32_t y_start = std::max(ix.y - 1, 0);
            int32_t y_end = std::min(ix.y + 1, shape_y - 1);
            int32_t x_start = std::max(ix.x - 1, 0);
            int32_t x_end = std::min(ix.x + 1, shape_x - 1);

            for (int32_t y = y_start; y <= y_end; ++y) {
                int32_t offset = y * shape_x;
                size_t i_start = neighbor_grid_indices[static_cast<size_t>(offset + x_start)];
                size_t i_end = neighbor_grid_indices[static_cast<size_t>(offset + x_end + 1)];
                for (size_t i = i_start; i < i_end; ++i) {
                    if (i != static_cast<size_t>(id)) {
                        pair_force(pos, e, peds.position[i], peds.velocity[i], acc);
                    }
                }
            }
        } else {  // sfm.rs:157-185
            for (long i = 0; i < n; ++i) {
                if (i != id) {
                    pair_force(pos, e, peds.position[i], peds.velocity[i], acc);
                }
            }
        }

        // Force from obstacles.
        if (use_distance_map) {  // sfm.rs:188-192
            float distance = field.get_obstacle_distance(pos);
            Vec2 direction = -normalize(field.get_obstacle_distance_grad(pos));
            Vec2 force = ((10.0f * 0.2f) * std::exp(-distance / 0.2f)) * direction;
            acc += force;
        } else {  // sfm.rs:193-237
            for (const Obstacle& obs : obstacles) {
                const Vec2* v = obs.line;
                float w = obs.width;
                Vec2 d = v[1] - v[0];
                float h = length(d);
                Vec2 nn = normalize_or_zero(vec2(d.y, -d.x)) * w * 0.5f;
                Vec2 lines[4][2] = {
                    {v[0] + nn, v[0] - nn},
                    {v[1] + nn, v[1] - nn},
                    {v[0] + nn, v[1] + nn},
                    {v[0] - nn, v[1] - nn},
                };
                Vec2 diffs[4];
                float distances[4];
                for (int k = 0; k < 4; ++k) {
                    diffs[k] = distance_from_line(pos, lines[k][0], lines[k][1]);
                    distances[k] = length(diffs[k]);
                }
                if (distances[0] < w && distances[1] < w && distances[2] < h && distances[3] < h) {
                    continue;  // (sic) sfm.rs:211-216
                }
                // `min_by(partial_cmp)`: first of equal minima (Iterator::min_by keeps the earlier).
                int min_index = 0;
                for (int k = 1; k < 4; ++k) {
                    if (distances[min_index] > distances[k]) min_index = k;
                }
                float min_d = distances[min_index];
                Vec2 direction = normalize(diffs[min_index]);

                Vec2 force = ((10.0f * 0.2f) * std::exp(-min_d / 0.2f)) * direction;
                acc += force;
            }
        }

        accelerations[id] = acc;
    }

    // Serial integration (sfm.rs:243-254).
    for (size_t i = 0; i < pedestrians.len(); ++i) {
        Vec2& pos = pedestrians.position[i];
        Vec2& vel = pedestrians.velocity[i];
        float desired_speed = pedestrians.desired_speed[i];

        Vec2 vel_prev = vel;
        vel += accelerations[i] * 0.1f;
        vel = clamp_length_max(vel, desired_speed * 1.3f);
        pos += (vel + vel_prev) * 0.05f;
    }
}

}  // namespace oracle

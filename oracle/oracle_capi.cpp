// TEST INFRASTRUCTURE — CPU oracle (see vec2.hpp header). Not product code.
// Plain-C entry points over sfm_oracle.{hpp,cpp} so tests/ and bench.py's cpu_baseline leg can
// drive the oracle through ctypes.
#include <chrono>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "sfm_oracle.hpp"

using namespace oracle;

namespace {
struct FieldArgs {
    float unit;
    int fy, fx, n_maps;
    const float* distance_map;
    const float* potential_maps;
};
Field to_field(const FieldArgs* f) { return Field{f->unit, f->fy, f->fx, f->n_maps, f->distance_map, f->potential_maps}; }
}  // namespace

extern "C" {

// ---- util.rs known-answer helpers
float oracle_bilinear(const float* grid, int ny, int nx, float x, float y) {
    return bilinear(Grid2{grid, ny, nx}, vec2(x, y));
}
void oracle_sobel_filter(const float* grid, int ny, int nx, float x, float y, float* out2) {
    Vec2 g = sobel_filter(Grid2{grid, ny, nx}, vec2(x, y));
    out2[0] = g.x;
    out2[1] = g.y;
}
void oracle_distance_from_line(float px, float py, float ax, float ay, float bx, float by, float* out2) {
    Vec2 d = distance_from_line(vec2(px, py), vec2(ax, ay), vec2(bx, by));
    out2[0] = d.x;
    out2[1] = d.y;
}
float oracle_get_potential(const FieldArgs* f, unsigned id, float x, float y) {
    return to_field(f).get_potential(id, vec2(x, y));
}

// ---- model (models/mod.rs:13-25 trait, sfm.rs impl)
void* oracle_model_new(float size_x, float size_y, float neighbor_unit, int use_neighbor_grid, int use_distance_map,
                       int n_obstacles, const float* obstacles /* x0,y0,x1,y1,w */) {
    std::vector<Obstacle> obs(n_obstacles);
    for (int k = 0; k < n_obstacles; ++k) {
        const float* o = obstacles + 5 * k;
        obs[k] = Obstacle{{vec2(o[0], o[1]), vec2(o[2], o[3])}, o[4]};
    }
    return new SocialForceModel(vec2(size_x, size_y), neighbor_unit, use_neighbor_grid != 0, use_distance_map != 0,
                                std::move(obs));
}
void oracle_model_free(void* m) { delete static_cast<SocialForceModel*>(m); }

void oracle_model_grid_shape(void* m, int* ny, int* nx) {
    auto* s = static_cast<SocialForceModel*>(m);
    *ny = static_cast<int>(s->grid.ny);
    *nx = static_cast<int>(s->grid.nx);
}

// PedestrianModel::spawn_pedestrians (append + grid rebuild + despawn + reorder)
void oracle_model_spawn(void* m, const FieldArgs* f, int n, const float* pos_xy, const uint32_t* dest,
                        const float* desired_speed) {
    static_cast<SocialForceModel*>(m)->spawn_pedestrians(to_field(f), static_cast<size_t>(n),
                                                         reinterpret_cast<const Vec2*>(pos_xy), dest, desired_speed);
}
// PedestrianModel::update_states
void oracle_model_update(void* m, const FieldArgs* f) { static_cast<SocialForceModel*>(m)->update_states(to_field(f)); }
// PedestrianModel::get_pedestrian_count
int oracle_model_count(void* m) { return static_cast<int>(static_cast<SocialForceModel*>(m)->pedestrians.len()); }

// list_pedestrians plus the private SoA columns (any pointer may be null).
void oracle_model_get(void* m, float* pos_xy, uint32_t* dest, float* vel_xy, float* desired_speed) {
    auto& p = static_cast<SocialForceModel*>(m)->pedestrians;
    size_t n = p.len();
    if (pos_xy) std::memcpy(pos_xy, p.position.data(), n * sizeof(Vec2));
    if (dest) std::memcpy(dest, p.destination.data(), n * sizeof(uint32_t));
    if (vel_xy) std::memcpy(vel_xy, p.velocity.data(), n * sizeof(Vec2));
    if (desired_speed) std::memcpy(desired_speed, p.desired_speed.data(), n * sizeof(float));
}
// Overwrite the SoA (tests: start both implementations from identical bits, including velocity).
void oracle_model_set(void* m, int n, const float* pos_xy, const uint32_t* dest, const float* vel_xy,
                      const float* desired_speed) {
    auto& p = static_cast<SocialForceModel*>(m)->pedestrians;
    const Vec2* pp = reinterpret_cast<const Vec2*>(pos_xy);
    const Vec2* vv = reinterpret_cast<const Vec2*>(vel_xy);
    p.position.assign(pp, pp + n);
    p.destination.assign(dest, dest + n);
    p.velocity.assign(vv, vv + n);
    p.desired_speed.assign(desired_speed, desired_speed + n);
}
int oracle_model_indices_len(void* m) {
    return static_cast<int>(static_cast<SocialForceModel*>(m)->neighbor_grid_indices.size());
}
void oracle_model_indices(void* m, uint32_t* out) {
    auto& v = static_cast<SocialForceModel*>(m)->neighbor_grid_indices;
    std::memcpy(out, v.data(), v.size() * sizeof(uint32_t));
}
void oracle_model_accelerations(void* m, float* acc_xy) {
    auto& v = static_cast<SocialForceModel*>(m)->last_accelerations;
    std::memcpy(acc_xy, v.data(), v.size() * sizeof(Vec2));
}

// Timed tick loop for the CPU baseline: `steps` x (spawn_pedestrians(no new agents) + update_states),
// the reference's own split (lib.rs:85-91). Returns sum of active counts; times in seconds.
long long oracle_model_run(void* m, const FieldArgs* f, int steps, double* time_spawn, double* time_calc_state) {
    auto* s = static_cast<SocialForceModel*>(m);
    Field field = to_field(f);
    long long updates = 0;
    double ts = 0.0, tc = 0.0;
    for (int k = 0; k < steps; ++k) {
        auto t0 = std::chrono::steady_clock::now();
        s->spawn_pedestrians(field, 0, nullptr, nullptr, nullptr);
        auto t1 = std::chrono::steady_clock::now();
        s->update_states(field);
        auto t2 = std::chrono::steady_clock::now();
        ts += std::chrono::duration<double>(t1 - t0).count();
        tc += std::chrono::duration<double>(t2 - t1).count();
        updates += static_cast<long long>(s->pedestrians.len());
    }
    *time_spawn = ts;
    *time_calc_state = tc;
    return updates;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"

// force.cuh — fused force + integration + next-tick cell key (sm_100a).
//
// Replaces SocialForceModel::update_states (sfm.rs:91-255): steering toward the destination
// (sfm.rs:106-109), pair repulsion over the 3x3 neighbor-cell block (sfm.rs:112-156), wall repulsion
// from the distance map (sfm.rs:188-192) or from obstacle segments (sfm.rs:193-237), and explicit
// integration (sfm.rs:243-254). The reference computes all accelerations first and integrates in a
// second serial loop; here the state is double-buffered (read `in`, write `out`) so that agent i
// sees every neighbour's PRE-integration position/velocity and one pass over the 24-byte state does
// both. The epilogue also evaluates the next rebuild's cell key and despawn predicate
// (neighbor_grid.rs:27-33, sfm.rs:69) on the just-integrated position, which is bit-identical to
// evaluating it at the start of the next tick on the stored value.
#pragma once
#include "grid_sort.cuh"

namespace pedoni {

constexpr float kCosPhi = -0.17364817766693036f;  // sfm.rs:16
constexpr int kEdgeFloats = 24;                   // per obstacle: 4 x (l0.x, l0.y, b.x, b.y, |b|^2), w, h, pad

struct ForceParams {
    AgentArrays in;              // cell-sorted state (pre-integration)
    AgentArrays out;             // integrated state, same indexing
    const uint32_t* d_range;     // device [begin, end): agents this launch integrates
    const uint32_t* d_owned;     // device [begin, end): agents this handle owns (the updates counter counts these)
    uint32_t count_upper;        // host upper bound of end - begin (grid size)
    const uint32_t* cell_start;  // local cell table (neighbor_grid_indices, sfm.rs:22)
    GridView grid;
    FieldView field;
    uint32_t* keys_out;          // next rebuild's keys, indexed like the arrays
    uint32_t* error_flag;
    unsigned long long* updates_total;  // += owned agents of this launch (thread 0 of block 0)
    const float* obstacle_edges;  // segment-wall variant only
    int n_obstacles;
};

// sfm.rs:129-155. `self` is the agent being updated, `o` the other pedestrian.
template <Math M>
__device__ __forceinline__ void pair_force(float2 pos, float2 e, float2 pos_o, float2 vel_o, float2& acc) {
    using O = Ops<M>;
    const float dx = O::sub(pos.x, pos_o.x), dy = O::sub(pos.y, pos_o.y);
    const float d2 = O::add(O::mul(dx, dx), O::mul(dy, dy));
    if (d2 > 4.0f) return;  // sfm.rs:133-135

    const float dist = O::sqrt(d2);
    const float rinv = O::rcp(dist);  // glam normalize = v * (1 / length)
    const float dirx = O::mul(dx, rinv), diry = O::mul(dy, rinv);

    const float t1x = O::sub(dx, O::mul(vel_o.x, 0.1f)), t1y = O::sub(dy, O::mul(vel_o.y, 0.1f));
    const float t1len = O::sqrt(O::add(O::mul(t1x, t1x), O::mul(t1y, t1y)));
    const float t2 = O::add(dist, t1len);
    const float vl = O::mul(O::sqrt(O::add(O::mul(vel_o.x, vel_o.x), O::mul(vel_o.y, vel_o.y))), 0.1f);
    const float b = O::mul(O::sqrt(O::sub(O::mul(t2, t2), O::mul(vl, vl))), 0.5f);

    // nabla_b = t2 * (direction + t1 / t1_length) / (4.0 * b)
    const float sx = O::add(dirx, O::div(t1x, t1len)), sy = O::add(diry, O::div(t1y, t1len));
    const float fb = O::mul(4.0f, b);
    const float nbx = O::div(O::mul(t2, sx), fb), nby = O::div(O::mul(t2, sy), fb);
    // force = 2.1 / 0.3 * exp(-b / 0.3) * nabla_b   (f32 constant 2.1/0.3 = 6.9999995)
    const float coef = O::mul(2.1f / 0.3f, O::exp(O::div(-b, 0.3f)));
    float fx = O::mul(coef, nbx), fy = O::mul(coef, nby);

    // anisotropy, sfm.rs:150-152
    const float lhs = O::add(O::mul(e.x, -fx), O::mul(e.y, -fy));
    const float flen = O::sqrt(O::add(O::mul(fx, fx), O::mul(fy, fy)));
    if (lhs < O::mul(flen, kCosPhi)) {
        fx = O::mul(fx, 0.5f);
        fy = O::mul(fy, 0.5f);
    }
    acc.x = O::add(acc.x, fx);
    acc.y = O::add(acc.y, fy);
}

// util.rs:92-103 with b = l1 - l0 and |b|^2 precomputed on the host (same f32 ops, same values).
template <Math M>
__device__ __forceinline__ float2 distance_from_edge(float2 p, const float* __restrict__ e) {
    using O = Ops<M>;
    const float ax = O::sub(p.x, e[0]), ay = O::sub(p.y, e[1]);
    const float bx = e[2], by = e[3], b_len2 = e[4];
    if (b_len2 == 0.0f) return make_float2(O::sub(ax, e[0]), O::sub(ay, e[1]));  // (sic) util.rs:98
    const float t = fminf(fmaxf(O::div(O::add(O::mul(ax, bx), O::mul(ay, by)), b_len2), 0.0f), 1.0f);
    return make_float2(O::sub(ax, O::mul(t, bx)), O::sub(ay, O::mul(t, by)));
}

template <Math M, bool kDistanceMap>
__global__ void __launch_bounds__(128) force_integrate_kernel(ForceParams p) {
    using O = Ops<M>;
    extern __shared__ float s_edges[];

    const uint32_t begin = p.d_range[0], end = p.d_range[1];
    const uint32_t id = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t lo = max(begin, p.d_owned[0]), hi = min(end, p.d_owned[1]);
        if (hi > lo) atomicAdd(p.updates_total, static_cast<unsigned long long>(hi - lo));
    }
    const bool live = id < end;
    if (kDistanceMap && !live) return;  // the segment variant needs every thread at its barriers

    float2 pos = make_float2(0.f, 0.f), vel = pos, e = pos, acc = pos;
    float v0 = 0.f;
    uint32_t dest = 0;
    int row = 0;
    if (live) {
        pos = p.in.pos[id];
        vel = p.in.vel[id];
        v0 = p.in.v0[id];
        dest = p.in.dest[id];

        // ---- steering (sfm.rs:106-109; field.rs:248-252)
        const float2 q = field_coord(pos, p.field.unit);
        {
            float gx, gy, unused;
            // dest < n_maps is guaranteed by the rebuild that admitted this agent (sort_key).
            sobel_sample<false>(p.field.potential_maps + static_cast<size_t>(dest) * p.field.fy * p.field.fx,
                                p.field.fy, p.field.fx, q, gx, gy, unused);
            const float rlen = O::rcp(O::sqrt(O::add(O::mul(gx, gx), O::mul(gy, gy))));
            e = make_float2(O::mul(gx, rlen), O::mul(gy, rlen));
            acc.x = O::add(acc.x, O::div(O::sub(O::mul(e.x, v0), vel.x), 0.5f));
            acc.y = O::add(acc.y, O::div(O::sub(O::mul(e.y, v0), vel.y), 0.5f));
        }

        // ---- pair repulsion (sfm.rs:112-156)
        {
            const int2 c = cell_of(pos, p.grid.unit);
            row = c.y;
            // Reference clamps to the grid; the local table may start at row_base (slabs) and always
            // holds every row a live agent can reach (its own rows plus one halo row each side).
            const int ly = c.y - p.grid.row_base;
            const int y_start = max(ly - 1, 0), y_end = min(ly + 1, p.grid.table_rows - 1);
            const int x_start = min(max(c.x - 1, 0), p.grid.nx - 1), x_end = max(min(c.x + 1, p.grid.nx - 1), 0);
            for (int y = y_start; y <= y_end; ++y) {
                const uint32_t* row = p.cell_start + static_cast<size_t>(y) * p.grid.nx;
                const uint32_t i_start = __ldg(row + x_start), i_end = __ldg(row + x_end + 1);
                for (uint32_t i = i_start; i < i_end; ++i) {
                    if (i != id) pair_force<M>(pos, e, __ldg(p.in.pos + i), __ldg(p.in.vel + i), acc);
                }
            }
        }

        // ---- walls, distance-map variant (sfm.rs:188-192; field.rs:242-245,255-258)
        if (kDistanceMap) {
            float gx, gy, distance;
            sobel_sample<true>(p.field.distance_map, p.field.fy, p.field.fx, q, gx, gy, distance);
            const float rlen = O::rcp(O::sqrt(O::add(O::mul(gx, gx), O::mul(gy, gy))));
            const float coef = O::mul(10.0f * 0.2f, O::exp(O::div(-distance, 0.2f)));
            acc.x = O::add(acc.x, O::mul(coef, -O::mul(gx, rlen)));
            acc.y = O::add(acc.y, O::mul(coef, -O::mul(gy, rlen)));
        }
    }

    // ---- walls, segment variant (sfm.rs:193-237): obstacles staged through shared memory
    if (!kDistanceMap) {
        constexpr int kChunk = 64;  // obstacles per stage: 64 * 24 * 4 B = 6 KB
        for (int o0 = 0; o0 < p.n_obstacles; o0 += kChunk) {
            const int n = min(kChunk, p.n_obstacles - o0);
            __syncthreads();
            for (int k = threadIdx.x; k < n * kEdgeFloats; k += blockDim.x)
                s_edges[k] = __ldg(p.obstacle_edges + static_cast<size_t>(o0) * kEdgeFloats + k);
            __syncthreads();
            if (!live) continue;
            for (int o = 0; o < n; ++o) {
                const float* ob = s_edges + o * kEdgeFloats;
                const float w = ob[20], h = ob[21];
                float2 diffs[4];
                float dists[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    diffs[k] = distance_from_edge<M>(pos, ob + 5 * k);
                    dists[k] = O::sqrt(O::add(O::mul(diffs[k].x, diffs[k].x), O::mul(diffs[k].y, diffs[k].y)));
                }
                if (dists[0] < w && dists[1] < w && dists[2] < h && dists[3] < h) continue;  // (sic) sfm.rs:211-216
                int mi = 0;  // first of equal minima (Iterator::min_by)
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (dists[mi] > dists[k]) mi = k;
                float min_d = dists[0];
                float2 md = diffs[0];
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (mi == k) {
                        min_d = dists[k];
                        md = diffs[k];
                    }
                const float rlen = O::rcp(min_d);
                const float coef = O::mul(10.0f * 0.2f, O::exp(O::div(-min_d, 0.2f)));
                acc.x = O::add(acc.x, O::mul(coef, O::mul(md.x, rlen)));
                acc.y = O::add(acc.y, O::mul(coef, O::mul(md.y, rlen)));
            }
        }
        if (!live) return;
    }

    // ---- integration (sfm.rs:243-254), dt = 0.1
    float2 vn = make_float2(O::add(vel.x, O::mul(acc.x, 0.1f)), O::add(vel.y, O::mul(acc.y, 0.1f)));
    {
        const float vmax = O::mul(v0, 1.3f);
        const float len2 = O::add(O::mul(vn.x, vn.x), O::mul(vn.y, vn.y));
        if (len2 > O::mul(vmax, vmax)) {  // glam clamp_length_max: max * (v / sqrt(len2))
            const float len = O::sqrt(len2);
            vn = make_float2(O::mul(vmax, O::div(vn.x, len)), O::mul(vmax, O::div(vn.y, len)));
        }
    }
    const float2 pn = make_float2(O::add(pos.x, O::mul(O::add(vn.x, vel.x), 0.05f)),
                                  O::add(pos.y, O::mul(O::add(vn.y, vel.y), 0.05f)));

    p.out.pos[id] = pn;
    p.out.vel[id] = vn;
    p.out.v0[id] = v0;
    p.out.dest[id] = dest;
    p.keys_out[id] = sort_key(p.grid, p.field, pn, dest, p.error_flag);
    // Slab handles exchange two ghost rows per tick, which covers every move of less than one grid row
    // (1.4 m per 0.1 s); anything faster would silently vanish at a slab boundary, so flag it.
    if (p.grid.slab) {
        const int new_row = __float2int_rz(S::div(pn.y, p.grid.unit));
        if (abs(new_row - row) >= 2 && pn.y == pn.y) atomicOr(p.error_flag, kErrRowJump);
    }
}

}  // namespace pedoni

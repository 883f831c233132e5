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
//
// Kernel shape (one WARP = 32 consecutive agents of the cell-sorted arrays, one thread per agent;
// warps are autonomous: no block-wide barrier on the distance-map path):
//   1. every lane loads its agent and the six cell-table entries bounding its three row ranges
//      (rows cy-1, cy, cy+1; sfm.rs:117-127). Ranges are monotone in the agent index, so the warp's
//      union per row offset is ONE contiguous index window [first lane's start, last lane's end)
//      — also across a row end, because consecutive rows are adjacent in the sorted arrays. The bounds
//      travel by warp shuffle.
//   2. the three windows (position and velocity) are staged in the warp's slice of shared memory by six
//      bulk copies (cp.async.bulk / TMA, completion on a per-warp mbarrier: no registers, no LSU work, no
//      waiting) instead of ~36 gathers per agent, while the lanes fetch and evaluate the steering / wall
//      field samples — 4x4 texel footprints, four texture gathers each in fast math.
//   3. SCAN: each lane walks its candidates in the tile, two per iteration (LDS.64 and the 2 m cut-off
//      test, branch free, PTX) and appends the in-range ones to a private list in shared memory, in index
//      order.
//   4. FORCE: each lane evaluates the Helbing-Molnar term for its list, in the same order as the
//      reference sums it. Splitting scan from force keeps the expensive body converged: a warp runs
//      max-over-lanes(in-range) ~ 18 heavy iterations instead of sum-over-rows max-over-lanes
//      (candidates) ~ 29 with three fifths of the lanes idle (profiles/r01a_force_integrate_full.md:
//      18.5 of 32 lanes active).
//   5. integration, key of the new position, coalesced stores.
// Warps whose windows exceed their tile (a cell holding ~100 agents) read candidates from global
// memory instead; that is a correctness path, not a fast path.
#pragma once
#include "grid_sort.cuh"

namespace pedoni {

constexpr float kCosPhi = -0.17364817766693036f;  // sfm.rs:16
constexpr int kEdgeFloats = 24;                   // per obstacle: 4 x (l0.x, l0.y, b.x, b.y, |b|^2), w, h, pad

// Tunables (overridable at build time for sweeps: scripts/sweep_force.py).
#ifndef PEDONI_FORCE_THREADS
#define PEDONI_FORCE_THREADS 128
#endif
#ifndef PEDONI_FORCE_MIN_BLOCKS
#define PEDONI_FORCE_MIN_BLOCKS 9
#endif
#ifndef PEDONI_FORCE_MIN_BLOCKS_STRICT
#define PEDONI_FORCE_MIN_BLOCKS_STRICT 8  // the IEEE path (fp64 exp, divides) needs the registers more than the warps
#endif
#ifndef PEDONI_FORCE_UNROLL
#define PEDONI_FORCE_UNROLL 2  // neighbours in flight in the FORCE loop
#endif
#ifndef PEDONI_TILE_ENTRIES
#define PEDONI_TILE_ENTRIES 192
#endif
#ifndef PEDONI_LIST_DEPTH
#define PEDONI_LIST_DEPTH 32
#endif
#ifndef PEDONI_FAST_FIELD
#define PEDONI_FAST_FIELD 1  // 0: PEDONI_MATH_FAST keeps the reference-order field gradient (experiments)
#endif
#ifndef PEDONI_FAST_PAIR
#define PEDONI_FAST_PAIR 1   // 0: PEDONI_MATH_FAST keeps the reference-order pair term (experiments)
#endif
#ifndef PEDONI_FAR_LOOKUP_EARLY
#define PEDONI_FAR_LOOKUP_EARLY 1  // ask the far-from-walls mask before the potential map's gathers (B200, 10 M: 0.615 vs 0.623 ms after them)
#endif
#ifndef PEDONI_WALL_EARLY_ADD
#define PEDONI_WALL_EARLY_ADD 1    // fast math: add the wall term to the acceleration before the pair loops (0.615 vs 0.619 ms)
#endif
#ifndef PEDONI_PREFETCH_AHEAD
#define PEDONI_PREFETCH_AHEAD 0  // pedestrians ahead whose state a warp pulls into L2 (0 = off; experiments)
#endif
constexpr int kForceThreads = PEDONI_FORCE_THREADS;
constexpr int kForceUnroll = PEDONI_FORCE_UNROLL;
constexpr int kForceWarps = kForceThreads / 32;
constexpr int kTileEntries = PEDONI_TILE_ENTRIES;  // agents per WARP tile: 3 windows of ~(32 + 2 cells) agents (~108 at 1 ped/m^2)
constexpr int kListDepth = PEDONI_LIST_DEPTH;     // in-range neighbours per agent per round (mean 12 at 1 ped/m^2)
#ifndef PEDONI_BULK_STAGE
#define PEDONI_BULK_STAGE 1  // stage the tile with cp.async.bulk (TMA) instead of per-lane cp.async
#endif
constexpr int kTileAlloc = kTileEntries + 2;  // +1 spare slot for the scan's read-ahead, +1 keeps 16-byte alignment
// two tiles (pos, vel) + the neighbour lists (+1 row: see pair_forces_tiled) + the warp's mbarrier
constexpr size_t kWarpMbarOffset = 2 * sizeof(float2) * kTileAlloc + sizeof(uint16_t) * (kListDepth + 1) * 32;
constexpr size_t kWarpSmemBytes = kWarpMbarOffset + 16;
static_assert(kTileAlloc % 2 == 0 && kWarpSmemBytes % 16 == 0, "bulk copies need 16-byte aligned tiles");
constexpr size_t kForceSmemBytes = kWarpSmemBytes * kForceWarps;
static_assert(kWarpSmemBytes * (kForceThreads / 32) + 2048 <= 65536, "list entries are 16-bit addresses in the CTA's shared window");
static_assert(kWarpSmemBytes % 16 == 0, "warp slices stay 16-byte aligned");

struct ForceParams {
    AgentArrays in;              // cell-sorted state (pre-integration)
    AgentArrays out;             // integrated state, same indexing
    const uint32_t* d_range;     // device [begin, end): agents this launch integrates (CTAs with blockIdx.y == 0)
    const uint32_t* d_range_hi;  // a second range for the CTAs with blockIdx.y == 1 (both edges of a slab in one launch)
    const uint32_t* d_owned;     // device [begin, end): agents this handle owns (the updates counter counts these)
    uint32_t count_upper;        // host upper bound of end - begin (grid size)
    uint32_t cap;                // elements allocated per array (bounds checks of debug builds)
    uint32_t table_cells;        // entries of cell_start - 1
    const uint32_t* cell_start;  // local cell table (neighbor_grid_indices, sfm.rs:22)
    GridView grid;
    FieldView field;
    CellSort cs;                 // the next rebuild's cell membership (the enrolment is fused here)
    uint32_t* error_flag;
    unsigned long long* updates_total;  // += owned agents of this launch (thread 0 of block 0)
    unsigned long long* arrived;        // [16] cumulative arrivals by destination (owned agents only)
    const float* obstacle_edges;  // segment-wall variant only
    int n_obstacles;
};

// ---- pair term, sfm.rs:129-155 ----------------------------------------------------------------------
// (po, vo) = position and velocity of the other pedestrian; the caller has already applied the cut-off
// `|d|^2 > 4 -> skip` (sfm.rs:133-135) and the self test (sfm.rs:130).
template <Math M>
struct PairTerm;

// Reference operation order, IEEE ops, no contraction.
template <>
struct PairTerm<Math::Strict> {
    static __device__ __forceinline__ void add(float2 pos, float2 e, float2 po, float2 vo, float2& acc) {
        using O = Ops<Math::Strict>;
        const float dx = O::sub(pos.x, po.x), dy = O::sub(pos.y, po.y);
        const float d2 = O::add(O::mul(dx, dx), O::mul(dy, dy));
        const float dist = O::sqrt(d2);
        const float rinv = O::rcp(dist);  // glam normalize = v * (1 / length)
        const float dirx = O::mul(dx, rinv), diry = O::mul(dy, rinv);

        const float t1x = O::sub(dx, O::mul(vo.x, 0.1f)), t1y = O::sub(dy, O::mul(vo.y, 0.1f));
        const float t1len = O::sqrt(O::add(O::mul(t1x, t1x), O::mul(t1y, t1y)));
        const float t2 = O::add(dist, t1len);
        const float vl = O::mul(O::sqrt(O::add(O::mul(vo.x, vo.x), O::mul(vo.y, vo.y))), 0.1f);
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
};

// Same term with the algebra folded for the SFU/FMA pipes: 3 rsqrt + 1 ex2 and ~35 FP32 ops.
//   2b = sqrt(q), q = t2^2 - |0.1 v_i|^2;  force = s * n with n = d/|d| + t1/|t1| and the positive scalar
//   s = (2.1/0.3) exp(-b/0.3) t2 / (4b) = 3.5 * exp2(-(log2 e / 0.6) sqrt(q)) * t2 / sqrt(q).
//   Anisotropy test e.(-f) < |f| cos(phi)  <=>  e.n > -cos(phi) |n|  <=>  e.n > 0 and (e.n)^2 > cos^2(phi) |n|^2
//   (cos(phi) < 0), which needs no square root of |f|.
template <>
struct PairTerm<Math::Fast> {
    static __device__ __forceinline__ float rsqrt(float a) {
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float ex2(float a) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ void add(float2 pos, float2 e, float2 po, float2 vo, float2& acc) {
        // Vec2 operations as packed fp32x2 (FFMA2): one issue slot for both components.
        const f32x2 d2v = sub2(pack2(pos), pack2(po));
        const float2 d = unpack2(d2v);
        const float d2 = fmaf(d.y, d.y, d.x * d.x);
        const float rinv = rsqrt(d2);
        const f32x2 t1v = fma2(pack2(vo), pack2(-0.1f, -0.1f), d2v);  // d - 0.1 v_i
        const float2 t1 = unpack2(t1v);
        const float vv = fmaf(vo.y, vo.y, vo.x * vo.x);
        const float t1sq = fmaf(t1.y, t1.y, t1.x * t1.x);
        const float rt1 = rsqrt(t1sq);
        const float t2 = fmaf(d2, rinv, t1sq * rt1);  // |d| + |t1|
        const float q = fmaf(vv, -0.01f, t2 * t2);    // t2^2 - |0.1 v_i|^2
        const float rq = rsqrt(q);
        constexpr float kC = -1.4426950408889634f / 0.6f;  // exp(-b/0.3), b = sqrt(q)/2, as exp2
        constexpr float kLog2K = 1.8073549220576042f;      // log2(2.1 / 0.3 / 2)
        float s = ex2(fmaf(q * rq, kC, kLog2K)) * (t2 * rq);
        const f32x2 nv = fma2(t1v, pack2(rt1, rt1), mul2(d2v, pack2(rinv, rinv)));  // d/|d| + t1/|t1|
        const float2 n = unpack2(nv);
        const float en = fmaf(e.y, n.y, e.x * n.x);
        const float n2 = fmaf(n.y, n.y, n.x * n.x);
        constexpr float kCos2 = kCosPhi * kCosPhi;
        if (en > 0.0f && en * en > kCos2 * n2) s *= 0.5f;
        acc = unpack2(fma2(pack2(s, s), nv, pack2(acc)));
    }
};

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

// 1 / |(x, y)| as glam's normalize computes it (1.0 / sqrt(x*x + y*y)); one rsqrt in fast mode.
template <Math M>
__device__ __forceinline__ float inv_length(float x, float y) {
    if (M == Math::Fast) return PairTerm<Math::Fast>::rsqrt(fmaf(y, y, x * x));
    using O = Ops<M>;
    return O::rcp(O::sqrt(O::add(O::mul(x, x), O::mul(y, y))));
}

// Field gradient for the force terms. Strict: the reference's 8 (9) bilinear samples, bit for bit.
// Fast: the same Sobel-of-bilinear evaluated separably on the 4x4 texel footprint with FMAs (~50 ops
// instead of ~130). Three cases keep the reference's operation order even in fast mode:
//   - map borders (out-of-bounds taps read 1e12, util.rs:45);
//   - footprints holding through-wall values (>= 1e5 * unit; obstacle cells cost 1e6 * unit,
//     field.rs:102): there the reference's sums of ~2.5e5-sized samples cancel catastrophically and
//     the rounding noise IS the behaviour — pedestrians in sealed pockets random-walk on it. The
//     separable form is more accurate there, so it is only used where that noise is below 0.1 % of a texel;
//   - (near-)vanishing gradients, where the reference yields either noise or an exact zero (-> NaN ->
//     the pedestrian disappears). Without this guard evacuation.toml, whose spawn lines lie exactly on
//     corridor mid-lines, kept 13 pedestrians the reference loses in the first tick (40 %-evacuation
//     time 31.6 +- 1.8 s instead of 22.6 +- 1.2 s over 20 seeds; scripts/exp_fast_stats.py).
//
// kTex: the 4x4 footprint comes from four texture gathers instead of sixteen loads. The kernel is bound by
// the L1 data pipe (92 % busy: ncu, profiles/r01g_force_10M_full.md), and sixteen 4-byte loads with one
// address per lane cost ~8 wavefronts each; a gather returns a 2x2 block per lane and request. Same texel
// values, so everything downstream is unchanged.
//
// tld4 returns the 2x2 block a bilinear fetch at the given coordinate would blend, as (x, y, z, w) =
// texels (i, j+1), (i+1, j+1), (i+1, j), (i, j). With unnormalised coordinates texel i covers [i, i+1),
// so the block whose lower texel is (i, j) is addressed at (i + 1, j + 1): half a texel away from any
// switch-over, exact in fp32. All maps live in one atlas so that the whole warp uses ONE texture handle: with
// a handle per destination the compiler wraps every tld4 in a loop over the distinct handles of the warp.
// pedoni_create verifies this layout against plain loads on the actual maps before it enables the path
// (footprint_check_kernel).
__device__ __forceinline__ void footprint_gather(cudaTextureObject_t tex, int x0, int y0, float (&t)[4][4]) {
    const float fx = static_cast<float>(x0) + 1.0f, fy = static_cast<float>(y0) + 1.0f;
#pragma unroll
    for (int by = 0; by < 2; ++by)
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) {
            const float4 g = tex2Dgather<float4>(tex, fx + 2.0f * bx, fy + 2.0f * by, 0);
            t[2 * by + 1][2 * bx] = g.x;
            t[2 * by + 1][2 * bx + 1] = g.y;
            t[2 * by][2 * bx + 1] = g.z;
            t[2 * by][2 * bx] = g.w;
        }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {  // sm_100: one FMNMX3; NaN operands are skipped
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// `tile`: texel offset of the map inside the atlas (kTex only).
template <Math M, bool WithCentre, bool kTex>
__device__ __forceinline__ void field_gradient(const float* __restrict__ g, cudaTextureObject_t tex, int2 tile, int ny, int nx,
                                               float2 q, float noise_limit, float flat_limit2, float& gx, float& gy,
                                               float& centre) {
    if (M == Math::Fast && PEDONI_FAST_FIELD) {
        const float bx = floorf(q.x), by = floorf(q.y);
        const int x0 = __float2int_rz(bx) - 1, y0 = __float2int_rz(by) - 1;
        if (x0 >= 0 && y0 >= 0 && x0 + 3 < nx && y0 + 3 < ny) {
            const float tx = q.x - bx, ty = q.y - by, sx = 1.0f - tx, sy = 1.0f - ty;
            float t[4][4];
            if (kTex) {
                footprint_gather(tex, x0 + tile.x, y0 + tile.y, t);
            } else {
                const float* base = g + static_cast<size_t>(y0) * nx + x0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) t[r][c] = __ldg(base + static_cast<size_t>(r) * nx + c);
            }
            // (The max over all 16 texels also keeps the 16 loads in flight together: testing a single
            // central texel is 15 instructions shorter and measurably SLOWER, 0.846 vs 0.809 ms at 10 M.)
            float big = fmax3(t[0][0], t[0][1], t[0][2]);  // FMNMX3: 8 instructions for the 16 texels
            big = fmax3(big, t[0][3], t[1][0]);
            big = fmax3(big, t[1][1], t[1][2]);
            big = fmax3(big, t[1][3], t[2][0]);
            big = fmax3(big, t[2][1], t[2][2]);
            big = fmax3(big, t[2][3], t[3][0]);
            big = fmax3(big, t[3][1], t[3][2]);
            big = fmaxf(big, t[3][3]);
            if (big < noise_limit) {
            // u[r][c] = sum_b sum_a wy_b wx_a t[r+b][c+a]; gx = sum_r k_r (u[r][0] - u[r][2]), k = (1, 2, 1)
            float h[4], v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) h[r] = fmaf(tx, t[r][1] - t[r][3], sx * (t[r][0] - t[r][2]));
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = fmaf(ty, t[1][c] - t[3][c], sy * (t[0][c] - t[2][c]));
            const float hx0 = fmaf(ty, h[1], sy * h[0]), hx1 = fmaf(ty, h[2], sy * h[1]), hx2 = fmaf(ty, h[3], sy * h[2]);
            const float vy0 = fmaf(tx, v[1], sx * v[0]), vy1 = fmaf(tx, v[2], sx * v[1]), vy2 = fmaf(tx, v[3], sx * v[2]);
            gx = hx0 + 2.0f * hx1 + hx2;
            gy = vy0 + 2.0f * vy1 + vy2;
            if (WithCentre)
                centre = fmaf(sy, fmaf(sx, t[1][1], tx * t[1][2]), ty * fmaf(sx, t[2][1], tx * t[2][2]));
            else
                centre = 0.0f;
            // A (near-)vanishing gradient — the mid-line of a corridor in the distance map, a spawn line
            // placed exactly on it (evacuation.toml) — is where the reference's result is either rounding
            // noise or exactly (0, 0), and (0, 0) normalises to NaN and removes the pedestrian at the next
            // rebuild (sfm.rs:108,190). Only the reference's own operation order
            // reproduces which of the two happens, so such samples are redone below.
            if (fmaf(gx, gx, gy * gy) > flat_limit2) return;
            }
        }
    }
    sobel_sample<WithCentre>(g, ny, nx, q, gx, gy, centre);
}

// Pair repulsion of one agent against its three candidate ranges held in the shared-memory tile
// (sfm.rs:112-156). cur/stop are TILE indices. Rounds of SCAN (branch-free: the cut-off test in IEEE ops so
// the neighbor SET is exact in both modes, an unconditional 16-bit store and a predicated advance of the
// write position) and FORCE (the converged heavy loop, two neighbours in flight for ILP, summed in index
// order). A round scans at most as many candidates as the list has free slots, so any density works; at
// 1 ped/m^2 one round covers everything.
//
// Everything is addressed by 32-bit addresses in the CTA's shared window (LDS/STS with immediate
// displacements) and the list holds the address of the neighbour's tile entry (16 bits are enough: a CTA owns
// ~21 KB), so neither loop spends instructions turning indices into addresses. SCAN takes two candidates per iteration
// (rows hold ~6 per lane, ~9 per warp): the second one is masked off past the end of the range; its
// unconditional store may land one slot past the list (hence kListDepth + 1 rows) and its read one entry
// past the range (the tile has a spare entry).
__device__ __forceinline__ float2 lds_f2(uint32_t sa) {  // sa: address in the shared window
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa));
    return v;
}
template <uint32_t kDelta>
__device__ __forceinline__ float2 lds_f2_at(uint32_t sa) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(sa), "n"(kDelta));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t sa) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(sa));
    return v;
}

// SCAN of the two candidates at shared addresses sa and sa + 8 (sfm.rs:130-135): squared distance in IEEE
// ops (packed: sub, mul, then one add), `!(d2 > 4)` (true for NaN, like the reference's `continue` on `>`),
// the second candidate masked off at sa_end, `self` excluded in the middle row. Each candidate's address is
// stored at the write position unconditionally and the position advances by one row if it is in range.
// Written in PTX because the compiler's version of this loop carried three induction variables, a 16-bit
// shadow of the index and select + add instead of a predicated add: 19 instructions per candidate, now ~10.
#define PEDONI_SCAN_PAIR_ASM(SELF_TEST)                                                                       \
    asm volatile(                                                                                             \
        "{\n\t"                                                                                               \
        ".reg .pred p0, p1;\n\t"                                                                              \
        ".reg .b64 q0, q1;\n\t"                                                                               \
        ".reg .f32 a0, b0, a1, b1;\n\t"                                                                       \
        ".reg .b16 h0, h1;\n\t"                                                                               \
        ".reg .u32 s1;\n\t"                                                                                   \
        "ld.shared.b64 q0, [%1];\n\t"                                                                         \
        "ld.shared.b64 q1, [%1+8];\n\t"                                                                       \
        "add.u32 s1, %1, 8;\n\t"                                                                              \
        "sub.rn.f32x2 q0, %3, q0;\n\t"                                                                        \
        "sub.rn.f32x2 q1, %3, q1;\n\t"                                                                        \
        "mul.rn.f32x2 q0, q0, q0;\n\t"                                                                        \
        "mul.rn.f32x2 q1, q1, q1;\n\t"                                                                        \
        "mov.b64 {a0, b0}, q0;\n\t"                                                                           \
        "mov.b64 {a1, b1}, q1;\n\t"                                                                           \
        "add.rn.f32 a0, a0, b0;\n\t"                                                                          \
        "add.rn.f32 a1, a1, b1;\n\t"                                                                          \
        "setp.lt.u32 p1, s1, %2;\n\t"                                                                         \
        "setp.leu.f32 p0, a0, 0f40800000;\n\t"                                                                \
        "setp.leu.and.f32 p1, a1, 0f40800000, p1;\n\t" SELF_TEST                                              \
        "cvt.u16.u32 h0, %1;\n\t"                                                                             \
        "cvt.u16.u32 h1, s1;\n\t"                                                                             \
        "st.shared.u16 [%0], h0;\n\t"                                                                         \
        "@p0 add.u32 %0, %0, 64;\n\t"                                                                         \
        "st.shared.u16 [%0], h1;\n\t"                                                                         \
        "@p1 add.u32 %0, %0, 64;\n\t"                                                                         \
        "}"                                                                                                   \
        : "+r"(wp)                                                                                            \
        : "r"(sa), "r"(sa_end), "l"(pos2), "r"(self_sa)                                                       \
        : "memory")

template <bool kExcludeSelf>
__device__ __forceinline__ void scan_pair(uint32_t sa, uint32_t sa_end, uint32_t self_sa, f32x2 pos2, uint32_t& wp) {
    if (kExcludeSelf)
        PEDONI_SCAN_PAIR_ASM("setp.ne.and.u32 p0, %1, %4, p0;\n\tsetp.ne.and.u32 p1, s1, %4, p1;\n\t");
    else
        PEDONI_SCAN_PAIR_ASM("");
}
#undef PEDONI_SCAN_PAIR_ASM

// One row range of one lane: scan from cur up to stop or until the list is full (`more`).
template <bool kExcludeSelf>
__device__ __forceinline__ void scan_row(uint32_t& cur, uint32_t stop, uint32_t tile_sa, uint32_t col_sa, uint32_t self_sa,
                                         f32x2 pos2, uint32_t& wp, bool& more) {
    if (more) return;
    const uint32_t room = (col_sa + static_cast<uint32_t>(kListDepth) * 64u - wp) / 64u;
    const uint32_t lim = min(stop, cur + room);
    const uint32_t sa_end = tile_sa + lim * 8u;
#pragma unroll 1
    for (uint32_t sa = tile_sa + cur * 8u; sa < sa_end; sa += 16u) scan_pair<kExcludeSelf>(sa, sa_end, self_sa, pos2, wp);
    cur = lim;
    more = lim < stop;
}

template <Math M>
__device__ __forceinline__ void pair_forces_tiled(float2 pos, float2 e, uint32_t tile_sa, uint32_t col_sa,
                                                  uint32_t (&cur)[3], const uint32_t (&stop)[3], uint32_t self,
                                                  float2& acc) {
    constexpr uint32_t kSlot = 32u * sizeof(uint16_t);           // one list row: a 16-bit entry per lane (the 64 in scan_pair)
    constexpr uint32_t kVelDelta = sizeof(float2) * kTileAlloc;  // tile_vel entry = tile_pos entry + this
    static_assert(kSlot == 64, "scan_pair advances the write position by 64 bytes");
    const f32x2 pos2 = pack2(pos);
    const uint32_t self_sa = tile_sa + self * 8u;
    bool more;
    do {
        uint32_t wp = col_sa;  // this lane's next free list slot
        more = false;
        scan_row<false>(cur[0], stop[0], tile_sa, col_sa, self_sa, pos2, wp, more);
        scan_row<true>(cur[1], stop[1], tile_sa, col_sa, self_sa, pos2, wp, more);
        scan_row<false>(cur[2], stop[2], tile_sa, col_sa, self_sa, pos2, wp, more);
#pragma unroll kForceUnroll
        for (uint32_t r = col_sa; r < wp; r += kSlot) {
            const uint32_t o = lds_u16(r);
            const float2 po = lds_f2(o), vo = lds_f2_at<kVelDelta>(o);
            PairTerm<(PEDONI_FAST_PAIR ? M : Math::Strict)>::add(pos, e, po, vo, acc);
        }
    } while (more);
}

// Same, candidates read from global memory (CTAs whose windows do not fit the tile). Correctness path.
template <Math M>
__device__ __forceinline__ void pair_forces_global(float2 pos, float2 e, const float2* __restrict__ g_pos,
                                                   const float2* __restrict__ g_vel, const uint32_t (&beg)[3],
                                                   const uint32_t (&stop)[3], uint32_t self, float2& acc) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (uint32_t c = beg[d]; c < stop[d]; ++c) {
            if (d == 1 && c == self) continue;
            const float2 po = __ldg(g_pos + c);
            const float dx = S::sub(pos.x, po.x), dy = S::sub(pos.y, po.y);
            if (S::add(S::mul(dx, dx), S::mul(dy, dy)) > 4.0f) continue;
            PairTerm<M>::add(pos, e, po, __ldg(g_vel + c), acc);
        }
    }
}

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- bulk (TMA) staging: one mbarrier per warp, used for exactly one phase -----------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar_sa, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_sa), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void mbar_inval(uint32_t mbar_sa) {  // before the memory is used for anything else
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar_sa) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar_sa, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar_sa, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(mbar_sa), "r"(parity)
        : "memory");
    return done != 0;
}
// global -> shared, 16-byte aligned on both sides, bytes a multiple of 16; completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_sa, const void* src, uint32_t bytes, uint32_t mbar_sa) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_sa),
                 "l"(src), "r"(bytes), "r"(mbar_sa)
                 : "memory");
}

// Fast math: is the wall term below 1e-17 m/s^2 at field coordinate q (FieldView::far_mask)? The mask is a few hundred
// KB and the warps of a CTA ask for the same words: an L1 hit almost always.
__device__ __forceinline__ bool far_from_walls_at(const FieldView& f, float2 q) {
    const int tx = __float2int_rz(floorf(q.x)), ty = __float2int_rz(floorf(q.y));
    if (tx < 0 || ty < 0 || tx >= f.fx || ty >= f.fy) return false;
    const uint32_t block = static_cast<uint32_t>(ty >> kFarShift) * static_cast<uint32_t>(f.far_bw) + static_cast<uint32_t>(tx >> kFarShift);
    return ((__ldg(f.far_mask + (block >> 5)) >> (block & 31u)) & 1u) != 0u;
}

template <Math M, bool kDistanceMap, bool kTex>
__global__ void __launch_bounds__(kForceThreads, M == Math::Fast ? PEDONI_FORCE_MIN_BLOCKS : PEDONI_FORCE_MIN_BLOCKS_STRICT)
    force_integrate_kernel(ForceParams p) {
    static_assert(!(kTex && M == Math::Strict), "strict math never reads the maps through textures");
    using O = Ops<M>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* tile_pos = reinterpret_cast<float2*>(smem_raw + kWarpSmemBytes * warp);
    float2* tile_vel = tile_pos + kTileAlloc;
    (void)tile_vel;  // only the cp.async staging variant names it

    const uint32_t* range = blockIdx.y == 0 ? p.d_range : p.d_range_hi;
    const uint32_t begin = range[0], end = range[1];
    const uint32_t block_first = begin + blockIdx.x * kForceThreads;
    if (block_first >= end) return;  // whole CTA beyond the live range (grids are sized from an upper bound)
    const uint32_t warp_first = block_first + warp * 32;
    // The segment-wall variant has block-wide barriers further down: its idle warps must stay.
    if (kDistanceMap && warp_first >= end) return;
    const uint32_t id = warp_first + lane;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t lo = max(begin, p.d_owned[0]), hi = min(end, p.d_owned[1]);
        if (hi > lo) atomicAdd(p.updates_total, static_cast<unsigned long long>(hi - lo));
    }
    const bool live = id < end;
    const int last_lane = warp_first < end ? static_cast<int>(min(32u, end - warp_first)) - 1 : 0;
    const uint32_t tile_sa = static_cast<uint32_t>(__cvta_generic_to_shared(tile_pos));
    const uint32_t mbar_sa = tile_sa + static_cast<uint32_t>(kWarpMbarOffset);
    if (PEDONI_BULK_STAGE) {
        if (lane == 0) mbar_init(mbar_sa, 1);
        __syncwarp();
    }

    if (PEDONI_PREFETCH_AHEAD > 0) {
        // the state of the warp that runs about one wave later: DRAM -> L2 now, so that its first loads are L2 hits
        const uint32_t ahead = warp_first + static_cast<uint32_t>(PEDONI_PREFETCH_AHEAD);
        if (ahead + 32u <= end) {
            if ((lane & 15) == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in.pos + ahead + lane));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in.vel + ahead + lane));
            }
            if (lane == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in.v0 + ahead));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in.dest + ahead));
            }
        }
    }

    float2 pos = make_float2(0.f, 0.f), vel = pos, e = pos, acc = pos;
    float v0 = 0.f;
    uint32_t dest = 0;
    int row = 0;
    uint32_t r_beg[3] = {0, 0, 0}, r_end[3] = {0, 0, 0};  // candidate index ranges, rows cy-1, cy, cy+1
    if (live) {
        pos = p.in.pos[id];
        vel = p.in.vel[id];
        v0 = p.in.v0[id];
        dest = p.in.dest[id];

        // ---- neighbour ranges (sfm.rs:113-127). The reference clamps to the grid; the local table may
        // start at row_base (slabs) and always holds every row a live agent can reach.
        const int2 c = cell_of(pos, p.grid.unit);
        row = c.y;
        const int ly = c.y - p.grid.row_base;
        const int x_start = min(max(c.x - 1, 0), p.grid.nx - 1), x_end = max(min(c.x + 1, p.grid.nx - 1), 0);
        const uint32_t unx = static_cast<uint32_t>(p.grid.nx), table_end = static_cast<uint32_t>(p.grid.table_rows) * unx;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            // Two unconditional loads per row with selected 32-bit indices. A row off the table is an empty range,
            // placed (table start / table end) so that the warp's windows stay monotone.
            const int y = ly + d - 1;
            const uint32_t row0 = static_cast<uint32_t>(y) * unx;
            const bool below = y < 0, above = y >= p.grid.table_rows;
            const uint32_t ib = below ? 0u : (above ? table_end : row0 + static_cast<uint32_t>(x_start));
            const uint32_t ie = below ? 0u : (above ? table_end : row0 + static_cast<uint32_t>(x_end) + 1u);
            if (PEDONI_IN_BOUNDS(ib <= p.table_cells && ie <= p.table_cells, p.error_flag)) {
                r_beg[d] = __ldg(p.cell_start + ib);
                r_end[d] = __ldg(p.cell_start + ie);
            }
        }
    }

    // ---- the warp's three index windows -> its tile [w0 | w1 | w2], staged asynchronously
    const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, r_beg[0], 0), w1 = __shfl_sync(0xFFFFFFFFu, r_beg[1], 0),
                   w2 = __shfl_sync(0xFFFFFFFFu, r_beg[2], 0);
    const uint32_t e0 = __shfl_sync(0xFFFFFFFFu, r_end[0], last_lane), e1 = __shfl_sync(0xFFFFFFFFu, r_end[1], last_lane),
                   e2 = __shfl_sync(0xFFFFFFFFu, r_end[2], last_lane);
#if PEDONI_BULK_STAGE
    // Windows rounded outward to even indices: both ends of a bulk copy must be 16-byte aligned (the arrays
    // have two spare entries for this). a_d: first staged index, n_d: staged entries, window d starts at
    // tile index n_0 + ... + n_{d-1}.
    const uint32_t a0 = w0 & ~1u, a1 = w1 & ~1u, a2 = w2 & ~1u;
    const uint32_t n0 = ((e0 + 1u) & ~1u) - a0, n1 = ((e1 + 1u) & ~1u) - a1, n2 = ((e2 + 1u) & ~1u) - a2;
    const bool tiled = e0 >= w0 && e1 >= w1 && e2 >= w2 &&
                       (static_cast<uint64_t>(n0) + n1 + n2) <= static_cast<uint64_t>(kTileEntries) &&
                       // (debug builds) the staged windows, rounded outward, stay inside the arrays' cap + 2 entries
                       PEDONI_IN_BOUNDS(static_cast<uint64_t>(a0) + n0 <= p.cap + 2ull && static_cast<uint64_t>(a1) + n1 <= p.cap + 2ull &&
                                            static_cast<uint64_t>(a2) + n2 <= p.cap + 2ull, p.error_flag);
    if (tiled && lane == 0) {
        constexpr uint32_t kVel = sizeof(float2) * kTileAlloc;
        mbar_expect_tx(mbar_sa, (n0 + n1 + n2) * 16u);
        if (n0) {
            bulk_g2s(tile_sa, p.in.pos + a0, n0 * 8u, mbar_sa);
            bulk_g2s(tile_sa + kVel, p.in.vel + a0, n0 * 8u, mbar_sa);
        }
        if (n1) {
            bulk_g2s(tile_sa + n0 * 8u, p.in.pos + a1, n1 * 8u, mbar_sa);
            bulk_g2s(tile_sa + kVel + n0 * 8u, p.in.vel + a1, n1 * 8u, mbar_sa);
        }
        if (n2) {
            bulk_g2s(tile_sa + (n0 + n1) * 8u, p.in.pos + a2, n2 * 8u, mbar_sa);
            bulk_g2s(tile_sa + kVel + (n0 + n1) * 8u, p.in.vel + a2, n2 * 8u, mbar_sa);
        }
    }
#else
    const uint32_t a0 = w0, a1 = w1, a2 = w2;
    const uint32_t n0 = e0 - w0, n1 = e1 - w1, n2 = e2 - w2;
    const bool tiled = e0 >= w0 && e1 >= w1 && e2 >= w2 &&
                       (static_cast<uint64_t>(n0) + n1 + n2) <= static_cast<uint64_t>(kTileEntries);
    if (tiled) {
        for (uint32_t k = lane; k < n0 + n1 + n2; k += 32) {
            const uint32_t i = k < n0 ? w0 + k : (k < n0 + n1 ? w1 + (k - n0) : w2 + (k - n0 - n1));
            cp_async_8(tile_pos + k, p.in.pos + i);
            cp_async_8(tile_vel + k, p.in.vel + i);
        }
    }
#endif

    // ---- steering (sfm.rs:106-109; field.rs:248-252) and walls from the distance map (sfm.rs:188-192;
    // field.rs:242-245,255-258), evaluated while the tile copies are in flight.
    float2 wall = make_float2(0.f, 0.f);
    if (live) {
        const float2 q = field_coord(pos, p.field);
        // fast-path guards of field_gradient: through-wall values, and gradients below 0.1 % of the
        // nominal Sobel magnitude 8 * unit (|grad| = 1 for a distance-like map)
        const float noise_limit = 1.0e5f * p.field.unit;
        const float flat_limit2 = (8.0e-3f * p.field.unit) * (8.0e-3f * p.field.unit);
        float gx, gy, unused;
        bool far_from_walls = false;
        if (PEDONI_FAR_LOOKUP_EARLY && M == Math::Fast && kDistanceMap && p.field.far_mask != nullptr)
            far_from_walls = far_from_walls_at(p.field, q);
        // dest < n_maps is guaranteed by the rebuild that admitted this agent (sort_key).
        int2 tile = make_int2(0, 0);
        if (kTex) {
            const int t = 1 + static_cast<int>(dest);
            tile = make_int2((t & (p.field.atlas_tiles_x - 1)) * p.field.fx, (t >> p.field.atlas_shift) * p.field.fy);
        }
        field_gradient<M, false, kTex>(p.field.potential_maps + static_cast<size_t>(dest) * p.field.fy * p.field.fx,
                                       p.field.atlas, tile, p.field.fy, p.field.fx, q, noise_limit, flat_limit2, gx, gy, unused);
        const float rlen = inv_length<M>(gx, gy);
        e = make_float2(O::mul(gx, rlen), O::mul(gy, rlen));
        acc.x = O::add(acc.x, O::div(O::sub(O::mul(e.x, v0), vel.x), 0.5f));  // x / 0.5 == x * 2 exactly
        acc.y = O::add(acc.y, O::div(O::sub(O::mul(e.y, v0), vel.y), 0.5f));
        if (!PEDONI_FAR_LOOKUP_EARLY && M == Math::Fast && kDistanceMap && p.field.far_mask != nullptr)
            far_from_walls = far_from_walls_at(p.field, q);
        if (kDistanceMap && !far_from_walls) {
            float dgx, dgy, distance;
            field_gradient<M, true, kTex>(p.field.distance_map, p.field.atlas, make_int2(0, 0), p.field.fy, p.field.fx, q,
                                          noise_limit, flat_limit2, dgx, dgy, distance);
            const float rl = inv_length<M>(dgx, dgy);
            const float coef = O::mul(10.0f * 0.2f, O::exp(O::div(-distance, 0.2f)));
            wall = make_float2(O::mul(coef, -O::mul(dgx, rl)), O::mul(coef, -O::mul(dgy, rl)));
            if (M == Math::Fast && PEDONI_WALL_EARLY_ADD) {  // fast math sums in any order: two registers fewer across the pair loops
                acc.x += wall.x;
                acc.y += wall.y;
            }
        }
    }
#if PEDONI_BULK_STAGE
    if (tiled) {  // bounded: a lost completion must not hang the GPU
        uint32_t spins = 0;
        while (!mbar_try_wait(mbar_sa, 0) && ++spins < (1u << 24)) {}
        if (spins >= (1u << 24)) atomicOr(p.error_flag, kErrStageTimeout);
    }
#else
    cp_async_wait_all();
#endif
    __syncwarp();
#if PEDONI_BULK_STAGE
    if (lane == 0) mbar_inval(mbar_sa);  // one phase per warp; the segment-wall variant reuses this memory below
#endif

    // ---- pair repulsion (sfm.rs:112-156)
    if (live) {
        if (tiled) {
            // candidate c of row offset d sits at tile index c - off[d]
            const uint32_t off[3] = {a0, a1 - n0, a2 - n0 - n1};
            uint32_t cur[3] = {r_beg[0] - off[0], r_beg[1] - off[1], r_beg[2] - off[2]};
            const uint32_t stop[3] = {r_end[0] - off[0], r_end[1] - off[1], r_end[2] - off[2]};
            const uint32_t col_sa = tile_sa + 2u * sizeof(float2) * kTileAlloc + 2u * lane;
            pair_forces_tiled<M>(pos, e, tile_sa, col_sa, cur, stop, id - off[1], acc);
        } else {
            pair_forces_global<M>(pos, e, p.in.pos, p.in.vel, r_beg, r_end, id, acc);
        }
        if (kDistanceMap && !(M == Math::Fast && PEDONI_WALL_EARLY_ADD)) {  // the reference's order: steering, pairs, wall (sfm.rs:106-192)
            acc.x = O::add(acc.x, wall.x);
            acc.y = O::add(acc.y, wall.y);
        }
    }

    // ---- walls, segment variant (sfm.rs:193-237): obstacles staged through shared memory (the tile is free now)
    if (!kDistanceMap) {
        float* s_edges = reinterpret_cast<float*>(smem_raw);
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
    }
    if (!live) return;

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

    if (!PEDONI_IN_BOUNDS(id < p.cap, p.error_flag)) return;
    p.out.pos[id] = pn;
    p.out.vel[id] = vn;
    p.out.v0[id] = v0;
    p.out.dest[id] = dest;
    // ghost-row agents are integrated twice (here and by their owner): only the owner counts an arrival
    const bool owned = id >= p.d_owned[0] && id < p.d_owned[1];
    // the logical index of a resident pedestrian in the next rebuild's input is its array index (see locate())
    enroll(p.cs, sort_key(p.grid, p.field, pn, dest, p.error_flag, p.arrived, owned, kTex), id, p.error_flag);
    // Slab handles exchange two ghost rows per tick, which covers every move of less than one grid row
    // (1.4 m per 0.1 s); anything faster would silently vanish at a slab boundary, so flag it.
    if (p.grid.slab) {
        const int new_row = __float2int_rz(S::div(pn.y, p.grid.unit));
        if (abs(new_row - row) >= 2 && pn.y == pn.y) atomicOr(p.error_flag, kErrRowJump);
    }
}

}  // namespace pedoni

// sfm_device.cuh — device-side arithmetic shared by the grid-sort and force kernels (sm_100a).
//
// Two arithmetic modes (include/pedoni_cuda.h PedoniMathMode):
//   Strict : every op is an IEEE round-to-nearest intrinsic (__fadd_rn, __fmul_rn, __fdiv_rn,
//            __fsqrt_rn) so nvcc cannot contract a*b+c into FMA — the Rust reference never fuses.
//            exp is evaluated in fp64 and rounded once, which reproduces glibc's (correctly rounded
//            in all but ~1e-9 of cases) expf that the reference calls through f32::exp.
//   Fast   : MUFU-based rcp / rsqrt / sqrt / ex2 approximations, FMA contraction allowed.
// Cell keys, the despawn predicate and all field sampling are Strict in BOTH modes: cell assignment
// and neighbor sets must be bit-exact given identical positions (neighbor_grid.rs:27, sfm.rs:69).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pedoni {

enum class Math : int { Strict = 0, Fast = 1 };

template <Math M>
struct Ops;

template <>
struct Ops<Math::Strict> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float rcp(float a) { return __fdiv_rn(1.0f, a); }
    static __device__ __forceinline__ float exp(float a) { return static_cast<float>(::exp(static_cast<double>(a))); }
};

template <>
struct Ops<Math::Fast> {
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }  // constants fold
    static __device__ __forceinline__ float sqrt(float a) {
        float r;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float rcp(float a) {
        float r;
        asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float exp(float a) { return __expf(a); }
};

using S = Ops<Math::Strict>;

// ---- packed fp32x2 (sm_100: FADD2 / FMUL2 / FFMA2) ---------------------------------------------------
// One issue slot for the x and y component of a Vec2 operation. `.rn` forms are IEEE per component,
// i.e. bit-identical to two scalar `_rn` intrinsics — usable on the exact paths too.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ f32x2 pack2(float2 v) { return pack2(v.x, v.y); }
__device__ __forceinline__ float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Field samplers — util.rs:44-58 `bilinear`, util.rs:61-75 `sobel_filter`, field.rs:235-258.
// Maps are row-major (y, x) f32 (ndarray Array2 of shape (fy, fx), util.rs:29-40). An out-of-bounds
// tap reads 1e12 (util.rs:45,53-56). No texture hardware: the reference's own OpenCL path uses it
// (sfm_gpu.cl:4-5) and diverges from the CPU model (9-bit weights, clamp addressing).
// ------------------------------------------------------------------------------------------------
constexpr int kFarShift = 3;  // 8 x 8 texels (2 m x 2 m at the default field unit) per bit of FieldView::far_mask
struct FieldView {
    float unit;
    float inv_unit;  // 1 / unit when unit is a power of two (then x * inv_unit == x / unit bit for bit), else 0
    int fy, fx;
    int n_maps;
    const float* __restrict__ distance_map;
    const float* __restrict__ potential_maps;
    // The same maps once more, tiled into ONE point-sampled 2D CUDA array (an atlas: one texture handle for
    // the whole warp) for the force kernel's 4x4 footprints (texture gather, see footprint_gather in
    // force.cuh). Tile 0 is the distance map, tile 1 + k potential map k; tile t sits at texel
    // ((t % atlas_tiles_x) * fx, (t / atlas_tiles_x) * fy), atlas_tiles_x a power of two. atlas == 0: not available (strict handles never
    // make one).
    cudaTextureObject_t atlas;
    int atlas_tiles_x;  // a power of two: tile t sits at ((t & (tiles_x - 1)) * fx, (t >> atlas_shift) * fy)
    int atlas_shift;    // log2(atlas_tiles_x)
    // Fast math, distance-map walls: one bit per block of kFarBlock x kFarBlock texels, set where the wall term
    // (10 * 0.2 * exp(-distance / 0.2), sfm.rs:191) is below 1e-17 m/s^2 for every position whose field coordinate
    // floors into the block AND its direction cannot be NaN there (far_mask_kernel in pedoni_cuda.cu). The force
    // kernel skips the distance map's footprint for such positions. nullptr: no mask (strict handles, segment
    // walls, PEDONI_WALL_CUTOFF=0). far_bw: blocks per row of the mask.
    int far_bw;
    const uint32_t* __restrict__ far_mask;
};

__device__ __forceinline__ float field_tap(const float* __restrict__ g, int ny, int nx, int x, int y) {
    return (x >= 0 && y >= 0 && x < nx && y < ny) ? __ldg(g + static_cast<size_t>(y) * nx + x) : 1e12f;
}

// One axis of `bilinear`'s setup for p: base = floor(p); t = p - base; s = 1 - t; i = base as i32.
struct Axis {
    float t, s;
    int i;
};
__device__ __forceinline__ Axis axis_of(float p) {
    float base = floorf(p);
    Axis a;
    a.t = S::sub(p, base);
    a.s = S::sub(1.0f, a.t);
    a.i = __float2int_rz(base);  // cvt.rzi.s32.f32: saturating, NaN -> 0 == Rust `as i32`
    return a;
}

// util.rs:52-57, taps g00=(ix,iy) g01=(ix+1,iy) g10=(ix,iy+1) g11=(ix+1,iy+1).
__device__ __forceinline__ float bilinear_combine(const Axis& ax, const Axis& ay, float g00, float g01, float g10,
                                                  float g11) {
    float y = 0.0f;
    y = S::add(y, S::mul(S::mul(ay.s, ax.s), g00));
    y = S::add(y, S::mul(S::mul(ay.s, ax.t), g01));
    y = S::add(y, S::mul(S::mul(ay.t, ax.s), g10));
    y = S::add(y, S::mul(S::mul(ay.t, ax.t), g11));
    return y;
}

__device__ __forceinline__ float bilinear(const float* __restrict__ g, int ny, int nx, float px, float py) {
    Axis ax = axis_of(px), ay = axis_of(py);
    return bilinear_combine(ax, ay, field_tap(g, ny, nx, ax.i, ay.i), field_tap(g, ny, nx, ax.i + 1, ay.i),
                            field_tap(g, ny, nx, ax.i, ay.i + 1), field_tap(g, ny, nx, ax.i + 1, ay.i + 1));
}

// `position / unit - 0.5` (field.rs:236,243,250,256): true divide, then subtract.
__device__ __forceinline__ float2 field_coord(float2 pos, const FieldView& f) {
    if (f.inv_unit != 0.0f)  // default unit 0.25: scaling by a power of two is exact, no IEEE divide needed
        return make_float2(S::sub(S::mul(pos.x, f.inv_unit), 0.5f), S::sub(S::mul(pos.y, f.inv_unit), 0.5f));
    return make_float2(S::sub(S::div(pos.x, f.unit), 0.5f), S::sub(S::div(pos.y, f.unit), 0.5f));
}

// field.rs:235-239
__device__ __forceinline__ float get_potential(const FieldView& f, uint32_t waypoint, float2 pos, bool use_atlas = false) {
    float2 q = field_coord(pos, f);
    if (use_atlas) {
        // the 2x2 footprint of a bilinear sample is exactly one texture gather on the atlas (component
        // order: see footprint_gather in force.cuh); same texels, same combine -> same bits
        const Axis ax = axis_of(q.x), ay = axis_of(q.y);
        if (ax.i >= 0 && ay.i >= 0 && ax.i + 1 < f.fx && ay.i + 1 < f.fy) {
            const int t = 1 + static_cast<int>(waypoint);
            const float4 g = tex2Dgather<float4>(f.atlas, static_cast<float>((t & (f.atlas_tiles_x - 1)) * f.fx + ax.i) + 1.0f,
                                                 static_cast<float>((t >> f.atlas_shift) * f.fy + ay.i) + 1.0f, 0);
            return bilinear_combine(ax, ay, g.w, g.z, g.x, g.y);
        }
    }
    return bilinear(f.potential_maps + static_cast<size_t>(waypoint) * f.fy * f.fx, f.fy, f.fx, q.x, q.y);
}

// Sobel gradient of the bilinear interpolant at q (util.rs:61-75), optionally also the centre
// sample (field.rs:242-245 for the wall term). The nine bilinear evaluations of the reference touch
// a 4x4 texel footprint; the common case (consecutive bases, footprint inside the map) loads those
// 16 texels once. Each of the 8 (9) samples is then combined with exactly the reference's
// per-sample weights and operation order, so the result is bit-identical to 8 (9) independent
// `bilinear` calls. Rounding in `q + 1.0` can make bases non-consecutive; that and the map border
// take the generic path.
template <bool WithCentre>
__device__ __forceinline__ void sobel_sample(const float* __restrict__ g, int ny, int nx, float2 q, float& gx,
                                             float& gy, float& centre) {
    Axis ax[3], ay[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        ax[o] = axis_of(S::add(q.x, static_cast<float>(o - 1)));
        ay[o] = axis_of(S::add(q.y, static_cast<float>(o - 1)));
    }
    float u[3][3];  // u[r][c]: r = y offset + 1, c = x offset + 1 (util.rs:62-69)
    const int x0 = ax[0].i, y0 = ay[0].i;
    const bool consecutive = (ax[1].i == x0 + 1) && (ax[2].i == x0 + 2) && (ay[1].i == y0 + 1) && (ay[2].i == y0 + 2);
    const bool inside = x0 >= 0 && y0 >= 0 && x0 + 3 < nx && y0 + 3 < ny;
    if (consecutive && inside) {
        float tex[4][4];
        const float* base = g + static_cast<size_t>(y0) * nx + x0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) tex[r][c] = __ldg(base + static_cast<size_t>(r) * nx + c);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!WithCentre && r == 1 && c == 1) continue;
                u[r][c] = bilinear_combine(ax[c], ay[r], tex[r][c], tex[r][c + 1], tex[r + 1][c], tex[r + 1][c + 1]);
            }
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!WithCentre && r == 1 && c == 1) continue;
                u[r][c] = bilinear_combine(ax[c], ay[r], field_tap(g, ny, nx, ax[c].i, ay[r].i),
                                           field_tap(g, ny, nx, ax[c].i + 1, ay[r].i),
                                           field_tap(g, ny, nx, ax[c].i, ay[r].i + 1),
                                           field_tap(g, ny, nx, ax[c].i + 1, ay[r].i + 1));
            }
    }
    // util.rs:72-73, left to right.
    gx = S::sub(S::sub(S::sub(S::sub(S::add(S::add(S::add(u[0][0], u[1][0]), u[1][0]), u[2][0]), u[0][2]), u[1][2]),
                       u[1][2]),
                u[2][2]);
    gy = S::sub(S::sub(S::sub(S::sub(S::add(S::add(S::add(u[0][0], u[0][1]), u[0][1]), u[0][2]), u[2][0]), u[2][1]),
                       u[2][1]),
                u[2][2]);
    centre = WithCentre ? u[1][1] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Neighbor grid — neighbor_grid.rs:14-36, util.rs:29-36.
// ------------------------------------------------------------------------------------------------
struct GridView {
    float unit;      // neighbor_grid_unit
    int nx, ny;      // global grid shape (neighbor_grid.rs:15-16)
    int row_base;    // global row of local table row 0 (0 on a whole-domain handle)
    int table_rows;  // rows covered by the local cell table (ny on a whole-domain handle)
    int own_row0;    // global rows [own_row0, own_row1) are owned by this handle
    int own_row1;
    int slab;        // 1 on a slab handle (ghost rows exist), 0 on a whole-domain handle
};

constexpr uint32_t kKeyDrop = 0xFFFFFFFFu;  // out of grid, despawned, or in a row this handle does not own
constexpr uint32_t kKeyFirstSpecial = 0xFFFFFFF0u;

// Device error flag bits (PedoniModel::d_error).
constexpr uint32_t kErrBadDestination = 1u;  // destination >= n_potential_maps (reference: index panic, field.rs:237)
constexpr uint32_t kErrRowJump = 2u;         // slab handle: a pedestrian crossed >= 2 grid rows in one step
constexpr uint32_t kErrHaloOverflow = 4u;    // slab handle: two boundary rows hold more agents than halo_capacity
constexpr uint32_t kErrHaloTimeout = 8u;     // slab handle: a neighbour's strip did not arrive within 20 s
constexpr uint32_t kErrDebugBounds = 128u;   // PEDONI_DEBUG_CHECKS builds: an index left its array (the access is skipped)
constexpr uint32_t kErrSpawnBound = 64u;     // device-side Poisson draw above the host's bound (mean + 10 sigma + 10)
constexpr uint32_t kErrSortOverflow = 32u;   // rebuild: the overflow list of the cell slots ran out (a bug: it is as long as the arrays)
constexpr uint32_t kErrStageTimeout = 16u;   // force kernel: a warp's bulk copies never completed (a bug, not a user error)

// Bounds checks of our own (compute-sanitizer is closed on this pool): a build with -DPEDONI_DEBUG_CHECKS=1
// (scripts/sweep_force.py build --set debug) tests every index the kernels derive from device-side data against its
// array's capacity, raises kErrDebugBounds and SKIPS the access; the GPU test suite is then run against that
// library (PEDONI_CUDA_LIB). The product build compiles the checks out.
#ifndef PEDONI_DEBUG_CHECKS
#define PEDONI_DEBUG_CHECKS 0
#endif
#if PEDONI_DEBUG_CHECKS
#define PEDONI_IN_BOUNDS(cond, error_flag) ((cond) ? true : (atomicOr((error_flag), kErrDebugBounds), false))
#else
#define PEDONI_IN_BOUNDS(cond, error_flag) true
#endif

// `(pos / unit).as_ivec2()` (neighbor_grid.rs:27, sfm.rs:113): IEEE divide, truncate toward zero.
__device__ __forceinline__ int2 cell_of(float2 pos, float unit) {
    return make_int2(__float2int_rz(S::div(pos.x, unit)), __float2int_rz(S::div(pos.y, unit)));
}

// Sort key of an agent for the next rebuild: local cell id, or kKeyDrop.
//  - outside the grid -> dropped (neighbor_grid.rs:29-33, util.rs:31)
//  - potential(dest, pos) > 0.25 is false (incl. NaN) -> despawned (sfm.rs:69)
//  - destination >= n_maps would panic in the reference (index out of bounds); here it drops the
//    agent and raises the device error flag.
//  - slab handles: an agent whose row belongs to another slab is dropped HERE; the owner of that row
//    integrates the same agent redundantly (it holds it as a ghost) and adopts it in its own rebuild.
//  - `arrived` (16 cumulative counters by destination, the last one lumps destinations >= 15): a
//    pedestrian removed BY THE PREDICATE (it reached its destination) is counted once, by the handle in
//    whose rows `count_arrival` says it belongs — the observable "flow" (SURVEY.md section 8, row f3).
__device__ __forceinline__ uint32_t sort_key(const GridView& g, const FieldView& f, float2 pos, uint32_t dest,
                                             uint32_t* error_flag, unsigned long long* arrived = nullptr,
                                             bool count_arrival = false, bool use_atlas = false) {
    int2 c = cell_of(pos, g.unit);
    if (c.x < 0 || c.y < 0 || c.x >= g.nx || c.y >= g.ny) return kKeyDrop;
    if (dest >= static_cast<uint32_t>(f.n_maps)) {
        atomicOr(error_flag, kErrBadDestination);
        return kKeyDrop;
    }
    const float potential = get_potential(f, dest, pos, use_atlas);
    if (!(potential > 0.25f)) {
        if (count_arrival && potential == potential) atomicAdd(arrived + min(dest, 15u), 1ull);
        return kKeyDrop;
    }
    if (c.y < g.own_row0 || c.y >= g.own_row1) return kKeyDrop;
    return static_cast<uint32_t>(c.y - g.row_base) * static_cast<uint32_t>(g.nx) + static_cast<uint32_t>(c.x);
}

}  // namespace pedoni

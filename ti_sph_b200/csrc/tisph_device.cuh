// tisph_device.cuh -- device-side helpers shared by the kernels of libtisph.so.
// Hand-written for sm_100a; no Taichi, no Triton, no tensor cores (the step is a
// cutoff-filtered gather, not a contraction).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tisph {

// Debug build (-DTISPH_CHECKS, `TISPH_CHECKS=1 python -m ti_sph_b200.build --force`): index bounds
// of the shared-memory tiles, pending lists and list-pool rows are checked on the device and
// counted; tisph_get_param(TISPH_P_STAT_CHECK_FAILURES) reads {count, first failing line}.
// compute-sanitizer is not available on the B200 pool, so this is the memory checker of the repo.
__device__ int g_tisph_check[2];
#ifdef TISPH_CHECKS
#define TISPH_CHECK(cond)                                                             \
    do {                                                                              \
        if (!(cond)) { if (atomicAdd(&g_tisph_check[0], 1) == 0) g_tisph_check[1] = __LINE__; } \
    } while (0)
#else
#define TISPH_CHECK(cond) do { } while (0)
#endif

constexpr int MAT_BOUNDARY = 0;   // partice_systemv4.py:24
constexpr int MAT_FLUID = 1;      // partice_systemv4.py:25

// Everything a kernel needs to know about the scene; passed by value (constant bank).
struct SimParams {
    int n;                 // particle_num
    int ncell;
    int gx, gy, gz;        // grid_num
    int dim;
    float h;               // support_length == grid_size
    float inv_h;
    float d2_cut;          // smallest f32 t with sqrtf(t) >= h  (norm(x_ij) < h  <=>  d2 < d2_cut)
    float k_w, k_dw;       // kernel normalisations
    float dt;
    float g[3];
    float pad;
    float wall_hi[3];
    float rho0, ps_density0, stiffness, exponent;
    float visc_fluid_c, visc_bound_c, eps_h2;
    float g1_visc_c, g1_mass, g1_press_c, m_V0;
    int density_mode, volume_mode;
    int int_exponent;      // exponent if it is a small positive integer, else 0
    int lists_only;        // 1: the density walk builds the neighbour lists but skips the (discarded) kernel sum
    int walls;             // 1: the force stage applies enforce_boundary itself (default), 0: TISPH_STAGE_WALLS does
    float one;             // 1.0f the compiler cannot see (tisph_lists.cuh: uncontracted packed sums)
    // slab sharding (tisph_shard.cuh): cell-key ranges [lo, hi).  Unsharded: [0, INT_MAX).
    int own_key_lo, own_key_hi;     // cells whose particles this rank advances
    int walk_key_lo, walk_key_hi;   // cells that get work items (own planes + one ghost plane each side)
    int ghost_walk;                 // 0: ghost cells need no density walk (rho = mass W(0), no boundary volumes)
};

// cell = (int)(x / grid_size): IEEE f32 division then truncation (partice_systemv4.py:86-92)
__device__ __forceinline__ int cell_coord(float x, float h) { return (int)__fdiv_rn(x, h); }

// d2 in the reference's evaluation order with no FMA contraction, so that the neighbour
// predicate is bit-identical to sqrtf(dx*dx+dy*dy+dz*dz) < h evaluated in IEEE f32.
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ float dist2_exact2(float dx, float dy) {
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// Cubic spline W(q)/k (sph_basev2.py:19-36), q in [0,1]
__device__ __forceinline__ float spline_w(float q) {
    float a = fmaf(6.0f * q * q, q - 1.0f, 1.0f);    // 6(q^3-q^2)+1
    float f = 1.0f - q;
    float b = 2.0f * f * f * f;
    return q <= 0.5f ? a : b;
}
// dW/dq / (6k) (sph_basev2.py:53-60): q(3q-2) | -(1-q)^2
__device__ __forceinline__ float spline_dw(float q) {
    float a = q * fmaf(3.0f, q, -2.0f);
    float f = 1.0f - q;
    float b = -f * f;
    return q <= 0.5f ? a : b;
}

__device__ __forceinline__ float fast_rcp(float x) { return __fdividef(1.0f, x); }

// x^e for the Tait EOS (wcsphv2.py:47). Integer exponents by repeated multiplication
// (<= 2 ulp, tighter than powf); anything else through powf.
__device__ __forceinline__ float eos_pow(float x, float e, int ie) {
    if (ie > 0) {
        float r = 1.0f, b = x;
        int k = ie;
        while (k) { if (k & 1) r *= b; b *= b; k >>= 1; }
        return r;
    }
    return powf(x, e);
}

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace tisph

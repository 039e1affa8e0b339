// tisph_gen1.cuh -- the gen-1 (2D) step: ParticleSystem / ParticleSystemV2 + WCSPH of the reference.
//
// Gen-1 never reorders particles.  Every step (`ps.init()`, partice_system.py:211-215) it rebuilds
// dense per-cell particle lists and an explicit neighbour table particle_neighbors[i][100]
// (allocate_particles_to_grid :127-132, search_neighbors :102-121), and the solver kernels walk
// that table (wcsph.py:18-72).  Here the cell lists are one cell-sorted index array (counting
// sort with the gen-2 kernels k_bin / k_scan_* / k_place, then k_g1_order makes the order inside a
// cell ascending = what the serial reference gets from its atomic slot counter), the neighbour
// table is kept in the reference's layout and order, and the three solver passes are one
// thread-per-particle kernel each.  At the sizes gen-1 supports (<= 2^15 particles) everything is
// launch-latency bound; the kernels are written for parity, not for a roofline.
#pragma once
#include "tisph_device.cuh"

namespace tisph {

constexpr int G1_MAX_PER_CELL = 100;      // partice_system.py:25
constexpr int G1_MAX_NEIGHBORS = 100;     // partice_system.py:26

// ids_sorted[cell segment] = particle indices of the cell in ascending order
__global__ void __launch_bounds__(256)
k_g1_order(int n, const int* __restrict__ keys, const int* __restrict__ ids, const int* __restrict__ cell_end,
           int* __restrict__ ids_sorted) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int id = ids[s];
    int key = keys[id];
    int b = key > 0 ? cell_end[key - 1] : 0, e = cell_end[key];
    int cnt = 0;
    for (int t = b; t < e; ++t) cnt += (ids[t] < id) ? 1 : 0;
    ids_sorted[b + cnt] = id;
}

// search_neighbors (partice_system.py:102-121): 3x3 cells in row-major order, `break` out of the
// whole walk at the first invalid cell (quirk Q8), p_j != p_i and norm(x_ij) < support_radius.
// err[1] counts cell-list / neighbour-list overflows (undefined behaviour in the reference).
__global__ void __launch_bounds__(128)
k_g1_neighbors(SimParams sp, const float4* __restrict__ P, const float4* __restrict__ Q,
               const int* __restrict__ cell_end, const int* __restrict__ ids_sorted,
               int* __restrict__ nbr, int* __restrict__ nbr_num, int* __restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    if (__float_as_int(Q[i].z) == MAT_BOUNDARY) return;            // :104-105 (count keeps its old value)
    const float4 pi = P[i];
    const int cx = cell_coord(pi.x, sp.h), cy = cell_coord(pi.y, sp.h);
    int cnt = 0;
    bool stop = false;
    int* row = nbr + (size_t)i * G1_MAX_NEIGHBORS;
    for (int ox = -1; ox <= 1 && !stop; ++ox)
        for (int oy = -1; oy <= 1; ++oy) {
            const int ax = cx + ox, ay = cy + oy;
            if (ax < 0 || ax >= sp.gx || ay < 0 || ay >= sp.gy) { stop = true; break; }
            const int key = ax * sp.gy + ay;
            const int b = key > 0 ? cell_end[key - 1] : 0;
            int e = cell_end[key];
            if (e - b > G1_MAX_PER_CELL) { atomicAdd(err + 1, 1); e = b + G1_MAX_PER_CELL; }
            for (int t = b; t < e; ++t) {
                const int j = ids_sorted[t];
                if (j == i) continue;
                const float4 pj = P[j];
                const float d2 = dist2_exact2(pi.x - pj.x, pi.y - pj.y);
                if (!(d2 < sp.d2_cut)) continue;                    // norm >= support_radius
                if (cnt >= G1_MAX_NEIGHBORS) { atomicAdd(err + 1, 1); continue; }
                row[cnt++] = j;
            }
        }
    nbr_num[i] = cnt;
}

__device__ __forceinline__ float g1_norm(float dx, float dy) { return __fsqrt_rn(dist2_exact2(dx, dy)); }

// compute_densities (wcsph.py:18-32) + clamp / Tait EOS (wcsph.py:37-40) -> D
__global__ void __launch_bounds__(128)
k_g1_density(SimParams sp, const float4* __restrict__ P, const float4* __restrict__ Q,
             const int* __restrict__ nbr, const int* __restrict__ nbr_num, float4* __restrict__ D,
             float* __restrict__ S, int* __restrict__ ncount) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    const float4 pi = P[i];
    const int cnt = nbr_num[i];
    const int* row = nbr + (size_t)i * G1_MAX_NEIGHBORS;
    float rho = 0.f;
    for (int t = 0; t < cnt; ++t) {
        const int j = row[t];
        if (__float_as_int(Q[j].z) != MAT_FLUID) continue;
        const float4 pj = P[j];
        const float q = fminf(g1_norm(pi.x - pj.x, pi.y - pj.y) * sp.inv_h, 1.0f);
        rho += sp.m_V0 * (sp.k_w * spline_w(q));
    }
    const float rho_raw = rho * sp.rho0;                            // :32
    const float rho_c = fmaxf(rho_raw, sp.rho0);                    // :37
    const float pr = sp.stiffness * (eos_pow(rho_c / sp.rho0, sp.exponent, sp.int_exponent) - 1.0f);
    D[i] = make_float4(rho_raw, pr / (rho_c * rho_c), rho_c, pr);
    S[i] = rho_raw;
    ncount[i] = cnt;
}

// compute_non_pressure_force (wcsph.py:52-65, sph_base.py:77-84), compute_pressure_force launch B
// (wcsph.py:42-49, sph_base.py:63-74), advert (wcsph.py:67-72).  enforce_boundary is a no-op in
// the reference (sph_base.py:161-166).
__global__ void __launch_bounds__(128)
k_g1_force(SimParams sp, const float4* __restrict__ Pin, const float4* __restrict__ Vin,
           const float4* __restrict__ Qin, const float4* __restrict__ D, const int* __restrict__ nbr,
           const int* __restrict__ nbr_num, float4* __restrict__ Pout, float4* __restrict__ Vout,
           float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
           float4* __restrict__ a_p_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    const float4 pi = Pin[i], vi = Vin[i], qi = Qin[i], di = D[i];
    float4 pout = pi, vout = vi;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (__float_as_int(qi.z) == MAT_FLUID) {
        const int cnt = nbr_num[i];
        const int* row = nbr + (size_t)i * G1_MAX_NEIGHBORS;
        float ax = 0.f, ay = sp.g[1];                               // d_v[dim-1] = const.g
        float px = 0.f, py = 0.f;
        for (int t = 0; t < cnt; ++t) {
            const int j = row[t];
            const float4 pj = Pin[j], vj = Vin[j], dj = D[j];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y;
            const float d2 = dist2_exact2(dx, dy);
            const float r = __fsqrt_rn(d2);
            const float q = r * sp.inv_h;
            float gfac = 0.f;                                       // gradW = gfac * x_ij  (sph_base.py:37-60)
            if (r > 1e-5f && q <= 1.0f) gfac = sp.k_dw * spline_dw(q) / (r * sp.h);
            const float v_xy = (vi.x - vj.x) * dx + (vi.y - vj.y) * dy;
            const float s = sp.g1_visc_c * (sp.g1_mass / dj.x) * v_xy / (d2 + sp.eps_h2) * gfac;   // :77-84
            ax = fmaf(s, dx, ax); ay = fmaf(s, dy, ay);
            if (__float_as_int(Qin[j].z) == MAT_FLUID) {
                const float sp_ = sp.g1_press_c * (di.y + dj.y) * gfac;                             // :63-70
                px = fmaf(sp_, dx, px); py = fmaf(sp_, dy, py);
            }
        }
        if (a_np_out) {
            a_np_out[i] = make_float4(ax, ay, 0.f, 0.f);
            a_p_out[i] = make_float4(px, py, 0.f, 0.f);
        }
        acc.x = ax + px; acc.y = ay + py;
        vout.x = vi.x + sp.dt * acc.x; vout.y = vi.y + sp.dt * acc.y;
        pout.x = pi.x + sp.dt * vout.x; pout.y = pi.y + sp.dt * vout.y;
    } else if (a_np_out) {
        a_np_out[i] = acc;
        a_p_out[i] = acc;
    }
    Pout[i] = pout;
    Vout[i] = vout;
    Qout[i] = make_float4(di.z, di.w, qi.z, qi.w);                  // clamped rho, p, material, id
    dvel[i] = acc;
}

}  // namespace tisph

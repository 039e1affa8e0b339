// tisph_kernels.cuh -- the sm_100a kernels of the WCSPH step.
//
// Data layout in HBM (all arrays float4 / 16-byte records, capacity-sized, two copies
// "cur" and "other" that ping-pong inside a step):
//   P = {x, y, z, mass}      V = {vx, vy, vz, volume}
//   Q = {density, pressure, material (i32 bits), orig_id (i32 bits)}
//   D = {rho_raw, p/rho_c^2, rho_c, p}   scratch written by the density kernel
// A step is:  bin -> scan -> place -> reorder (cur -> other, now sorted by cell key)
//             density (+boundary volume +EOS)  -> D, S, neighbour count
//             force+advect+walls               -> writes the other copy (sorted order)
#pragma once
#include "tisph_device.cuh"

namespace tisph {

// ---------------------------------------------------------------------------------------
// K1  bin: cell key per particle + histogram       (partice_systemv4.py:206-215)
// Warp-aggregated atomics: lanes with equal keys elect one leader that adds the group size.
// `arrival` is the particle's arrival rank inside its cell (arbitrary order, fixed up in
// k_reorder so that the final order is the stable, serial-reference order).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bin(SimParams sp, const float4* __restrict__ P, int* __restrict__ keys,
      int* __restrict__ arrival, int* __restrict__ cell_count, int* __restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < sp.n;
    int key = 0;
    if (valid) {
        float4 p = P[i];
        int cx = cell_coord(p.x, sp.h);
        int cy = cell_coord(p.y, sp.h);
        int cz = sp.dim == 3 ? cell_coord(p.z, sp.h) : 0;
        bool bad = cx < 0 || cx >= sp.gx || cy < 0 || cy >= sp.gy || cz < 0 || cz >= sp.gz;
        if (bad) {   // reference: out-of-bounds access (UB). Here: flagged, clamped.
            atomicAdd(err, 1);
            cx = min(max(cx, 0), sp.gx - 1);
            cy = min(max(cy, 0), sp.gy - 1);
            cz = min(max(cz, 0), sp.gz - 1);
        }
        key = (cx * sp.gy + cy) * sp.gz + cz;
    }
    unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        int lane = threadIdx.x & 31;
        unsigned peers = __match_any_sync(act, key);
        int leader = __ffs(peers) - 1;
        int rank = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (lane == leader) base = atomicAdd(&cell_count[key], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        keys[i] = key;
        arrival[i] = base + rank;
    }
}

// ---------------------------------------------------------------------------------------
// K2  inclusive scan of the histogram (partice_systemv4.py:255; Taichi PrefixSumExecutor)
// reduce -> spine -> apply; 2048 cells per block.
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const int* __restrict__ in, int n, int* __restrict__ sums) {
    __shared__ int wsum[SCAN_THREADS / 32];
    int base = blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int idx = base + k * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
        sums[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the per-block sums, in place
__global__ void __launch_bounds__(1024)
k_scan_spine(int* __restrict__ sums, int nb) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int idx = base + threadIdx.x;
        int v = idx < nb ? sums[idx] : 0;
        int inc = warp_inclusive_scan(v, lane);
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int t = wtot[lane];
            int ti = warp_inclusive_scan(t, lane);
            wtot[lane] = ti - t;          // exclusive warp offsets
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + wtot[warp] + inc - v;
        if (idx < nb) sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const int* __restrict__ in, int n, const int* __restrict__ sums,
             int* __restrict__ out) {
    __shared__ int wtot[SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    if (idx0 + SCAN_ITEMS <= n) {
        int4 a = *reinterpret_cast<const int4*>(in + idx0);
        int4 b = *reinterpret_cast<const int4*>(in + idx0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (idx0 + k < n) ? in[idx0 + k] : 0;
    }
#pragma unroll
    for (int k = 1; k < SCAN_ITEMS; ++k) v[k] += v[k - 1];
    int tot = v[SCAN_ITEMS - 1];
    int inc = warp_inclusive_scan(tot, lane);
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) woff += (w < warp) ? wtot[w] : 0;
    int off = sums[blockIdx.x] + woff + inc - tot;
    if (idx0 + SCAN_ITEMS <= n) {
        int4 a = make_int4(v[0] + off, v[1] + off, v[2] + off, v[3] + off);
        int4 b = make_int4(v[4] + off, v[5] + off, v[6] + off, v[7] + off);
        *reinterpret_cast<int4*>(out + idx0) = a;
        *reinterpret_cast<int4*>(out + idx0 + 4) = b;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (idx0 + k < n) out[idx0 + k] = v[k] + off;
    }
}

// ---------------------------------------------------------------------------------------
// K3  place: ids[start(key) + arrival] = i      (first half of resort, :219-224)
// K4  reorder: canonicalise each cell segment to ascending original index (== the stable
//     order a serial execution of the reference produces) and move the particle records
//     once, cur -> other (the reference scatters 10 arrays and copies them back, :226-249).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_start(const int* __restrict__ cell_end, int c) {
    return c > 0 ? cell_end[c - 1] : 0;
}

__global__ void __launch_bounds__(256)
k_place(int n, const int* __restrict__ keys, const int* __restrict__ arrival,
        const int* __restrict__ cell_end, int* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keys[i];
    ids[cell_start(cell_end, key) + arrival[i]] = i;
}

__global__ void __launch_bounds__(256)
k_reorder(int n, const int* __restrict__ keys, const int* __restrict__ ids,
          const int* __restrict__ cell_end, const float4* __restrict__ Pin,
          const float4* __restrict__ Vin, const float4* __restrict__ Qin,
          float4* __restrict__ Pout, float4* __restrict__ Vout, float4* __restrict__ Qout,
          int* __restrict__ keys_sorted) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int id = ids[s];
    int key = keys[id];
    int b = cell_start(cell_end, key), e = cell_end[key];
    int cnt = 0;
    for (int t = b; t < e; ++t) cnt += (ids[t] < id) ? 1 : 0;
    int dst = b + cnt;
    Pout[dst] = Pin[id];
    Vout[dst] = Vin[id];
    Qout[dst] = Qin[id];
    keys_sorted[dst] = key;
}

// ---------------------------------------------------------------------------------------
// Neighbour walks.  One CTA per cell.  The 27 neighbour cells are 9 contiguous ranges of the
// sorted arrays (z is the fastest key digit, so cells (x,y,cz-1..cz+1) are adjacent); they
// are staged into shared memory tile by tile and every thread walks them as broadcast
// reads.  The CTA's threads are arranged as  [split][target lane]: 32 or 64 target particles
// of the cell, each walked by 8 or 4 "split" threads that take interleaved 32-candidate
// chunks; partial sums are combined through shared memory.
//
// Candidate range of cell c is [cell_end[max(0,c-1)], cell_end[c])  (partice_systemv4.py:343),
// which makes cell 0 invisible as a neighbour (reference quirk, reproduced). Cells outside
// the grid are empty (the reference reads out of bounds there).
// ---------------------------------------------------------------------------------------
constexpr int NB_THREADS = 256;

struct CellRanges {
    int gb[9];     // first sorted index of each range
    int off[10];   // tile offsets (prefix sum of lengths)
};

__device__ __forceinline__ void compute_cell_ranges(const SimParams& sp,
                                                    const int* __restrict__ cell_end, int c,
                                                    CellRanges& R) {
    __shared__ int s_len[9];
    int tid = threadIdx.x;
    if (tid < 9) {
        int cz = c % sp.gz;
        int cy = (c / sp.gz) % sp.gy;
        int cx = c / (sp.gz * sp.gy);
        int x = cx + tid / 3 - 1, y = cy + tid % 3 - 1;
        int gb = 0, len = 0;
        if (x >= 0 && x < sp.gx && y >= 0 && y < sp.gy) {
            int zlo = max(cz - 1, 0), zhi = min(cz + 1, sp.gz - 1);
            int clo = (x * sp.gy + y) * sp.gz + zlo;
            int chi = clo + (zhi - zlo);
            gb = cell_end[max(clo - 1, 0)];
            len = cell_end[chi] - gb;
        }
        R.gb[tid] = gb;
        s_len[tid] = len;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { R.off[k] = o; o += s_len[k]; }
        R.off[9] = o;
    }
    __syncthreads();
}

__device__ __forceinline__ int tile_to_global(const CellRanges& R, int e) {
    int k = 0;
#pragma unroll
    for (int t = 1; t < 9; ++t) k += (e >= R.off[t]) ? 1 : 0;
    return R.gb[k] + (e - R.off[k]);
}

// ---------------------------------------------------------------------------------------
// K5  density: S_i, neighbour count, boundary volume, clamp + Tait EOS
//     (wcsphv2.py:18-34,45-47 ; sph_basev2.py:190-201)
// ---------------------------------------------------------------------------------------
constexpr int DENS_TCAP = 2048;

__global__ void __launch_bounds__(NB_THREADS)
k_density(SimParams sp, const int* __restrict__ cell_end, const float4* __restrict__ P,
          float4* __restrict__ V, const float4* __restrict__ Q, float4* __restrict__ D,
          float* __restrict__ S, int* __restrict__ ncount) {
    __shared__ CellRanges R;
    __shared__ float4 tile[DENS_TCAP];
    __shared__ float red_w[NB_THREADS];
    __shared__ float red_b[NB_THREADS];
    __shared__ int red_c[NB_THREADS];

    const int c = blockIdx.x;
    const int tb = cell_start(cell_end, c), te = cell_end[c];
    if (te <= tb) return;
    compute_cell_ranges(sp, cell_end, c, R);
    const int total = R.off[9];
    const int tid = threadIdx.x;
    const int nT = te - tb;
    const int tl = nT <= 32 ? 32 : 64;            // target lanes per pass
    const int nsplit = NB_THREADS / tl;
    const int t_local = tid % tl, split = tid / tl;
    const bool akinci = sp.volume_mode == 1;
    const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];

    for (int pass = 0; pass < nT; pass += tl) {
        const int i = tb + pass + t_local;
        const bool active = i < te;
        float4 pi = active ? P[i] : make_float4(1e18f, 1e18f, 1e18f, 0.f);
        int mat_i = active ? __float_as_int(Q[i].z) : MAT_FLUID;
        int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        float wsum = 0.f, wbsum = 0.f;
        int cnt = 0;
        for (int tile0 = 0; tile0 < total; tile0 += DENS_TCAP) {
            const int tile_n = min(DENS_TCAP, total - tile0);
            __syncthreads();
            for (int e = tid; e < tile_n; e += NB_THREADS) {
                int g = tile_to_global(R, tile0 + e);
                float4 p = P[g];
                p.w = akinci ? Q[g].z : 0.f;
                tile[e] = p;
            }
            __syncthreads();
            const int self_t = self_e - tile0;
            for (int cb = split * 32; cb < tile_n; cb += nsplit * 32) {
                const int ce = min(cb + 32, tile_n);
#pragma unroll 4
                for (int e = cb; e < ce; ++e) {
                    float4 cj = tile[e];
                    float dx = pi.x - cj.x, dy = pi.y - cj.y, dz = pi.z - cj.z;
                    float d2 = sp.dim == 3 ? dist2_exact(dx, dy, dz) : dist2_exact2(dx, dy);
                    if (d2 < sp.d2_cut && e != self_t) {
                        cnt++;
                        float r = d2 * rsqrtf(fmaxf(d2, 1e-30f));
                        float w = spline_w(r * sp.inv_h);
                        wsum += w;
                        if (__float_as_int(cj.w) == MAT_BOUNDARY) wbsum += w;
                    }
                }
            }
        }
        red_w[tid] = wsum;
        red_b[tid] = wbsum;
        red_c[tid] = cnt;
        __syncthreads();
        if (split == 0 && active) {
            for (int s = 1; s < nsplit; ++s) {
                wsum += red_w[s * tl + t_local];
                wbsum += red_b[s * tl + t_local];
                cnt += red_c[s * tl + t_local];
            }
            float4 qi = Q[i];
            float rho_raw;
            float s_i = 0.f;
            if (mat_i == MAT_FLUID) {
                float self = pi.w * sp.k_w;                 // mass_i * W(0)
                s_i = pi.w * (sp.k_w * wsum);               // sum_j mass_i W(r_ij)   (Q2)
                rho_raw = sp.density_mode == 1 ? self + s_i : self;
            } else {
                rho_raw = qi.x;                             // boundary keeps its stored density
                // sph_basev2.py:195-201: volume = 1/(W(0) [+ sum over boundary neighbours])
                float delta = sp.k_w + (akinci ? sp.k_w * wbsum : 0.f);
                float4 vi = V[i];
                vi.w = 1.0f / delta;
                V[i] = vi;
            }
            float rho_c = fmaxf(rho_raw, sp.rho0);                                     // :46
            float pr = sp.stiffness * (eos_pow(rho_c / sp.rho0, sp.exponent, sp.int_exponent) - 1.0f);  // :47
            D[i] = make_float4(rho_raw, pr / (rho_c * rho_c), rho_c, pr);
            S[i] = s_i;
            ncount[i] = cnt;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K6  forces + advect + walls, fused
//     (wcsphv2.py:56-93 non-pressure, :43-54 + sph_basev2.py:64-78 pressure,
//      wcsphv2.py:95-100 advert, sph_basev2.py:158-189 walls)
// Reads the sorted copy (Pin,Vin,Qin,D), writes the other copy in the same (sorted) order.
// ---------------------------------------------------------------------------------------
constexpr int FORCE_TCAP = 1792;
constexpr size_t FORCE_SMEM = (size_t)FORCE_TCAP * 3 * sizeof(float4);

__global__ void __launch_bounds__(NB_THREADS, 2)
k_force(SimParams sp, const int* __restrict__ cell_end, const float4* __restrict__ Pin,
        const float4* __restrict__ Vin, const float4* __restrict__ Qin,
        const float4* __restrict__ D, float4* __restrict__ Pout, float4* __restrict__ Vout,
        float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
        float4* __restrict__ a_p_out) {
    extern __shared__ float4 dyn_smem[];
    float4* tP = dyn_smem;                      // {x,y,z,mass}
    float4* tV = dyn_smem + FORCE_TCAP;         // {vx,vy,vz,volume}
    float4* tA = dyn_smem + 2 * FORCE_TCAP;     // {rho_raw, p/rho_c^2, material, -}
    __shared__ CellRanges R;
    __shared__ float red[6][NB_THREADS];

    const int c = blockIdx.x;
    const int tb = cell_start(cell_end, c), te = cell_end[c];
    if (te <= tb) return;
    compute_cell_ranges(sp, cell_end, c, R);
    const int total = R.off[9];
    const int tid = threadIdx.x;
    const int nT = te - tb;
    const int tl = nT <= 32 ? 32 : 64;
    const int nsplit = NB_THREADS / tl;
    const int t_local = tid % tl, split = tid / tl;
    const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];

    for (int pass = 0; pass < nT; pass += tl) {
        const int i = tb + pass + t_local;
        const bool active = i < te;
        float4 pi = active ? Pin[i] : make_float4(1e18f, 1e18f, 1e18f, 1.f);
        float4 vi = active ? Vin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 di = active ? D[i] : make_float4(1.f, 0.f, 1.f, 0.f);
        float4 qi = active ? Qin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int mat_i = __float_as_int(qi.z);
        const bool walker = active && mat_i == MAT_FLUID && i >= sp.owned_lo && i < sp.owned_hi;
        int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        const float coh_i = 0.01f / pi.w;                         // wcsphv2.py:64
        const float nub_i = sp.visc_bound_c / (2.0f * di.x);      // wcsphv2.py:76
        const float rho_i = di.x, pr_i = di.y;
        float anx = 0.f, any = 0.f, anz = 0.f;   // sum of non-pressure terms (to subtract)
        float apx = 0.f, apy = 0.f, apz = 0.f;   // sum of pressure terms
        for (int tile0 = 0; tile0 < total; tile0 += FORCE_TCAP) {
            const int tile_n = min(FORCE_TCAP, total - tile0);
            __syncthreads();
            for (int e = tid; e < tile_n; e += NB_THREADS) {
                int g = tile_to_global(R, tile0 + e);
                tP[e] = Pin[g];
                tV[e] = Vin[g];
                float4 d = D[g];
                tA[e] = make_float4(d.x, d.y, Qin[g].z, 0.f);
            }
            __syncthreads();
            if (walker) {
                const int self_t = self_e - tile0;
                for (int cb = split * 32; cb < tile_n; cb += nsplit * 32) {
                    const int ce = min(cb + 32, tile_n);
#pragma unroll 2
                    for (int e = cb; e < ce; ++e) {
                        float4 pj = tP[e];
                        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        float d2 = dist2_exact(dx, dy, dz);
                        if (d2 < sp.d2_cut && e != self_t) {
                            float4 vj = tV[e];
                            float4 aj = tA[e];
                            float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                            float r = d2 * rinv;
                            float q = r * sp.inv_h;
                            // gradW = k_dw * dw(q) * x_ij / (r h); zero for r <= 1e-5 (sph_basev2.py:53)
                            float gfac = r > 1e-5f ? sp.k_dw * spline_dw(q) * rinv * sp.inv_h : 0.f;
                            float dot = (vi.x - vj.x) * dx + (vi.y - vj.y) * dy + (vi.z - vj.z) * dz;
                            float mn = fminf(dot, 0.f) * fast_rcp(d2 + sp.eps_h2);
                            float cn, cp;
                            if (__float_as_int(aj.z) == MAT_FLUID) {
                                float w = sp.k_w * spline_w(q);
                                float nu = sp.visc_fluid_c * fast_rcp(rho_i + aj.x);     // :69
                                float pi_ij = -nu * mn;                                   // :72
                                cn = coh_i * pj.w * w + pj.w * pi_ij * gfac;              // :64 + :73
                                cp = -pj.w * (pr_i + aj.y) * gfac;                        // sph_basev2.py:71-73
                            } else {
                                float pi_ij = -nub_i * mn;                                // :78
                                cn = sp.ps_density0 * vj.w * pi_ij * gfac;                // :80
                                cp = -sp.rho0 * vj.w * pr_i * gfac;                       // sph_basev2.py:75
                            }
                            anx = fmaf(cn, dx, anx); any = fmaf(cn, dy, any); anz = fmaf(cn, dz, anz);
                            apx = fmaf(cp, dx, apx); apy = fmaf(cp, dy, apy); apz = fmaf(cp, dz, apz);
                        }
                    }
                }
            }
        }
        red[0][tid] = anx; red[1][tid] = any; red[2][tid] = anz;
        red[3][tid] = apx; red[4][tid] = apy; red[5][tid] = apz;
        __syncthreads();
        if (split == 0 && active) {
            float4 pout = pi, vout = vi, acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (walker) {
                for (int s = 1; s < nsplit; ++s) {
                    int o = s * tl + t_local;
                    anx += red[0][o]; any += red[1][o]; anz += red[2][o];
                    apx += red[3][o]; apy += red[4][o]; apz += red[5][o];
                }
                float nx = sp.g[0] - anx, ny = sp.g[1] - any, nz = sp.g[2] - anz;   // wcsphv2.py:89-93
                if (a_np_out) {
                    a_np_out[i] = make_float4(nx, ny, nz, 0.f);
                    a_p_out[i] = make_float4(apx, apy, apz, 0.f);
                }
                acc.x = nx + apx; acc.y = ny + apy; acc.z = nz + apz;              // wcsphv2.py:53
                // advert (wcsphv2.py:98-99): v += dt a ; x += dt v
                vout.x = vi.x + sp.dt * acc.x; vout.y = vi.y + sp.dt * acc.y; vout.z = vi.z + sp.dt * acc.z;
                float px = pi.x + sp.dt * vout.x, py = pi.y + sp.dt * vout.y, pz = pi.z + sp.dt * vout.z;
                // enforce_boundary_3D_v1 (sph_basev2.py:158-189), tests use the pre-clamp position
                float cnx = 0.f, cny = 0.f, cnz = 0.f;
                pout.x = px; pout.y = py; pout.z = pz;
                if (px > sp.wall_hi[0]) { cnx += 1.f; pout.x = sp.wall_hi[0]; }
                if (px <= sp.pad)       { cnx -= 1.f; pout.x = sp.pad; }
                if (py > sp.wall_hi[1]) { cny += 1.f; pout.y = sp.wall_hi[1]; }
                if (py <= sp.pad)       { cny -= 1.f; pout.y = sp.pad; }
                if (pz > sp.wall_hi[2]) { cnz += 1.f; pout.z = sp.wall_hi[2]; }
                if (pz <= sp.pad)       { cnz -= 1.f; pout.z = sp.pad; }
                float len = sqrtf(cnx * cnx + cny * cny + cnz * cnz);
                if (len > 1e-6f) {
                    float ux = cnx / len, uy = cny / len, uz = cnz / len;
                    float sdot = 1.5f * (vout.x * ux + vout.y * uy + vout.z * uz);   // :151-156
                    vout.x -= sdot * ux; vout.y -= sdot * uy; vout.z -= sdot * uz;
                }
            } else if (a_np_out) {
                a_np_out[i] = acc;
                a_p_out[i] = acc;
            }
            Pout[i] = pout;
            Vout[i] = vout;
            Qout[i] = make_float4(di.z, di.w, qi.z, qi.w);     // clamped rho, p, material, orig id
            dvel[i] = acc;
        }
        __syncthreads();
    }
}

// =======================================================================================
// Two-phase neighbour walks (default).  At the reference's spacing only ~13 % of the 1728
// candidates of a 27-cell walk pass the cutoff, so a single loop runs the expensive pair body
// at ~13 % lane utilisation.  Here every thread first FILTERS its candidates (exact cutoff
// test only) and appends the survivors to a private list in shared memory, laid out
// [slot][thread] so that appends and reads are bank-conflict free; whenever a list could
// overflow, and at the end of every tile, the warp drains its lists running the pair body on
// (nearly) full warps.
// =======================================================================================
constexpr int LCAP = 64;               // pending-list slots per thread
constexpr int CHUNK = 32;              // candidates filtered between drain checks
constexpr float FAR = 1e18f;           // padding candidates: never within the cutoff

constexpr int D2_TCAP = 2048;
constexpr size_t D2_SMEM = (size_t)D2_TCAP * sizeof(float4) + (size_t)LCAP * NB_THREADS * sizeof(unsigned short);

__global__ void __launch_bounds__(NB_THREADS, 3)
k_density2(SimParams sp, const int* __restrict__ cell_end, const float4* __restrict__ P,
           float4* __restrict__ V, const float4* __restrict__ Q, float4* __restrict__ D,
           float* __restrict__ S, int* __restrict__ ncount) {
    extern __shared__ float4 dyn_smem[];
    float4* tile = dyn_smem;                                           // {x,y,z,material}
    unsigned short* L = reinterpret_cast<unsigned short*>(dyn_smem + D2_TCAP);
    __shared__ CellRanges R;
    __shared__ float red_w[NB_THREADS];
    __shared__ float red_b[NB_THREADS];
    __shared__ int red_c[NB_THREADS];

    const int c = blockIdx.x;
    const int tb = cell_start(cell_end, c), te = cell_end[c];
    if (te <= tb) return;
    compute_cell_ranges(sp, cell_end, c, R);
    const int total = R.off[9];
    const int tid = threadIdx.x;
    const int nT = te - tb;
    const int tl = nT <= 32 ? 32 : 64;
    const int nsplit = NB_THREADS / tl;
    const int t_local = tid % tl, split = tid / tl;
    const bool akinci = sp.volume_mode == 1;
    const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
    unsigned short* myL = L + tid;

    for (int pass = 0; pass < nT; pass += tl) {
        const int i = tb + pass + t_local;
        const bool active = i < te;
        float4 pi = active ? P[i] : make_float4(-FAR, -FAR, -FAR, 0.f);
        int mat_i = active ? __float_as_int(Q[i].z) : MAT_FLUID;
        int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        float wsum = 0.f, wbsum = 0.f;
        int cnt = 0;
        for (int tile0 = 0; tile0 < total; tile0 += D2_TCAP) {
            const int tile_n = min(D2_TCAP, total - tile0);
            const int tile_pad = (tile_n + CHUNK - 1) & ~(CHUNK - 1);
            __syncthreads();
            for (int e = tid; e < tile_pad; e += NB_THREADS) {
                float4 p = make_float4(FAR, FAR, FAR, 0.f);
                if (e < tile_n) {
                    int g = tile_to_global(R, tile0 + e);
                    p = P[g];
                    p.w = akinci ? Q[g].z : 0.f;
                }
                tile[e] = p;
            }
            __syncthreads();
            const int self_t = self_e - tile0;
            int pend = 0;
            for (int cb = split * CHUNK; cb < tile_pad; cb += nsplit * CHUNK) {
#pragma unroll 8
                for (int k = 0; k < CHUNK; ++k) {
                    float4 cj = tile[cb + k];
                    float dx = pi.x - cj.x, dy = pi.y - cj.y, dz = pi.z - cj.z;
                    float d2 = sp.dim == 3 ? dist2_exact(dx, dy, dz) : dist2_exact2(dx, dy);
                    if (d2 < sp.d2_cut) { myL[pend * NB_THREADS] = (unsigned short)(cb + k); ++pend; }
                }
                const bool last = cb + nsplit * CHUNK >= tile_pad;
                if (last || __any_sync(0xffffffffu, pend > LCAP - CHUNK)) {
                    for (int k = 0; k < pend; ++k) {
                        int e = myL[k * NB_THREADS];
                        if (e == self_t) continue;
                        float4 cj = tile[e];
                        float dx = pi.x - cj.x, dy = pi.y - cj.y, dz = pi.z - cj.z;
                        float d2 = sp.dim == 3 ? dist2_exact(dx, dy, dz) : dist2_exact2(dx, dy);
                        float r = d2 * rsqrtf(fmaxf(d2, 1e-30f));
                        float w = spline_w(r * sp.inv_h);
                        cnt++;
                        wsum += w;
                        if (__float_as_int(cj.w) == MAT_BOUNDARY) wbsum += w;
                    }
                    pend = 0;
                }
            }
        }
        red_w[tid] = wsum;
        red_b[tid] = wbsum;
        red_c[tid] = cnt;
        __syncthreads();
        if (split == 0 && active) {
            for (int s = 1; s < nsplit; ++s) {
                wsum += red_w[s * tl + t_local];
                wbsum += red_b[s * tl + t_local];
                cnt += red_c[s * tl + t_local];
            }
            float4 qi = Q[i];
            float rho_raw;
            float s_i = 0.f;
            if (mat_i == MAT_FLUID) {
                float self = pi.w * sp.k_w;                 // mass_i * W(0)
                s_i = pi.w * (sp.k_w * wsum);               // sum_j mass_i W(r_ij)   (Q2)
                rho_raw = sp.density_mode == 1 ? self + s_i : self;
            } else {
                rho_raw = qi.x;
                float delta = sp.k_w + (akinci ? sp.k_w * wbsum : 0.f);   // sph_basev2.py:195-201
                float4 vi = V[i];
                vi.w = 1.0f / delta;
                V[i] = vi;
            }
            float rho_c = fmaxf(rho_raw, sp.rho0);                                     // :46
            float pr = sp.stiffness * (eos_pow(rho_c / sp.rho0, sp.exponent, sp.int_exponent) - 1.0f);  // :47
            D[i] = make_float4(rho_raw, pr / (rho_c * rho_c), rho_c, pr);
            S[i] = s_i;
            ncount[i] = cnt;
        }
        __syncthreads();
    }
}

// tile record of the force walk: 40 B per candidate
//   tP = {x, y, z, psi}   psi = +mass_j for a fluid neighbour, -volume_j for a boundary one
//   tV = {vx, vy, vz, rho_raw}
//   tR = p_j / rho_c_j^2
constexpr int F2_TCAP = 1792;
constexpr size_t F2_SMEM = (size_t)F2_TCAP * (2 * sizeof(float4) + sizeof(float)) +
                           (size_t)LCAP * NB_THREADS * sizeof(unsigned short);

__global__ void __launch_bounds__(NB_THREADS, 2)
k_force2(SimParams sp, const int* __restrict__ cell_end, const float4* __restrict__ Pin,
         const float4* __restrict__ Vin, const float4* __restrict__ Qin,
         const float4* __restrict__ D, float4* __restrict__ Pout, float4* __restrict__ Vout,
         float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
         float4* __restrict__ a_p_out) {
    extern __shared__ float4 dyn_smem[];
    float4* tP = dyn_smem;
    float4* tV = dyn_smem + F2_TCAP;
    float* tR = reinterpret_cast<float*>(dyn_smem + 2 * F2_TCAP);
    unsigned short* L = reinterpret_cast<unsigned short*>(tR + F2_TCAP);
    __shared__ CellRanges R;
    __shared__ float red[6][NB_THREADS];

    const int c = blockIdx.x;
    const int tb = cell_start(cell_end, c), te = cell_end[c];
    if (te <= tb) return;
    compute_cell_ranges(sp, cell_end, c, R);
    const int total = R.off[9];
    const int tid = threadIdx.x;
    const int nT = te - tb;
    const int tl = nT <= 32 ? 32 : 64;
    const int nsplit = NB_THREADS / tl;
    const int t_local = tid % tl, split = tid / tl;
    const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
    unsigned short* myL = L + tid;

    for (int pass = 0; pass < nT; pass += tl) {
        const int i = tb + pass + t_local;
        const bool active = i < te;
        float4 pi = active ? Pin[i] : make_float4(-FAR, -FAR, -FAR, 1.f);
        float4 vi = active ? Vin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 di = active ? D[i] : make_float4(1.f, 0.f, 1.f, 0.f);
        float4 qi = active ? Qin[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int mat_i = __float_as_int(qi.z);
        const bool walker = active && mat_i == MAT_FLUID && i >= sp.owned_lo && i < sp.owned_hi;
        const float xi = walker ? pi.x : -FAR, yi = walker ? pi.y : -FAR, zi = walker ? pi.z : -FAR;
        int self_e = (active && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
        const float coh_i = 0.01f / pi.w;                         // wcsphv2.py:64
        const float rho_i = di.x, pr_i = di.y;
        const float nub_i = sp.visc_bound_c / (2.0f * rho_i);     // wcsphv2.py:76
        float anx = 0.f, any = 0.f, anz = 0.f;
        float apx = 0.f, apy = 0.f, apz = 0.f;
        for (int tile0 = 0; tile0 < total; tile0 += F2_TCAP) {
            const int tile_n = min(F2_TCAP, total - tile0);
            const int tile_pad = (tile_n + CHUNK - 1) & ~(CHUNK - 1);
            __syncthreads();
            for (int e = tid; e < tile_pad; e += NB_THREADS) {
                float4 p = make_float4(FAR, FAR, FAR, 0.f), v = make_float4(0.f, 0.f, 0.f, 1.f);
                float pr = 0.f;
                if (e < tile_n) {
                    int g = tile_to_global(R, tile0 + e);
                    p = Pin[g];
                    v = Vin[g];
                    float4 d = D[g];
                    if (__float_as_int(Qin[g].z) != MAT_FLUID) p.w = -v.w;   // psi = -volume
                    v.w = d.x;
                    pr = d.y;
                }
                tP[e] = p; tV[e] = v; tR[e] = pr;
            }
            __syncthreads();
            const int self_t = self_e - tile0;
            int pend = 0;
            for (int cb = split * CHUNK; cb < tile_pad; cb += nsplit * CHUNK) {
#pragma unroll 8
                for (int k = 0; k < CHUNK; ++k) {
                    float4 pj = tP[cb + k];
                    float dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
                    float d2 = dist2_exact(dx, dy, dz);
                    if (d2 < sp.d2_cut) { myL[pend * NB_THREADS] = (unsigned short)(cb + k); ++pend; }
                }
                const bool last = cb + nsplit * CHUNK >= tile_pad;
                if (last || __any_sync(0xffffffffu, pend > LCAP - CHUNK)) {
                    for (int k = 0; k < pend; ++k) {
                        int e = myL[k * NB_THREADS];
                        if (e == self_t) continue;
                        float4 pj = tP[e];
                        float4 vj = tV[e];
                        float prj = tR[e];
                        float dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
                        float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                        float r = d2 * rinv;
                        float q = r * sp.inv_h;
                        // gradW = k_dw dw(q) x_ij / (r h); zero for r <= 1e-5 (sph_basev2.py:53)
                        float gfac = r > 1e-5f ? sp.k_dw * spline_dw(q) * rinv * sp.inv_h : 0.f;
                        float dot = (vi.x - vj.x) * dx + (vi.y - vj.y) * dy + (vi.z - vj.z) * dz;
                        float mn = fminf(dot, 0.f) * fast_rcp(d2 + sp.eps_h2);
                        float cn, cp;
                        if (pj.w > 0.f) {                                                  // fluid j
                            float w = sp.k_w * spline_w(q);
                            float nu = sp.visc_fluid_c * fast_rcp(rho_i + vj.w);          // :69
                            cn = pj.w * (coh_i * w - nu * mn * gfac);                      // :64 + :72-73
                            cp = -pj.w * (pr_i + prj) * gfac;                              // sph_basev2.py:71-73
                        } else {                                                           // boundary j
                            float vol = -pj.w;
                            cn = sp.ps_density0 * vol * (-nub_i * mn) * gfac;              // :78-80
                            cp = -sp.rho0 * vol * pr_i * gfac;                             // sph_basev2.py:75
                        }
                        anx = fmaf(cn, dx, anx); any = fmaf(cn, dy, any); anz = fmaf(cn, dz, anz);
                        apx = fmaf(cp, dx, apx); apy = fmaf(cp, dy, apy); apz = fmaf(cp, dz, apz);
                    }
                    pend = 0;
                }
            }
        }
        red[0][tid] = anx; red[1][tid] = any; red[2][tid] = anz;
        red[3][tid] = apx; red[4][tid] = apy; red[5][tid] = apz;
        __syncthreads();
        if (split == 0 && active) {
            float4 pout = pi, vout = vi, acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (walker) {
                for (int s = 1; s < nsplit; ++s) {
                    int o = s * tl + t_local;
                    anx += red[0][o]; any += red[1][o]; anz += red[2][o];
                    apx += red[3][o]; apy += red[4][o]; apz += red[5][o];
                }
                float nx = sp.g[0] - anx, ny = sp.g[1] - any, nz = sp.g[2] - anz;   // wcsphv2.py:89-93
                if (a_np_out) {
                    a_np_out[i] = make_float4(nx, ny, nz, 0.f);
                    a_p_out[i] = make_float4(apx, apy, apz, 0.f);
                }
                acc.x = nx + apx; acc.y = ny + apy; acc.z = nz + apz;              // wcsphv2.py:53
                vout.x = vi.x + sp.dt * acc.x; vout.y = vi.y + sp.dt * acc.y; vout.z = vi.z + sp.dt * acc.z;
                float px = pi.x + sp.dt * vout.x, py = pi.y + sp.dt * vout.y, pz = pi.z + sp.dt * vout.z;
                float cnx = 0.f, cny = 0.f, cnz = 0.f;                            // sph_basev2.py:158-189
                pout.x = px; pout.y = py; pout.z = pz;
                if (px > sp.wall_hi[0]) { cnx += 1.f; pout.x = sp.wall_hi[0]; }
                if (px <= sp.pad)       { cnx -= 1.f; pout.x = sp.pad; }
                if (py > sp.wall_hi[1]) { cny += 1.f; pout.y = sp.wall_hi[1]; }
                if (py <= sp.pad)       { cny -= 1.f; pout.y = sp.pad; }
                if (pz > sp.wall_hi[2]) { cnz += 1.f; pout.z = sp.wall_hi[2]; }
                if (pz <= sp.pad)       { cnz -= 1.f; pout.z = sp.pad; }
                float len = sqrtf(cnx * cnx + cny * cny + cnz * cnz);
                if (len > 1e-6f) {
                    float ux = cnx / len, uy = cny / len, uz = cnz / len;
                    float sdot = 1.5f * (vout.x * ux + vout.y * uy + vout.z * uz);   // :151-156
                    vout.x -= sdot * ux; vout.y -= sdot * uy; vout.z -= sdot * uz;
                }
            } else if (a_np_out) {
                a_np_out[i] = acc;
                a_p_out[i] = acc;
            }
            Pout[i] = pout;
            Vout[i] = vout;
            Qout[i] = make_float4(di.z, di.w, qi.z, qi.w);
            dvel[i] = acc;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Host<->device layout conversion (add_particles :171-204, dump/copy_to_numpy :279-307)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_particles(int n, int dim, int first, float m_V0, const float* __restrict__ pos,
                 const float* __restrict__ vel, const float* __restrict__ density,
                 const float* __restrict__ pressure, const int* __restrict__ material,
                 float4* __restrict__ P, float4* __restrict__ V, float4* __restrict__ Q) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pos[i * dim], y = pos[i * dim + 1], z = dim == 3 ? pos[i * dim + 2] : 0.f;
    float vx = vel[i * dim], vy = vel[i * dim + 1], vz = dim == 3 ? vel[i * dim + 2] : 0.f;
    float rho = density[i];
    float volume = m_V0;                         // :203
    float mass = volume * rho;                   // :204
    P[first + i] = make_float4(x, y, z, mass);
    V[first + i] = make_float4(vx, vy, vz, volume);
    Q[first + i] = make_float4(rho, pressure[i], __int_as_float(material[i]),
                               __int_as_float(first + i));
}

__global__ void __launch_bounds__(256)
k_upload_xv(int n, int dim, const float* __restrict__ pos, const float* __restrict__ vel,
            float4* __restrict__ P, float4* __restrict__ V) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = P[i], v = V[i];
    p.x = pos[i * dim]; p.y = pos[i * dim + 1]; p.z = dim == 3 ? pos[i * dim + 2] : 0.f;
    v.x = vel[i * dim]; v.y = vel[i * dim + 1]; v.z = dim == 3 ? vel[i * dim + 2] : 0.f;
    P[i] = p; V[i] = v;
}

// dst[i*ncomp + k] = word (comp0+k) of src[i]; 32-bit words, so f32 and i32 alike
__global__ void __launch_bounds__(256)
k_unpack(int n, const float4* __restrict__ src, int comp0, int ncomp, uint32_t* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = src[i];
    uint32_t w[4] = {__float_as_uint(s.x), __float_as_uint(s.y), __float_as_uint(s.z),
                     __float_as_uint(s.w)};
    for (int k = 0; k < ncomp; ++k) dst[(size_t)i * ncomp + k] = w[comp0 + k];
}

// colour is kept in insertion order and gathered through orig_id at dump time
__global__ void __launch_bounds__(256)
k_gather_color(int n, int ncomp, const float4* __restrict__ Q, const int* __restrict__ color,
               int* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int id = __float_as_int(Q[i].w);
    for (int k = 0; k < ncomp; ++k) dst[(size_t)i * ncomp + k] = color[(size_t)id * ncomp + k];
}

}  // namespace tisph

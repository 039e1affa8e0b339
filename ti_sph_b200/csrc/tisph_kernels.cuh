// tisph_kernels.cuh -- binning / scan / sort kernels and layout conversion of the WCSPH step
// (the neighbour walks are in tisph_walk.cuh, gen-1 in tisph_gen1.cuh, sharding in tisph_shard.cuh).
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

// `rank_key` is what the canonical intra-cell order sorts by: the index in the unsorted array
// (the reference's serial order) or, when Qin is given (sharded runs, where array positions are
// rank-local), the global original id of the particle.
__global__ void __launch_bounds__(256)
k_place(int n, const int* __restrict__ keys, const int* __restrict__ arrival,
        const int* __restrict__ cell_end, int* __restrict__ ids, const float4* __restrict__ Qin,
        int* __restrict__ rank_key) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keys[i];
    int slot = cell_start(cell_end, key) + arrival[i];
    ids[slot] = i;
    rank_key[slot] = Qin ? __float_as_int(Qin[i].w) : i;
}

__global__ void __launch_bounds__(256)
k_reorder(int n, const int* __restrict__ keys, const int* __restrict__ ids,
          const int* __restrict__ rank_key, const int* __restrict__ cell_end, const float4* __restrict__ Pin,
          const float4* __restrict__ Vin, const float4* __restrict__ Qin,
          float4* __restrict__ Pout, float4* __restrict__ Vout, float4* __restrict__ Qout,
          int* __restrict__ keys_sorted, int* __restrict__ new_index) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int id = ids[s];
    int key = keys[id];
    int b = cell_start(cell_end, key), e = cell_end[key];
    int cnt = 0;
    const int mine = rank_key[s];
    for (int t = b; t < e; ++t) cnt += (rank_key[t] < mine) ? 1 : 0;
    int dst = b + cnt;
    Pout[dst] = Pin[id];
    Vout[dst] = Vin[id];
    Qout[dst] = Qin[id];
    keys_sorted[dst] = key;
    if (new_index) new_index[id] = dst;           // paritcle_index_temp (diagnostics)
}

// ---------------------------------------------------------------------------------------
// Geometry shared by the neighbour walks (tisph_walk.cuh).  The 27 neighbour cells of a cell are 9
// contiguous ranges of the sorted arrays (z is the fastest key digit, so cells (x,y,cz-1..cz+1)
// are adjacent).  Candidate range of cell c is [cell_end[max(0,c-1)], cell_end[c])
// (partice_systemv4.py:343), which makes cell 0 invisible as a neighbour (reference quirk,
// reproduced).  Cells outside the grid are empty (the reference reads out of bounds there).
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

// The same for a thread that visits its candidates in ASCENDING order (the staging loops): the range of a
// candidate is found from the previous one's -- k only moves forward, one shared-memory compare per call
// instead of eight.  k = 0 at the start of an item.  (R.off is non-decreasing, so "the last t with
// off[t] <= e" is what tile_to_global counts; empty ranges are stepped over.)
__device__ __forceinline__ int tile_to_global_fwd(const CellRanges& R, int e, int& k) {
    while (k < 8 && e >= R.off[k + 1]) ++k;
    return R.gb[k] + (e - R.off[k]);
}

// ---------------------------------------------------------------------------------------
// Host<->device layout conversion (add_particles :171-204, dump/copy_to_numpy :279-307)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_particles(int n, int dim, int first, int id_first, float m_V0, const float* __restrict__ pos,
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
                               __int_as_float(id_first + i));
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

// dst[i*ncomp + k] = word (comp0+k) of src[i]; 32-bit words, so f32 and i32 alike.  floor_x > -inf: the x
// component is raised to floor_x first (the clamped density max(rho, rho0) of wcsphv2.py:46)
__global__ void __launch_bounds__(256)
k_unpack(int n, const float4* __restrict__ src, int comp0, int ncomp, uint32_t* __restrict__ dst, float floor_x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = src[i];
    s.x = fmaxf(s.x, floor_x);
    uint32_t w[4] = {__float_as_uint(s.x), __float_as_uint(s.y), __float_as_uint(s.z),
                     __float_as_uint(s.w)};
    for (int k = 0; k < ncomp; ++k) dst[(size_t)i * ncomp + k] = w[comp0 + k];
}

// dump() in one launch: x | v | material | colour (through orig_id) | orig_id of n particles, each in the
// reference's layout, into one staging block (any destination may be null)
__global__ void __launch_bounds__(256)
k_dump_pack(int n, int dim, int ncolor, const float4* __restrict__ P, const float4* __restrict__ V,
            const float4* __restrict__ Q, const int* __restrict__ color, float* __restrict__ out_x,
            float* __restrict__ out_v, int* __restrict__ out_mat, int* __restrict__ out_col,
            int* __restrict__ out_id) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = P[i], v = V[i], q = Q[i];
    const float pw[3] = {p.x, p.y, p.z}, vw[3] = {v.x, v.y, v.z};
    for (int k = 0; k < dim; ++k) {
        if (out_x) out_x[(size_t)i * dim + k] = pw[k];
        if (out_v) out_v[(size_t)i * dim + k] = vw[k];
    }
    if (out_mat) out_mat[i] = __float_as_int(q.z);
    const int id = __float_as_int(q.w);
    if (out_id) out_id[i] = id;
    if (out_col)
        for (int k = 0; k < ncolor; ++k) out_col[(size_t)i * ncolor + k] = color[(size_t)id * ncolor + k];
}

// max |v|^2 over the fluid particles, as float bits (non-negative floats order like unsigned ints):
// input of the optional CFL time step (an extension; the reference's dt is fixed, sph_basev2.py:14-15)
__global__ void __launch_bounds__(256)
k_vmax(int n, const float4* __restrict__ V, const float4* __restrict__ Q, unsigned int* __restrict__ out) {
    unsigned int m = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (__float_as_int(Q[i].z) != MAT_FLUID) continue;
        float4 v = V[i];
        m = max(m, __float_as_uint(v.x * v.x + v.y * v.y + v.z * v.z));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0u) atomicMax(out, m);
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

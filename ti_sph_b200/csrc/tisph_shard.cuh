// tisph_shard.cuh -- spatial-slab sharding of the step across GPUs (one process per GPU).
//
// The reference is single-device (no distributed code).  The cell key is x-major
// (key = cx*ny*nz + cy*nz + cz, partice_systemv4.py:98-100), so a slab of cell rows (cx, cy) in
// [row_lo, row_hi), row = cx*ny + cy, is one contiguous range of the sorted arrays.  A rank
//   1. packs, from the particles it advanced in the previous step, those that now lie within
//      `ghost` cell layers of a neighbour's rows or inside them (migrants) into one message per neighbour,
//   2. appends what the neighbours sent to its own slice and sorts everything together,
//   3. runs the density walk on its rows plus one layer around them and the force walk on its own rows.
// Ownership is decided by position alone after the sort (row in [row_lo, row_hi)); a migrant
// stays in the sender's arrays for one more step, where it is a ghost.
#pragma once
#include "tisph_device.cuh"

namespace tisph {

constexpr int SHARD_REC_F4 = 3;     // one record = {P, V, Q} = 48 bytes

struct ShardCounters {
    int n_left, n_right;    // records packed for the left / right neighbour
    int overflow;           // records that did not fit the message buffers
    int lost;               // particles that jumped over the whole neighbouring slab in one step
};

// sorted-index range of the particles this rank owns: written after every sort
__global__ void k_owned_range(const int* __restrict__ cell_end, int ncell, int own_key_lo, int own_key_hi,
                              int* __restrict__ range) {
    int lo = own_key_lo > 0 ? cell_end[min(own_key_lo, ncell) - 1] : 0;
    int hi = own_key_hi > 0 ? cell_end[min(own_key_hi, ncell) - 1] : 0;
    range[0] = lo;
    range[1] = hi;
}

// Pack the messages.  `range` = {first, end} of the owned slice inside P/V/Q (device memory, so
// that no host round trip is needed between the step and the pack).
// Slabs are cut at (x-plane, y-row) granularity: a rank owns the cell ROWS r = cx * gy + cy in [row_lo, row_hi) --
// still one contiguous key range.  The left neighbour owns rows below row_lo: it needs every particle of mine that
// has one of ITS cells within `ghost` cell layers (in x and y), i.e. whose lowest such row
// max(cx - ghost, 0) * gy + max(cy - ghost, 0) is below row_lo; migrants (own row below row_lo) are among them.
__global__ void __launch_bounds__(256)
k_shard_pack(SimParams sp, int n_upper, const int* __restrict__ range, int row_lo, int row_hi,
             int ghost, int left_row_lo, int right_row_hi, int cap_records,
             const float4* __restrict__ P, const float4* __restrict__ V, const float4* __restrict__ Q,
             float4* __restrict__ send_left, float4* __restrict__ send_right,
             ShardCounters* __restrict__ ctr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int first = range[0], end = range[1];
    int i = first + t;
    const bool has_left = left_row_lo >= 0, has_right = right_row_hi >= 0;
    bool valid = t < n_upper && i < end;
    bool to_left = false, to_right = false;
    float4 p, v, q;
    if (valid) {
        p = P[i];
        const int cx = cell_coord(p.x, sp.h), cy = cell_coord(p.y, sp.h);
        const int row = cx * sp.gy + cy;
        const int lowest = max(cx - ghost, 0) * sp.gy + max(cy - ghost, 0);
        const int highest = min(cx + ghost, sp.gx - 1) * sp.gy + min(cy + ghost, sp.gy - 1);
        to_left = has_left && lowest < row_lo;
        to_right = has_right && highest >= row_hi;
        // a migrant may land anywhere inside the neighbour's slab [left_row_lo, row_lo) / [row_hi, right_row_hi);
        // beyond it the neighbour would not own it either
        if ((has_left && row < left_row_lo) || (has_right && row >= right_row_hi)) atomicAdd(&ctr->lost, 1);
        if (to_left || to_right) { v = V[i]; q = Q[i]; }
    }
    // warp-aggregated slot reservation
    unsigned lane = threadIdx.x & 31;
    unsigned ml = __ballot_sync(0xffffffffu, to_left), mr = __ballot_sync(0xffffffffu, to_right);
    int bl = 0, br = 0;
    if (lane == 0) {
        if (ml) bl = atomicAdd(&ctr->n_left, __popc(ml));
        if (mr) br = atomicAdd(&ctr->n_right, __popc(mr));
    }
    bl = __shfl_sync(0xffffffffu, bl, 0);
    br = __shfl_sync(0xffffffffu, br, 0);
    if (to_left) {
        int s = bl + __popc(ml & ((1u << lane) - 1u));
        if (s < cap_records) { send_left[3 * s] = p; send_left[3 * s + 1] = v; send_left[3 * s + 2] = q; }
        else atomicAdd(&ctr->overflow, 1);
    }
    if (to_right) {
        int s = br + __popc(mr & ((1u << lane) - 1u));
        if (s < cap_records) { send_right[3 * s] = p; send_right[3 * s + 1] = v; send_right[3 * s + 2] = q; }
        else atomicAdd(&ctr->overflow, 1);
    }
}

// Append received records behind the owned slice.
__global__ void __launch_bounds__(256)
k_shard_append(int n, const float4* __restrict__ recv, int dst0, float4* __restrict__ P,
               float4* __restrict__ V, float4* __restrict__ Q) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    P[dst0 + t] = recv[3 * t];
    V[dst0 + t] = recv[3 * t + 1];
    Q[dst0 + t] = recv[3 * t + 2];
}

}  // namespace tisph

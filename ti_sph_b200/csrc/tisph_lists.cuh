// tisph_lists.cuh -- the two neighbour walks of a WCSPH step on sm_100a, list path (the kernels
// every benchmark runs; tisph_walk.cuh holds what they share with the fallback kernels).
//
// Measured on B200 (scripts/ubench_g8*.cu), and what this file is built on:
//   * a packed f32x2 instruction (FADD2/FMUL2/FFMA2) occupies the FMA pipe for two cycles: packing
//     saves issue slots, not arithmetic throughput;
//   * LDS.128 with 32 distinct addresses costs 8 cycles per warp however the banks fall (12.8 for
//     random 16-byte gathers); LDS.64 costs 2 when each half-warp hits 16 distinct 8-byte bank
//     pairs, LDS.32 1.2-1.4 when conflict-free.  Gathers therefore use 8-byte records.
//
// Thread arrangement.  A work item is up to 64 target particles of one grid cell (k_items); the
// CTA's 256 threads walk it in passes of 32 targets x 8 lanes.  Lane j of a target's group owns the
// candidates whose tile slot s has s mod 8 == j; with m = s >> 3 (the slot's ROW, 0..253) a
// candidate is named by one byte.  Rows alternate between two PARITY streams (m even / m odd),
// and every lane consumes one entry of each stream per iteration, even groups in the order
// (even, odd), odd groups (odd, even).  The 16 lanes of an LDS.64 phase (one even and one odd group)
// then read 8-byte slots s = 8 m + j with s mod 16 = j (+8 for odd m): 16 distinct bank pairs,
// no conflict, for any neighbour pattern.
//
// Tile.  The 27 neighbour cells of a cell are 9 contiguous ranges of the sorted arrays (z is the
// fastest key digit); all of them are staged into shared memory once per item, candidate e in slot
// cand_to_slot(e) (one tile of LT_CAP = 2032 candidates: 27 cells x 64 at the reference spacing is 1728, and the
// lattice aliasing of a moving block reaches 2028; anything larger goes to the fallback kernels).
//
// Walk 1 (k_density_list): FILTER -- lane j tests its candidates two rows at a time (rows 2k and
// 2k+1 are the two halves of a packed f32x2 value: FADD2/FMUL2/FFMA2, loads shared by the four
// groups of a warp) against a cutoff widened by 1e-6 and pushes the survivors' row bytes to its two
// pending streams in shared memory; DRAIN -- one (even, odd) pair of entries per iteration, packed:
// exact IEEE test sqrt(d2) < h in the reference's evaluation order (bit-exact neighbour count) and
// the kernel sum.  The drained byte pairs are copied to the warp's rows of the global neighbour-list
// pool, four entries per 32-bit word, [word][lane]; an item in which one stream of one lane collects
// more than 40 neighbours (crowded cells) is left to the fallback kernels.
// Walk 2 (k_force_list): no filter -- every lane replays its byte list, gathers the neighbour's
// {x,y} {z,psi} {vx,vy} {vz,rho} p/rho^2 from 8-byte planes and evaluates cohesion, artificial
// viscosity and pressure for two neighbours at a time, branch-free; then advect + walls.
#pragma once
#include <cuda_fp16.h>

#include "tisph_walk.cuh"

namespace tisph {

constexpr int GL = 8;                            // lanes per target
constexpr int PASS_T = NB_THREADS / GL;          // targets per pass (32)
constexpr int LT_ROWS = 128;                     // tile rows of 16 candidates in the pair-packed filter arrays
constexpr int LT_SLOTS = 2048;                   // 8-byte slots per plane (rows m = 0..255 of 8 slots)
constexpr int LT_CAP = 2032;                     // candidates of one tile: rows m = 0..253
constexpr int M_DUMMY = 254;                     // rows 254 (even stream) and 255 (odd stream) are always FAR
constexpr int LCAP2 = 48;                        // pending entries per parity stream and lane (shared memory)
constexpr int LSTRIDE = 2 * LCAP2 + 4;           // bytes of a lane's pending list (25 words: odd, so lanes spread over banks)
constexpr int FCHUNK = 8;                        // row pairs filtered between drain checks
// Neighbour-list pool (global memory): rows of 32 words (128 B).  A warp (4 targets x 8 lanes) takes one
// count row plus as many list rows as its longest lane needs -- reserved after the walk, so nothing
// is padded to a worst case: ~12 rows per warp (0.4 KB per particle) at the reference spacing.

// Candidate -> slot.  While a block of particles is still lattice-like, the 16 candidates of an aligned
// run are a (y, z) plane of a cell's 4 x 4 x 4 sub-lattice, and the plain map slot = candidate would hand
// each lane (slot mod 8) and each parity stream (bit 3 of the slot) one z-column / half of the plane: the
// lists of a target's eight lanes would differ by factors.  The map below permutes every aligned run
// of 16 so that the 16 classes (slot mod 16) are diagonals of the sub-lattice, which a neighbourhood
// ball cuts evenly:  with x = bits 4-5, y = bits 2-3, z = bits 0-1 of the candidate,
//   slot = (e & ~15) | ((x - y) & 3) << 2 | ((x + y + z) & 3).
__device__ __forceinline__ int cand_to_slot(int e) {
    const int x = e >> 4, y = e >> 2, z = e;
    return (e & ~15) | (((x - y) & 3) << 2) | ((x + y + z) & 3);
}
__device__ __forceinline__ int slot_to_cand(int s) {
    const int x = s >> 4, chi = s >> 2, clo = s;
    const int y = (x - chi) & 3;
    return (s & ~15) | (y << 2) | ((clo - x - y) & 3);
}

__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// byte K of a word, zero-extended (one PRMT)
template <int K>
__device__ __forceinline__ uint32_t byte_of(uint32_t w) {
    uint32_t v;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(w), "n"(0x4440 + K));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// =======================================================================================
// Walk 1: boundary volume, density summation, clamp, Tait EOS; builds the neighbour lists
//
// FILTER tile, half precision.  The filter only has to find a SUPERSET of the neighbours (the drain applies
// the exact IEEE test), so it runs on binary16 copies of the positions, u = (x - cell centre) / h rounded to
// half (|u| <= 1.5 for every candidate, <= 0.5 for a target), in the form
//     n_j - 2 u_i.u_j  <  T_i         n_j = |u_j|^2 of the ROUNDED u_j, rounded to half;   T_i = HC2 - |u_i|^2
// as three chained HFMA2 (z, y, x) and one HSETP2 per two rows -- no differences to form.  Error budget FOR A PAIR
// WITHIN THE CUTOFF (only those must not be lost): the rounded positions are at most sqrt(3) (2^-13 + 2^-11) =
// 1.07e-3 further apart than the true ones (d2 < 1.00215); |u_j| < 1.87, so n_j < 4 is rounded by at most 2^-10;
// the partial sums are d2 - n_i plus at most two products of magnitude <= 1.5, i.e. below 4, 4 and 2 in
// magnitude: roundings of at most 2^-10, 2^-10, 2^-11 (the products inside an FMA are exact); T_i is rounded UP.
// Total 1.00215 + 3.4e-3 = 1.0056 < HC2 = 1.0068 = 1 + 7/1024.  A pair outside the cutoff may pass with a
// larger error (big n_j): it fails the exact test in the drain and evaluates to exact zeros in the force walk
// (~1 % extra candidates).  Padding rows carry n = 60000, idle targets T = -60000: never a survivor.
// Two rows (m = 2k, 2k+1: slots 16k+j, 16k+8+j) are the two halves of a half2; one 16-byte record per row pair k
// and lane class j:
//   T4[8 k + j] = {x(2k) x(2k+1), y(2k) y(2k+1), z(2k) z(2k+1), n(2k) n(2k+1)}
// Entries of the ODD stream hold the EVEN row of their pair (row = entry + 1): one register serves both pushes,
// and the walks add the row offset to the odd stream's base address for free.
// DRAIN planes, single precision:  GXY[s] = {x,y}, GZ[s] = z of slot s;  GM[s] = material (Akinci volumes only)
//
// A pass runs in two phases with different thread arrangements over the same (target, lane) lists:
//   FILTER  thread = (lane j = warp index, target = lane id): the 32 threads of a warp test the SAME
//           candidates against 32 targets, so the tile is read with broadcast loads;
//   DRAIN   thread = (target = tid / 8, lane j = tid % 8): the arrangement of the file header, whose
//           gathers are conflict-free.
// The pending list of (target t, lane j) is LSTRIDE bytes at word LIST_J * j + LIST_T * t: consecutive t
// (filter pushes) and the 4 x 8 (t, j) of a drain warp both fall into 32 different banks.
// =======================================================================================
constexpr int LIST_T = LSTRIDE / 4;              // 25 words
constexpr int LIST_J = PASS_T * LIST_T + 4;      // 804 words: = 4 (mod 32)
constexpr int ARENA_ROWS = 4096;                 // rows (512 KiB) a CTA takes from the list pool at a time (fewer for small pools)
constexpr int ARENA_MIN = GL * (LCAP2 / 2 + 1);  // ... and at least what one pass can need
constexpr float HC2 = 1.0f + 7.0f / 1024.0f;     // filter threshold on the half-precision d2 / h^2 (see above)
constexpr float N_NEVER = 60000.0f;              // n of padding rows, -T of idle targets (finite in half precision)
constexpr size_t DL_SMEM = (size_t)LT_ROWS * 8 * sizeof(uint4) + (size_t)LT_SLOTS * (sizeof(float2) + sizeof(float)) +
                           (size_t)GL * LIST_J * 4 + (size_t)NB_THREADS * 2;
constexpr size_t DL_SMEM_AKINCI = DL_SMEM + (size_t)LT_SLOTS * sizeof(int);

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// one row pair of the filter: s = n_j - 2 u_i.u_j for the two rows (xi, yi, zi hold -2 u_i; s < T = survivor)
__device__ __forceinline__ uint32_t filter_rows(__half2 xi, __half2 yi, __half2 zi, uint4 rec) {
    return h2_bits(__hfma2(xi, bits_h2(rec.x), __hfma2(yi, bits_h2(rec.y), __hfma2(zi, bits_h2(rec.z), bits_h2(rec.w)))));
}
// ... and the predicated pushes of FCHUNK = 8 row pairs: the EVEN row byte m0 + 2u of pair u goes to the pending
// stream of each of its two rows that survived (one asm block: the stream pointers stay in their registers)
__device__ __forceinline__ void filter_push8(uint32_t& pA, uint32_t& pB, const uint32_t (&sv)[8], uint32_t T, uint32_t m0) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b32 m;\n\t"
                 "setp.lt.f16x2 p|q, %2, %10;\n\t"
                 "@p st.shared.u8 [%0], %11;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], %11;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %3, %10;\n\tadd.u32 m, %11, 2;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %4, %10;\n\tadd.u32 m, %11, 4;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %5, %10;\n\tadd.u32 m, %11, 6;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %6, %10;\n\tadd.u32 m, %11, 8;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %7, %10;\n\tadd.u32 m, %11, 10;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %8, %10;\n\tadd.u32 m, %11, 12;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t"
                 "setp.lt.f16x2 p|q, %9, %10;\n\tadd.u32 m, %11, 14;\n\t"
                 "@p st.shared.u8 [%0], m;\n\t@p add.u32 %0, %0, 2;\n\t@q st.shared.u8 [%1], m;\n\t@q add.u32 %1, %1, 2;\n\t}"
                 : "+r"(pA), "+r"(pB)
                 : "r"(sv[0]), "r"(sv[1]), "r"(sv[2]), "r"(sv[3]), "r"(sv[4]), "r"(sv[5]), "r"(sv[6]), "r"(sv[7]), "r"(T), "r"(m0)
                 : "memory");
}
static_assert(FCHUNK == 8, "filter_push8 is written for chunks of 8 row pairs");

template <bool AKINCI>
__global__ void __launch_bounds__(NB_THREADS, 3)
k_density_list(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
               StepCounters* __restrict__ ctr, int pool_rows_cap, int arena_rows, int all_to_fallback,
               const float4* __restrict__ P, float4* __restrict__ V, const float4* __restrict__ Q,
               float4* __restrict__ D, float* __restrict__ S, int* __restrict__ ncount,
               uint32_t* __restrict__ Lg, int* __restrict__ item_row, unsigned char* __restrict__ flags,
               int* __restrict__ fb_d, int* __restrict__ fb_f) {
    extern __shared__ float4 dyn_smem[];
    uint4* T4 = reinterpret_cast<uint4*>(dyn_smem);
    float2* GXY = reinterpret_cast<float2*>(T4 + LT_ROWS * 8);
    float* GZ = reinterpret_cast<float*>(GXY + LT_SLOTS);
    unsigned char* L = reinterpret_cast<unsigned char*>(GZ + LT_SLOTS);
    unsigned short* LC = reinterpret_cast<unsigned short*>(L + GL * LIST_J * 4);     // entries per stream: nA | nB << 8
    int* GM = reinterpret_cast<int*>(LC + NB_THREADS);
    __shared__ CellRanges R2[2];                     // current item / the one being published (see tisph_walk.cuh)
    __shared__ ItemMeta M2[2];
    __shared__ int s_over, s_arena[2];               // s_arena: next free row / end of this CTA's chunk of the list pool

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_items = ctr->n_items;
    if (tid < 2) s_arena[tid] = 0;
    const float2 one2 = make_float2(sp.one, sp.one);       // see the drain: keeps ptxas from contracting the exact sum
    // ---- filter arrangement: lane jF = warp, target tF = lane
    const uint32_t fT = smem_u32(T4) + 16u * warp;
    const uint32_t fL = smem_u32(L) + 4u * (LIST_J * warp + LIST_T * lane);
    // a row-2k candidate goes to byte 2n + offA of the list, a row-2k+1 candidate to byte 2n + offB: the FIRST
    // entry of every byte pair has the parity of the target's drain group (target & 1)
    const uint32_t fA = fL + (lane & 1u), fB = fL + 1u - (lane & 1u);
    // ---- drain arrangement: target tid / 8, lane j = tid % 8
    const int j = tid & (GL - 1), tD = tid >> 3;
    const uint32_t c1 = tD & 1u;                           // parity of my group: which stream I consume first
    // plane addresses of my lane for the FIRST and the SECOND entry of a byte pair; the odd stream's entries name
    // the even row of their pair, so its base is one row further (odd stream = second entry of an even group)
    const uint32_t sG1 = smem_u32(GXY) + 8u * j + 64u * c1, sG2 = smem_u32(GXY) + 8u * j + 64u * (1u - c1);
    const uint32_t sGZ1 = smem_u32(GZ) + 4u * j + 32u * c1, sGZ2 = smem_u32(GZ) + 4u * j + 32u * (1u - c1);
    const uint32_t sGM1 = smem_u32(GM) + 4u * j + 32u * c1, sGM2 = smem_u32(GM) + 4u * j + 32u * (1u - c1);
    const uint32_t sL = smem_u32(L) + 4u * (LIST_J * j + LIST_T * tD);
    const uint32_t offA = c1, offB = 1u - c1;
    // the exported words hold TRUE rows: +1 on the bytes of the odd stream (no carry: entries <= 254)
    const uint32_t odd_plus1 = c1 ? 0x00010001u : 0x01000100u;
    // the two dummy rows are FAR in every array, for good (half: |u| = 1000 makes d2 overflow to +inf)
    if (tid < 16) { GXY[8 * M_DUMMY + tid] = make_float2(FAR, FAR); GZ[8 * M_DUMMY + tid] = FAR; }
    if (AKINCI && tid < 16) GM[8 * M_DUMMY + tid] = MAT_FLUID;

    ItemFetch nx;
    if (warp == 0) {
        nx = fetch_item(sp, cell_end, items, n_items, &ctr->work_d);
        publish_item(nx, R2[0], M2[0]);
    }
    __syncthreads();
    for (int buf = 0;; buf ^= 1) {
        const CellRanges& R = R2[buf];
        const ItemMeta& M = M2[buf];
        const int it = M.it;
        if (it >= n_items) break;
        ItemGeom G;
        item_geometry(M, R, G);
        if (warp == 0) nx = fetch_item(sp, cell_end, items, n_items, &ctr->work_d);   // the next item, behind this one's walk
        if (G.total > LT_CAP || all_to_fallback) {          // a candidate's row must fit one byte
            if (tid == 0) {
                flags[it] = 2;
                fb_d[atomicAdd(&ctr->n_fb_d, 1)] = it;
                fb_f[atomicAdd(&ctr->n_fb_f, 1)] = it;
            }
            if (warp == 0) publish_item(nx, R2[buf ^ 1], M2[buf ^ 1]);
            __syncthreads();
            continue;
        }
        const bool own = G.c >= sp.own_key_lo && G.c < sp.own_key_hi;   // ghost cells get no force walk
        const int npass = (G.nT + PASS_T - 1) / PASS_T;
        if (tid == 0) s_over = 0;
        // a ghost cell whose density does not depend on its neighbours (reference modes) skips the walk
        const int walk_total = (own || sp.ghost_walk) ? G.total : 0;
        // ---- stage the tile: rows of 16 candidates, padded with FAR to a whole filter chunk -------------
        const int nrow2 = ((walk_total + 15) / 16 + FCHUNK - 1) / FCHUNK * FCHUNK;      // row pairs (k) staged
        // centre of the target cell: origin of the half-precision coordinates
        const float ccx = ((float)(G.c / (sp.gz * sp.gy)) + 0.5f) * sp.h, ccy = ((float)((G.c / sp.gz) % sp.gy) + 0.5f) * sp.h,
                    ccz = ((float)(G.c % sp.gz) + 0.5f) * sp.h;
        // Two records (p, p + 256) per iteration: the four candidates' global rows first -- each of the two candidate
        // sequences of a thread ascends, so their ranges are found incrementally -- then all four loads, then the
        // conversions: the load latencies overlap instead of following one another.
        int ka = 0, kb = 0;
        const int n8 = nrow2 * 8;
        for (int p = tid; p < n8; p += 2 * NB_THREADS) {
            const bool two = p + NB_THREADS < n8;
            int sa[2], ga[2], gb[2];
            bool va[2], vb[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int pu = p + u * NB_THREADS;
                sa[u] = 16 * (pu >> 3) + (pu & 7);                            // slots sa, sa + 8: rows 2k and 2k+1, lane class jj
                const int ea = slot_to_cand(sa[u]), eb = slot_to_cand(sa[u] + 8);
                va[u] = (u == 0 || two) && ea < walk_total;                   // padding: never within the cutoff
                vb[u] = (u == 0 || two) && eb < walk_total;
                ga[u] = va[u] ? tile_to_global_fwd(R, ea, ka) : G.i0;         // (padding loads the first target: any valid row)
                gb[u] = vb[u] ? tile_to_global_fwd(R, eb, kb) : G.i0;
            }
            float4 a[2], b[2];
            int ma[2] = {MAT_FLUID, MAT_FLUID}, mb[2] = {MAT_FLUID, MAT_FLUID};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                a[u] = P[ga[u]];
                b[u] = P[gb[u]];
                if (AKINCI) { ma[u] = __float_as_int(Q[ga[u]].z); mb[u] = __float_as_int(Q[gb[u]].z); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int pu = p + u * NB_THREADS, sb = sa[u] + 8;
                TISPH_CHECK(!(u == 0 || two) || (sb < LT_SLOTS && pu < LT_ROWS * 8));
                const float4 far4 = make_float4(FAR, FAR, FAR, 0.f);
                const float4 pa = va[u] ? a[u] : far4, pb = vb[u] ? b[u] : far4;
                const float3 ua = va[u] ? make_float3((pa.x - ccx) * sp.inv_h, (pa.y - ccy) * sp.inv_h, (pa.z - ccz) * sp.inv_h)
                                        : make_float3(0.f, 0.f, 0.f);
                const float3 ub = vb[u] ? make_float3((pb.x - ccx) * sp.inv_h, (pb.y - ccy) * sp.inv_h, (pb.z - ccz) * sp.inv_h)
                                        : make_float3(0.f, 0.f, 0.f);
                // rounded coordinates and the norm of the ROUNDED vector (exact products in f32, one rounding to half)
                const __half2 hx = __floats2half2_rn(ua.x, ub.x), hy = __floats2half2_rn(ua.y, ub.y), hz = __floats2half2_rn(ua.z, ub.z);
                const float2 fx = __half22float2(hx), fy = __half22float2(hy), fz = __half22float2(hz);
                const float na = va[u] ? fmaf(fz.x, fz.x, fmaf(fy.x, fy.x, fx.x * fx.x)) : N_NEVER;
                const float nb = vb[u] ? fmaf(fz.y, fz.y, fmaf(fy.y, fy.y, fx.y * fx.y)) : N_NEVER;
                if (u == 0 || two) {                                          // (only the stores of the second record are conditional)
                    T4[pu] = make_uint4(h2_bits(hx), h2_bits(hy), h2_bits(hz), h2_bits(__floats2half2_rn(na, nb)));
                    GXY[sa[u]] = make_float2(pa.x, pa.y); GZ[sa[u]] = pa.z;
                    GXY[sb] = make_float2(pb.x, pb.y); GZ[sb] = pb.z;
                    if (AKINCI) { GM[sa[u]] = va[u] ? ma[u] : MAT_FLUID; GM[sb] = vb[u] ? mb[u] : MAT_FLUID; }
                }
            }
        }
        __syncthreads();
        const int self_lo = R.gb[4], self_len = R.off[5] - R.off[4];
        bool redo = false;

        for (int pass = 0; pass < npass; ++pass) {
            // the drain's target (arrangement: target tid / 8): asked for now, needed after the filter
            const int t_local = pass * PASS_T + tD;
            const int i = G.i0 + t_local;
            const bool active = t_local < G.nT;
            const float4 pi = active ? P[i] : make_float4(-FAR, -FAR, -FAR, 0.f);
            const int mat_i = active ? __float_as_int(Q[i].z) : MAT_FLUID;
            if (tid == 0 && own) {
                // list rows come from a per-CTA arena refilled here, behind the filter, so that the warps reserve
                // theirs with a shared-memory atomic (a pass needs at most 8 x (1 + LCAP2 / 2) rows)
                if (s_arena[1] - s_arena[0] < GL * (LCAP2 / 2 + 1)) {
                    const int base = atomicAdd(&ctr->pool_rows, arena_rows);
                    s_arena[0] = base;
                    s_arena[1] = base + arena_rows <= pool_rows_cap ? base + arena_rows : base;   // exhausted: an empty arena
                }
            }
            // ======== FILTER: my target against the candidates of lane `warp` ==============================
            int ovf = 0;
            {
                const int tf = pass * PASS_T + lane;
                __half2 xi = __float2half2_rn(0.f), yi = xi, zi = xi;
                uint32_t T = h2_bits(__float2half2_rn(-N_NEVER));      // idle lane: nothing survives
                if (tf < G.nT) {
                    // my target's position: it is in the tile (one of the cell's own particles), except in cell 0,
                    // which is invisible as a neighbour cell (Q3), and in ghost cells that skip the walk
                    const int itf = G.i0 + tf, st = itf - self_lo;
                    float3 pf;
                    if (walk_total > 0 && st >= 0 && st < self_len) {
                        const int s = cand_to_slot(R.off[4] + st);
                        const float2 xy = GXY[s];
                        pf = make_float3(xy.x, xy.y, GZ[s]);
                    } else {
                        const float4 q4 = P[itf];
                        pf = make_float3(q4.x, q4.y, q4.z);
                    }
                    const __half hx = __float2half_rn((pf.x - ccx) * sp.inv_h), hy = __float2half_rn((pf.y - ccy) * sp.inv_h),
                                 hz = __float2half_rn((pf.z - ccz) * sp.inv_h);
                    const float fx = __half2float(hx), fy = __half2float(hy), fz = __half2float(hz);
                    xi = __float2half2_rn(-2.0f * fx); yi = __float2half2_rn(-2.0f * fy); zi = __float2half2_rn(-2.0f * fz);   // (exact)
                    T = h2_bits(__half2half2(__float2half_ru(HC2 - fmaf(fz, fz, fmaf(fy, fy, fx * fx)))));
                }
                uint32_t pA = fA, pB = fB;                     // next free byte of the two pending streams
                for (int k0 = 0; k0 < nrow2; k0 += FCHUNK) {
                    // FCHUNK row pairs = FCHUNK records of 16 bytes; all loads first
                    uint4 rec[FCHUNK];
#pragma unroll
                    for (int u = 0; u < FCHUNK; ++u)
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rec[u].x), "=r"(rec[u].y), "=r"(rec[u].z), "=r"(rec[u].w)
                                     : "r"(fT + 128u * (uint32_t)(k0 + u)) : "memory");
                    uint32_t sv[FCHUNK];
#pragma unroll
                    for (int u = 0; u < FCHUNK; ++u) sv[u] = filter_rows(xi, yi, zi, rec[u]);
                    TISPH_CHECK(pA - fL < 2u * LCAP2 + 2u && pB - fL < 2u * LCAP2 + 2u);
                    filter_push8(pA, pB, sv, T, 2u * (uint32_t)k0);
                    // a stream that could overflow with the next chunk: the whole item goes to the fallback kernels
                    if (max(pA - fA, pB - fB) > 2u * (LCAP2 - FCHUNK)) { ovf = 1; break; }
                }
                LC[tid] = (unsigned short)(((pA - fA) >> 1) | (((pB - fB) >> 1) << 8));
            }
            if (__syncthreads_or(ovf)) { redo = true; break; }
            // ======== DRAIN: target tid / 8, lane tid % 8 ====================================================
            // the target's own slot in the tile
            const int self_t = (active && walk_total > 0 && i >= self_lo && i < self_lo + self_len) ? R.off[4] + (i - self_lo) : -1;
            const int self_s = self_t >= 0 ? cand_to_slot(self_t) : -1;
            float2 wsum2 = make_float2(0.f, 0.f);
            float wbsum = 0.f;
            int cnt = 0;
            // the filter thread of (target tD, lane j) was thread 32 j + tD
            const uint32_t nab = LC[32 * j + tD];
            uint32_t nA = nab & 0xffu, nB = nab >> 8;
            if (self_s >= 0) {
                // p_i != p_j (partice_systemv4.py:344): the target itself passed the filter (d2 = 0) and sits in the
                // list of lane (self_s & 7).  The 8 lanes of the group look for it together, each in every 8th
                // entry of that stream (independent loads); the one that finds it puts the dummy row in its place.
                const uint32_t ms = (uint32_t)self_s >> 3, js = (uint32_t)self_s & 7u;
                const uint32_t nss = LC[32 * js + tD];
                const uint32_t ns = (ms & 1u) ? nss >> 8 : nss & 0xffu;
                const uint32_t sb = smem_u32(L) + 4u * (LIST_J * js + LIST_T * tD) + ((ms & 1u) ? offB : offA);
#pragma unroll
                for (uint32_t r8 = 0; r8 < (uint32_t)LCAP2; r8 += 8u) {
                    const uint32_t k = r8 + (uint32_t)j;
                    if (k < ns && lds_u8(sb + 2u * k) == (ms & ~1u)) sts_u8(sb + 2u * k, M_DUMMY);
                }
            }
            __syncwarp();
            // The two streams are consumed in lockstep (here and in the force walk), so the longer one sets the
            // number of iterations: level them by moving tail entries across (the parity only matters for the
            // banks: an entry in the "wrong" stream costs its gathers a two-way conflict; ~1 in 10 moves).
            {
                const bool a_long = nA > nB;
                const uint32_t ns = a_long ? nA : nB, nd = a_long ? nB : nA;
                const uint32_t src = sL + (a_long ? offA : offB), dst = sL + (a_long ? offB : offA);
                const uint32_t mv = min(8u, (ns - nd) >> 1);                // (loads first: their latencies overlap)
                // even row m in the even stream = entry m - 1 in the odd stream, and back.  Only tail entries move
                // (at most half of the longer stream, which is ascending): never the entry 0 of an even stream.
                const uint32_t adj = a_long ? 0xffffffffu : 1u;
                uint32_t e[8];
#pragma unroll
                for (uint32_t u = 0; u < 8u; ++u) if (u < mv) e[u] = lds_u8(src + 2u * (ns - 1u - u));
#pragma unroll
                for (uint32_t u = 0; u < 8u; ++u) if (u < mv) sts_u8(dst + 2u * (nd + u), e[u] + adj);
                nA = a_long ? ns - mv : nd + mv;
                nB = a_long ? nd + mv : ns - mv;
            }
            // pad both streams with their dummy row to the longest list of the warp (whole words), then one
            // (first, second) pair of entries per iteration, branch-free
            const uint32_t n2 = (2u * max(nA, nB) + 2u) & ~3u;                  // 2 x pairs I hand to the force walk
            const uint32_t nmax = __reduce_max_sync(0xffffffffu, n2);
            for (uint32_t k = 2u * nA; k < nmax; k += 2u) sts_u8(sL + offA + k, M_DUMMY);
            for (uint32_t k = 2u * nB; k < nmax; k += 2u) sts_u8(sL + offB + k, M_DUMMY);       // (row M_DUMMY + 1)
            // one word = two (first, second) pairs of entries; a byte is taken out with one PRMT and scaled into its
            // plane address by one shift-add each
            auto drain2 = [&](uint32_t m1, uint32_t m2) {
                TISPH_CHECK(m1 < 256u && m2 < 256u);
                const float2 xy1 = lds_f32x2(sG1 + (m1 << 6));
                const float z1 = lds_f32(sGZ1 + (m1 << 5));
                const float2 xy2 = lds_f32x2(sG2 + (m2 << 6));
                const float z2 = lds_f32(sGZ2 + (m2 << 5));
                const float2 dx = make_float2(pi.x - xy1.x, pi.x - xy2.x);
                const float2 dy = make_float2(pi.y - xy1.y, pi.y - xy2.y);
                const float2 dz = make_float2(pi.z - z1, pi.z - z2);
                // d2 in the reference's evaluation order with every product rounded (dist2_exact, two at a
                // time).  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn, so the
                // sums are written as fma(y2, 1, x2) with a run-time 1: exact, and nothing left to contract.
                const float2 d2 = __ffma2_rn(__fmul2_rn(dz, dz), one2, __ffma2_rn(__fmul2_rn(dy, dy), one2, __fmul2_rn(dx, dx)));
                const float2 rinv = make_float2(rsqrt_approx(fmaxf(d2.x, 1e-30f)), rsqrt_approx(fmaxf(d2.y, 1e-30f)));
                // q = sqrt(d2) / h with one Newton step on top of MUFU.RSQ: the density sum feeds p = B (x^7 - 1),
                // which multiplies its relative error by 7 and more, so q is kept to ~1 ulp here:
                // r0 = d2 rinv, e = r0 rinv - 1, q = r0/h - (r0/2h) e
                const float2 r0 = __fmul2_rn(d2, rinv);
                const float2 e = __ffma2_rn(r0, rinv, make_float2(-1.0f, -1.0f));
                const float2 q0 = __fmul2_rn(r0, make_float2(sp.inv_h, sp.inv_h));
                float2 q = __ffma2_rn(__fmul2_rn(q0, make_float2(-0.5f, -0.5f)), e, q0);
                q.x = fminf(q.x, 1.0f); q.y = fminf(q.y, 1.0f);
                // cubic spline, branch-free:  W/k = 2 (1-q)^3 - 8 max(1/2 - q, 0)^3   (sph_basev2.py:19-36)
                const float2 nf = __fadd2_rn(q, make_float2(-1.0f, -1.0f));
                float2 g = make_float2(fmaxf(-0.5f - nf.x, 0.f), fmaxf(-0.5f - nf.y, 0.f));
                const float2 nf3 = __fmul2_rn(__fmul2_rn(nf, nf), nf);
                const float2 g3 = __fmul2_rn(__fmul2_rn(g, g), g);
                if (AKINCI) {
                    const float2 w2 = __ffma2_rn(g3, make_float2(-8.0f, -8.0f), __fmul2_rn(nf3, make_float2(-2.0f, -2.0f)));
                    wsum2 = __fadd2_rn(wsum2, w2);
                    wbsum += (int)lds_u32(sGM1 + (m1 << 5)) == MAT_BOUNDARY ? w2.x : 0.f;
                    wbsum += (int)lds_u32(sGM2 + (m2 << 5)) == MAT_BOUNDARY ? w2.y : 0.f;
                } else {
                    wsum2 = __ffma2_rn(nf3, make_float2(-2.0f, -2.0f), wsum2);
                    wsum2 = __ffma2_rn(g3, make_float2(-8.0f, -8.0f), wsum2);
                }
                asm("{\n\t.reg .pred p, q;\n\tsetp.lt.f32 p, %1, %3;\n\tsetp.lt.f32 q, %2, %3;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t@q add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(d2.x), "f"(d2.y), "f"(sp.d2_cut));
            };
            for (uint32_t k4 = 0; k4 < (sp.lists_only ? 0u : nmax); k4 += 4u) {
                const uint32_t w = lds_u32(sL + k4);
                TISPH_CHECK(k4 < 2u * LCAP2 + 4u);
                drain2(byte_of<0>(w), byte_of<1>(w));
                drain2(byte_of<2>(w), byte_of<3>(w));
            }
            // ---- hand my pairs to the force walk: the warp reserves a count row and the rows of its longest
            //      lane (rows of 32 words), then copies whole words of (first, second, first, second) entries
            if (own) {
                const int nw = (int)(n2 >> 2);                  // (levelled and padded above)
                const int rows = __reduce_max_sync(0xffffffffu, nw) + 1;
                int row = 0;
                if (lane == 0) {
                    row = atomicAdd(&s_arena[0], rows);
                    if (row + rows > s_arena[1]) { row = -1; s_over = 1; }     // pool exhausted: fallback force kernel
                    item_row[(2 * it + pass) * 8 + warp] = row;
                }
                row = __shfl_sync(0xffffffffu, row, 0);
                if (row >= 0) {
                    uint32_t* gp = Lg + (size_t)row * 32 + lane;
                    *gp = (uint32_t)nw;                        // count row, in words
                    for (int q4 = 0; q4 < nw; ++q4) {
                        gp += 32;
                        TISPH_CHECK(row + 1 + q4 < pool_rows_cap);
                        *gp = lds_u32(sL + 4u * q4) + odd_plus1;
                    }
                }
            }
            float wsum = wsum2.x + wsum2.y;
#pragma unroll
            for (int o = 1; o < GL; o <<= 1) {
                wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                if (AKINCI) wbsum += __shfl_xor_sync(0xffffffffu, wbsum, o);
            }
            if (j == 0 && active)
                density_epilogue(sp, i, pi.w, mat_i, wsum, wbsum, cnt, V, Q, D, S, ncount);
            __syncthreads();                                   // the lists are free for the next pass; s_over is complete
        }
        if (tid == 0) {
            if (redo) {                                        // a pending list would have overflowed: both walks fall back
                flags[it] = 2;
                fb_d[atomicAdd(&ctr->n_fb_d, 1)] = it;
                fb_f[atomicAdd(&ctr->n_fb_f, 1)] = it;
            } else {
                flags[it] = (unsigned char)s_over;
                if (s_over && own) fb_f[atomicAdd(&ctr->n_fb_f, 1)] = it;
            }
        }
        // the one barrier between two items: the next item's ranges go to the other buffer, which nobody reads now
        if (warp == 0) publish_item(nx, R2[buf ^ 1], M2[buf ^ 1]);
        __syncthreads();
    }
}

// =======================================================================================
// Walk 2: forces + advect + walls.  Tile = five planes of LT_SLOTS 8-byte (4-byte) slots:
//   P01 {x,y}   P23 {z,psi}   V01 {vx,vy}   RP {rho_raw, p/rho_c^2}   VZ vz
// every plane is a contiguous piece of one source record (P.xy, {P.z, D.z}, V.xy, D.xy, V.z), so the tile is staged
// with asynchronous global -> shared copies (cp.async: no registers in between, all of a thread's copies in flight
// at once)
//   psi = +mass_j (fluid j) / -volume_j (boundary j)
// =======================================================================================
constexpr uint32_t PLANE_B = (uint32_t)LT_SLOTS * sizeof(float2);
constexpr size_t FL_SMEM = (size_t)LT_SLOTS * (4 * sizeof(float2) + sizeof(float));

struct ForceAcc2 { float2 anx, any, anz, apx, apy, apz; };

// per-target constants of the pair evaluation.  K = -k_dw: gradW = K gfac0 x_ij with gfac0 = (dW/dq / -6k) / (r h);
// K is folded into the constants below and, for the pressure sum, applied once in the epilogue.
struct ForceConst {
    float c8;            // coh_i k_w W = c8 (g^3 + nf^3 / 4)   (c8 = -8 coh_kw ; wcsphv2.py:64)
    float cv;            // -nu_fluid K                      (wcsphv2.py:72-73)
    float cbK, pb, K;    // boundary j: rho0-scaled viscosity coefficient x K ; rho0 p_i/rho_i^2 ; K
    float nKpi;          // -K p_i / rho_i^2
};

// Branch-free evaluation of two neighbours (wcsphv2.py:56-80 ; sph_basev2.py:64-78).  The self pair and
// coincident particles give exactly zero (x_ij = 0 and gradW = 0 for r <= 1e-5, sph_basev2.py:53);
// entries outside the cutoff (filter band, list padding) have q clamped to 1, where W and gradW vanish.
//   dW/dq / (6k) = 4 max(1/2 - q, 0)^2 - (1-q)^2     (sph_basev2.py:53-60)
// SPLIT keeps the non-pressure and the pressure sums apart (diagnostics, kernel-by-kernel stepping); otherwise one
// accumulator takes  -(non-pressure) + K (pressure)  per pair: three FFMA2 less per two neighbours.
template <bool HAS_BOUNDARY, bool SPLIT>
__device__ __forceinline__ void pair_force2(const SimParams& sp, const ForceConst& C, float4 pi, float4 vi,
                                            float rho_i, float pr_i, float2 xy1, float2 zp1, float2 vxy1,
                                            float2 vzr1, float pr1, float2 xy2, float2 zp2, float2 vxy2, float2 vzr2,
                                            float pr2, ForceAcc2& A) {
    const float2 dx = make_float2(pi.x - xy1.x, pi.x - xy2.x);
    const float2 dy = make_float2(pi.y - xy1.y, pi.y - xy2.y);
    const float2 dz = make_float2(pi.z - zp1.x, pi.z - zp2.x);
    const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    const float2 rinv = make_float2(rsqrt_approx(fmaxf(d2.x, 1e-30f)), rsqrt_approx(fmaxf(d2.y, 1e-30f)));
    const float2 rih = __fmul2_rn(rinv, make_float2(sp.inv_h, sp.inv_h));           // 1 / (r h)
    float2 q = __fmul2_rn(d2, rih);
    q.x = fminf(q.x, 1.0f); q.y = fminf(q.y, 1.0f);                  // beyond the support W = gradW = 0: no mask needed
    const float2 nf = __fadd2_rn(q, make_float2(-1.0f, -1.0f));
    float2 g = __ffma2_rn(q, make_float2(-1.0f, -1.0f), make_float2(0.5f, 0.5f));
    g.x = fmaxf(g.x, 0.f); g.y = fmaxf(g.y, 0.f);
    const float2 f2 = __fmul2_rn(nf, nf), g2 = __fmul2_rn(g, g);
    const float2 ndw = __ffma2_rn(g2, make_float2(-4.0f, -4.0f), f2);                 // -dW/dq / (6k)
    float2 gfac = __fmul2_rn(ndw, rih);                                               // gradW = K gfac x_ij
    gfac.x = d2.x > 1e-10f ? gfac.x : 0.f;                                            // r > 1e-5 (sph_basev2.py:53)
    gfac.y = d2.y > 1e-10f ? gfac.y : 0.f;
    const float2 dvx = make_float2(vi.x - vxy1.x, vi.x - vxy2.x);
    const float2 dvy = make_float2(vi.y - vxy1.y, vi.y - vxy2.y);
    const float2 dvz = make_float2(vi.z - vzr1.x, vi.z - vzr2.x);
    float2 dot = __ffma2_rn(dvz, dz, __ffma2_rn(dvy, dy, __fmul2_rn(dvx, dx)));
    dot.x = fminf(dot.x, 0.f); dot.y = fminf(dot.y, 0.f);
    // min(v.x, 0) / (d2 + 0.01 h^2) / (rho_i + rho_j) with one reciprocal (wcsphv2.py:69-73)
    const float2 d2e = __fadd2_rn(d2, make_float2(sp.eps_h2, sp.eps_h2));
    const float2 rs = make_float2(rho_i + vzr1.y, rho_i + vzr2.y);
    const float2 den = __fmul2_rn(d2e, rs);
    const float2 mnr = __fmul2_rn(dot, make_float2(rcp_approx(den.x), rcp_approx(den.y)));
    const float2 nf3 = __fmul2_rn(f2, nf), g3 = __fmul2_rn(g2, g);
    if (SPLIT) {
        const float2 t = __fmul2_rn(mnr, gfac);
        const float2 ps = __fmul2_rn(__fadd2_rn(make_float2(pr1, pr2), make_float2(pr_i, pr_i)), gfac);   // sph_basev2.py:71-73 (/ K)
        const float2 cw = __fmul2_rn(__ffma2_rn(nf3, make_float2(0.25f, 0.25f), g3), make_float2(C.c8, C.c8));
        const float2 u = __ffma2_rn(t, make_float2(C.cv, C.cv), cw);                  // :64 + :72-73
        float2 cn = make_float2(zp1.y * u.x, zp2.y * u.y);                            // psi (coh_i W - nu mn gradW)
        float2 cp = make_float2(-zp1.y * ps.x, -zp2.y * ps.y);
        if (HAS_BOUNDARY) {
            const float2 mnb = __fmul2_rn(t, rs);                                     // min(v.x,0)/(d2+eps) gfac
            // boundary j: psi = -volume_j                     wcsphv2.py:78-80 ; sph_basev2.py:75
            cn.x = zp1.y > 0.f ? cn.x : C.cbK * zp1.y * mnb.x;
            cn.y = zp2.y > 0.f ? cn.y : C.cbK * zp2.y * mnb.y;
            cp.x = zp1.y > 0.f ? cp.x : C.pb * zp1.y * gfac.x;
            cp.y = zp2.y > 0.f ? cp.y : C.pb * zp2.y * gfac.y;
        }
        A.anx = __ffma2_rn(cn, dx, A.anx); A.any = __ffma2_rn(cn, dy, A.any); A.anz = __ffma2_rn(cn, dz, A.anz);
        A.apx = __ffma2_rn(cp, dx, A.apx); A.apy = __ffma2_rn(cp, dy, A.apy); A.apz = __ffma2_rn(cp, dz, A.apz);
    } else {
        // a_i += psi w x_ij,  w = -(coh_i W - nu mn gradW) - K (p_i/rho_i^2 + p_j/rho_j^2) gfac
        //                       = gfac (-K p_j/rho_j^2 - K p_i/rho_i^2 + nu K mn) - coh_i W
        float2 cwn = __fmul2_rn(__ffma2_rn(nf3, make_float2(0.25f, 0.25f), g3), make_float2(-C.c8, -C.c8));
        float2 in = __ffma2_rn(mnr, make_float2(-C.cv, -C.cv),
                               __ffma2_rn(make_float2(pr1, pr2), make_float2(-C.K, -C.K), make_float2(C.nKpi, C.nKpi)));
        if (HAS_BOUNDARY) {
            // boundary j: w = gfac (K rho0 p_i/rho_i^2 - cbK mn (rho_i + rho_j)), no cohesion   wcsphv2.py:78-80 ; sph_basev2.py:75
            const float pbK = C.pb * C.K;
            const float2 inb = __ffma2_rn(__fmul2_rn(mnr, rs), make_float2(-C.cbK, -C.cbK), make_float2(pbK, pbK));
            in.x = zp1.y > 0.f ? in.x : inb.x;   cwn.x = zp1.y > 0.f ? cwn.x : 0.f;
            in.y = zp2.y > 0.f ? in.y : inb.y;   cwn.y = zp2.y > 0.f ? cwn.y : 0.f;
        }
        const float2 w = __ffma2_rn(gfac, in, cwn);
        const float2 c = make_float2(zp1.y * w.x, zp2.y * w.y);
        A.anx = __ffma2_rn(c, dx, A.anx); A.any = __ffma2_rn(c, dy, A.any); A.anz = __ffma2_rn(c, dz, A.anz);
    }
}

template <bool HAS_BOUNDARY, bool SPLIT>
__global__ void __launch_bounds__(NB_THREADS, 3)
k_force_list(SimParams sp, const int* __restrict__ cell_end, const int2* __restrict__ items,
             StepCounters* __restrict__ ctr, const float4* __restrict__ Pin,
             const float4* __restrict__ Vin, const float4* __restrict__ Qin,
             const float4* __restrict__ D, float4* __restrict__ Pout, float4* __restrict__ Vout,
             float4* __restrict__ Qout, float4* __restrict__ dvel, float4* __restrict__ a_np_out,
             float4* __restrict__ a_p_out, const uint32_t* __restrict__ Lg,
             const int* __restrict__ item_row, const unsigned char* __restrict__ flags) {
    extern __shared__ float4 dyn_smem[];
    float2* P01 = reinterpret_cast<float2*>(dyn_smem);
    float2* P23 = P01 + LT_SLOTS;
    float2* V01 = P23 + LT_SLOTS;
    float2* RP = V01 + LT_SLOTS;
    float* VZ = reinterpret_cast<float*>(RP + LT_SLOTS);
    __shared__ CellRanges R2[2];                     // current item / the one being published
    __shared__ ItemMeta M2[2];
    const int tid = threadIdx.x;
    const int j = tid & (GL - 1);
    const int n_items = ctr->n_items;
    const uint32_t sP = smem_u32(P01) + 8u * j, sVZ = smem_u32(VZ) + 4u * j;
    const uint32_t tP01 = smem_u32(P01), tVZ = smem_u32(VZ);
    if (tid < 16) {                                         // the two dummy rows: FAR away, for good
        const int s = 8 * M_DUMMY + tid;
        P01[s] = make_float2(FAR, FAR); P23[s] = make_float2(FAR, 1.f);
        V01[s] = make_float2(0.f, 0.f); RP[s] = make_float2(1.f, 0.f);
        VZ[s] = 0.f;
    }

    ItemFetch nx;
    if (tid < 32) {
        nx = fetch_item(sp, cell_end, items, n_items, &ctr->work_f, flags);
        publish_item(nx, R2[0], M2[0]);
    }
    __syncthreads();
    for (int buf = 0;; buf ^= 1) {
        const CellRanges& R = R2[buf];
        const ItemMeta& M = M2[buf];
        const int it = M.it;
        if (it >= n_items) break;
        ItemGeom G;
        item_geometry(M, R, G);
        if (tid < 32) nx = fetch_item(sp, cell_end, items, n_items, &ctr->work_f, flags);   // the next item, behind this one's walk
        // items of k_force_fb, and ghost cells (not advanced here), are passed over
        if (M.flag || G.c < sp.own_key_lo || G.c >= sp.own_key_hi) {
            if (tid < 32) publish_item(nx, R2[buf ^ 1], M2[buf ^ 1]);
            __syncthreads();
            continue;
        }
        TISPH_CHECK(G.total <= LT_CAP);
        const int npass = (G.nT + PASS_T - 1) / PASS_T;
        // my warp's rows of the list pool, one block per pass: asked for now, needed after the staging
        const int row_p0 = item_row[(2 * it) * 8 + (tid >> 5)];
        const int row_p1 = npass > 1 ? item_row[(2 * it + 1) * 8 + (tid >> 5)] : row_p0;
        // ---- stage the tile: candidate e in slot cand_to_slot(e), six asynchronous copies each (psi comes ready-made
        //      from the density walk: D.z)
        int kr = 0;                                                              // range of my last candidate (ascending)
        for (int e = tid; e < G.total; e += NB_THREADS) {
            const int g = tile_to_global_fwd(R, e, kr);
            const uint32_t s8 = tP01 + 8u * (uint32_t)cand_to_slot(e);
            const float4 *pp = Pin + g, *vp = Vin + g, *dp = D + g;
            cp_async8(s8, pp);                                                   // P01 = {x, y}
            cp_async4(s8 + PLANE_B, &pp->z);                                     // P23 = {z, psi}
            cp_async4(s8 + PLANE_B + 4u, &dp->z);
            cp_async8(s8 + 2u * PLANE_B, vp);                                    // V01 = {vx, vy}
            cp_async8(s8 + 3u * PLANE_B, dp);                                    // RP  = {rho_raw, p / rho_c^2}
            cp_async4(tVZ + ((s8 - tP01) >> 1), &vp->z);                         // VZ  = vz
        }
        cp_async_commit();
        // the lists were written by the density walk long ago (DRAM): pull the count row and the first list rows
        // of both passes into L2 while the tile is being staged
        if (row_p0 >= 0 && row_p1 >= 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                prefetch_l2(Lg + (size_t)(row_p0 + r) * 32 + (tid & 31));
                prefetch_l2(Lg + (size_t)(row_p1 + r) * 32 + (tid & 31));
            }
        }
        // The head of a pass: the target's records, its list length and the first three list words.  None of it
        // depends on the tile, so pass 0's head is asked for BEFORE the tile barrier and pass 1's behind pass 0's pair
        // loop (before its reduction and epilogue): the latencies of these dependent loads (material and count row
        // -> list words) are off the critical path.  The two passes are two copies of the code (static registers:
        // a head that is still in flight is never moved).
        struct PassHead {
            float4 pi, vi, di, qi;
            const uint32_t* gl;
            int i, nw;
            uint32_t w0, w1, w2;
            bool active, walker;
        };
        auto load_head = [&](int pass) {
            PassHead H;
            const int t_local = pass * PASS_T + (tid >> 3);
            H.i = G.i0 + t_local;
            H.active = t_local < G.nT;
            const int row = pass ? row_p1 : row_p0;
            TISPH_CHECK(row >= 0);
            const uint32_t* gl = Lg + (size_t)row * 32 + (tid & 31);
            const int nw_row = (int)gl[0];                              // words of 4 entries (count row; written for every lane)
            H.pi = H.active ? Pin[H.i] : make_float4(-FAR, -FAR, -FAR, 1.f);
            H.vi = H.active ? Vin[H.i] : make_float4(0.f, 0.f, 0.f, 0.f);
            H.di = H.active ? D[H.i] : make_float4(1.f, 0.f, 0.f, 0.f);
            H.qi = H.active ? Qin[H.i] : make_float4(0.f, 0.f, 0.f, 0.f);
            H.walker = H.active && __float_as_int(H.qi.z) == MAT_FLUID;
            H.nw = H.walker ? nw_row : 0;
            TISPH_CHECK(H.nw >= 0 && H.nw <= LCAP2 / 2);
            gl += 32;
            // list words are fetched three ahead of their use (a word is ~170 instructions of work); the rows behind
            // the ones the staging prefetched are pulled into L2 now
            H.w0 = H.nw > 0 ? gl[0] : 0u; H.w1 = H.nw > 1 ? gl[32] : 0u; H.w2 = H.nw > 2 ? gl[64] : 0u;
            for (int r = 3; r < H.nw; ++r) prefetch_l2(gl + (size_t)r * 32);
            H.gl = gl;
            return H;
        };
        PassHead H0 = load_head(0), H1 = {};
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            if (pass >= npass) break;
            const PassHead& H = pass ? H1 : H0;
            const int i = H.i;
            const bool active = H.active, walker = H.walker;
            const float4 pi = H.pi, vi = H.vi, di = H.di, qi = H.qi;
            const float coh_kw = 0.01f / pi.w * sp.k_w;                // wcsphv2.py:64 (x the kernel normalisation)
            const float rho_i = di.x, pr_i = di.y;
            const float nub_i = sp.visc_bound_c / (2.0f * rho_i);     // wcsphv2.py:76
            ForceConst C;
            C.K = -sp.k_dw;
            C.c8 = -8.0f * coh_kw;
            C.cv = -sp.visc_fluid_c * C.K;
            C.cbK = sp.ps_density0 * nub_i * C.K;                     // wcsphv2.py:78-80
            C.pb = sp.rho0 * pr_i;                                    // sph_basev2.py:75
            C.nKpi = -C.K * pr_i;
            ForceAcc2 A;
            A.anx = A.any = A.anz = A.apx = A.apy = A.apz = make_float2(0.f, 0.f);
            const uint32_t* gl = H.gl;
            const int nw = H.nw;
            uint32_t w0 = H.w0, w1 = H.w1, w2 = H.w2;
            for (int k = 0; k < nw; ++k) {
                const uint32_t cur = w0;
                w0 = w1; w1 = w2;
                if (k + 3 < nw) w2 = gl[(size_t)(k + 3) * 32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t m1 = h ? byte_of<2>(cur) : byte_of<0>(cur), m2 = h ? byte_of<3>(cur) : byte_of<1>(cur);
                    const uint32_t a1 = sP + (m1 << 6), a2 = sP + (m2 << 6);
                    const float2 xy1 = lds_f32x2(a1), zp1 = lds_f32x2(a1 + PLANE_B);
                    const float2 vxy1 = lds_f32x2(a1 + 2u * PLANE_B), rp1 = lds_f32x2(a1 + 3u * PLANE_B);
                    const float vz1 = lds_f32(sVZ + (m1 << 5));
                    const float2 xy2 = lds_f32x2(a2), zp2 = lds_f32x2(a2 + PLANE_B);
                    const float2 vxy2 = lds_f32x2(a2 + 2u * PLANE_B), rp2 = lds_f32x2(a2 + 3u * PLANE_B);
                    const float vz2 = lds_f32(sVZ + (m2 << 5));
                    pair_force2<HAS_BOUNDARY, SPLIT>(sp, C, pi, vi, rho_i, pr_i, xy1, zp1, vxy1, make_float2(vz1, rp1.x), rp1.y,
                                                     xy2, zp2, vxy2, make_float2(vz2, rp2.x), rp2.y, A);
                }
            }
            if (pass == 0 && npass > 1) H1 = load_head(1);            // in flight during the reduction and the epilogue
            float a6[6] = {A.anx.x + A.anx.y, A.any.x + A.any.y, A.anz.x + A.anz.y,
                           A.apx.x + A.apx.y, A.apy.x + A.apy.y, A.apz.x + A.apz.y};
#pragma unroll
            for (int o = 1; o < GL; o <<= 1)
#pragma unroll
                for (int c = 0; c < (SPLIT ? 6 : 3); ++c) a6[c] += __shfl_xor_sync(0xffffffffu, a6[c], o);
            if (j == 0 && active) {
                if (SPLIT)      // a = g - (non-pressure sum) + K (pressure sum)
                    force_epilogue(sp, i, walker, pi, vi, di, qi, a6[0], a6[1], a6[2], C.K * a6[3], C.K * a6[4], C.K * a6[5],
                                   Pout, Vout, Qout, dvel, a_np_out, a_p_out);
                else            // a = g + (the one sum)
                    force_epilogue(sp, i, walker, pi, vi, di, qi, -a6[0], -a6[1], -a6[2], 0.f, 0.f, 0.f,
                                   Pout, Vout, Qout, dvel, nullptr, nullptr);
            }
        }
        // the one barrier between two items: the next item's ranges go to the other buffer, which nobody reads now
        if (tid < 32) publish_item(nx, R2[buf ^ 1], M2[buf ^ 1]);
        __syncthreads();
    }
}

}  // namespace tisph

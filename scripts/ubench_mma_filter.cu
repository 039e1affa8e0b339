// ubench_mma_filter.cu -- developer microbenchmark (not part of the product): the cutoff pre-filter of the
// density walk, half2 arithmetic (the round-2 form) against a formulation on mma.sync.m16n8k8 (f16 in, f32
// accumulate):  |u_j|^2 - 2 u_i.u_j  <  T - |u_i|^2.   Same shared-memory lists, same 32 targets x 1792
// candidates per pass, 256 threads, 3 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_mma_filter scripts/ubench_mma_filter.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int THREADS = 256;
constexpr int NPAIR = 112;                 // row pairs (16 candidates each): 1792 candidates
constexpr int LCAP2 = 48, LSTRIDE = 2 * LCAP2 + 4, LIST_T = LSTRIDE / 4, LIST_J = 32 * LIST_T + 4, GL = 8;
constexpr int H0 = 24;                     // entries of the low half of a stream (MMA form)
constexpr float HCUT = 1.0f + 7.0f / 1024.0f;
constexpr float HC2 = 1.0045f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

__device__ __forceinline__ void filter_rows(__half2 xi, __half2 yi, __half2 zi, __half2 ncut, uint32_t nx, uint32_t ny,
                                            uint32_t nz, uint32_t m, uint32_t& pA, uint32_t& pB) {
    const __half2 dx = __hadd2(xi, bits_h2(nx)), dy = __hadd2(yi, bits_h2(ny)), dz = __hadd2(zi, bits_h2(nz));
    const __half2 s = __hfma2(dz, dz, __hfma2(dy, dy, __hfma2(dx, dx, ncut)));
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\t"
                 "setp.lt.f16x2 p|q, %2, z;\n\t"
                 "@p st.shared.u8 [%0], %3;\n\t@p add.u32 %0, %0, 2;\n\t"
                 "@q st.shared.u8 [%1], %4;\n\t@q add.u32 %1, %1, 2;\n\t}"
                 : "+r"(pA), "+r"(pB) : "r"(h2_bits(s)), "r"(m), "r"(m + 1u) : "memory");
}

// candidates: U[1792] float4 (u in h units, relative to the cell centre); targets: TU[32] float4
struct Smem {
    uint4 HA[NPAIR / 2 * 8];
    uint2 HB[NPAIR / 2 * 8];
    uint2 T4[NPAIR * 16];                 // [pair][class g][k] -> {word(even row), word(odd row)}, k = 0: (x,y)  1: (z,n)
    uint32_t L[GL * LIST_J];
    uint32_t LC[THREADS];
};

__device__ void stage(Smem& S, const float4* U) {
    for (int p = threadIdx.x; p < NPAIR * 8; p += THREADS) {
        const int k = p >> 3, jj = p & 7;
        const float4 a = U[16 * k + jj], b = U[16 * k + 8 + jj];
        const uint32_t hx = h2_bits(__floats2half2_rn(-a.x, -b.x)), hy = h2_bits(__floats2half2_rn(-a.y, -b.y)),
                       hz = h2_bits(__floats2half2_rn(-a.z, -b.z));
        uint32_t* ha = reinterpret_cast<uint32_t*>(S.HA + 8 * (k >> 1) + jj);
        uint32_t* hb = reinterpret_cast<uint32_t*>(S.HB + 8 * (k >> 1) + jj);
        if ((k & 1) == 0) { ha[0] = hx; ha[1] = hy; ha[2] = hz; }
        else { ha[3] = hx; hb[0] = hy; hb[1] = hz; }
        // MMA tile: quantised coordinates and the norm of the QUANTISED vector
        const __half2 axy = __floats2half2_rn(a.x, a.y), bxy = __floats2half2_rn(b.x, b.y);
        const __half az = __float2half_rn(a.z), bz = __float2half_rn(b.z);
        const float2 fa = __half22float2(axy), fb = __half22float2(bxy);
        const float faz = __half2float(az), fbz = __half2float(bz);
        const float na = fa.x * fa.x + fa.y * fa.y + faz * faz, nb = fb.x * fb.x + fb.y * fb.y + fbz * fbz;
        S.T4[(k * 8 + jj) * 2 + 0] = make_uint2(h2_bits(axy), h2_bits(bxy));
        S.T4[(k * 8 + jj) * 2 + 1] = make_uint2(h2_bits(__halves2half2(az, __float2half_rn(na))),
                                                 h2_bits(__halves2half2(bz, __float2half_rn(nb))));
    }
}

__global__ void __launch_bounds__(THREADS, 3) k_half(const float4* U, const float4* TU, int reps, int* out) {
    extern __shared__ float4 dyn[];
    Smem& S = *reinterpret_cast<Smem*>(dyn);
    stage(S, U + (size_t)(blockIdx.x % 64) * NPAIR * 16);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t fHA = smem_u32(S.HA) + 16u * warp, fHB = smem_u32(S.HB) + 8u * warp;
    const uint32_t fL = smem_u32(S.L) + 4u * (LIST_J * warp + LIST_T * lane);
    const uint32_t fA = fL + (lane & 1u), fB = fL + 1u - (lane & 1u);
    const __half2 ncut = __float2half2_rn(-HCUT);
    int total = 0;
    for (int r = 0; r < reps; ++r) {
        const float4 pf = TU[(blockIdx.x * 7 + r) % 64 * 32 + lane];
        const __half2 xi = __float2half2_rn(pf.x), yi = __float2half2_rn(pf.y), zi = __float2half2_rn(pf.z);
        uint32_t pA = fA, pB = fB;
        for (int k0 = 0; k0 < NPAIR; k0 += 8) {
            uint4 ra[4];
            uint2 rb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t kk = (uint32_t)(k0 >> 1) + u;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ra[u].x), "=r"(ra[u].y), "=r"(ra[u].z), "=r"(ra[u].w)
                             : "r"(fHA + 128u * kk) : "memory");
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rb[u].x), "=r"(rb[u].y) : "r"(fHB + 64u * kk) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t m = 2u * (k0 + 2 * u);
                filter_rows(xi, yi, zi, ncut, ra[u].x, ra[u].y, ra[u].z, m, pA, pB);
                filter_rows(xi, yi, zi, ncut, ra[u].w, rb[u].x, rb[u].y, m + 2u, pA, pB);
            }
            if (max(pA - fA, pB - fB) > 2u * (LCAP2 - 8)) break;
        }
        S.LC[tid] = ((pA - fA) >> 1) | (((pB - fB) >> 1) << 8);
        __syncthreads();
        total += (S.LC[tid] & 0xff) + (S.LC[tid] >> 8);
        __syncthreads();
    }
    atomicAdd(out, total);
}

__device__ __forceinline__ void push(uint32_t& p, float d, uint32_t m, uint32_t step) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, 0f00000000;\n\t@p st.shared.u8 [%0], %2;\n\t@p add.u32 %0, %0, %3;\n\t}"
                 : "+r"(p) : "f"(d), "r"(m), "r"(step) : "memory");
}

// warp = (target octet t8 = warp & 3, row half hh = warp >> 2); thread (g = lane >> 2, k = lane & 3):
// candidate class g of the even and the odd row of a pair (A rows g, g + 8), targets 8 t8 + 2k, + 1 (D columns)
__global__ void __launch_bounds__(THREADS, 3) k_mma(const float4* U, const float4* TU, int reps, int* out) {
    extern __shared__ float4 dyn[];
    Smem& S = *reinterpret_cast<Smem*>(dyn);
    stage(S, U + (size_t)(blockIdx.x % 64) * NPAIR * 16);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3, t8 = warp & 3, hh = warp >> 2;
    const int t0 = 8 * t8 + 2 * k, t1 = t0 + 1;
    // the low half of the rows fills a stream upwards from entry 0, the high half downwards from entry LCAP2 - 1
    const uint32_t lb0 = smem_u32(S.L) + 4u * (LIST_J * g + LIST_T * t0) + (hh ? 2u * (LCAP2 - 1) : 0u);
    const uint32_t lb1 = smem_u32(S.L) + 4u * (LIST_J * g + LIST_T * t1) + (hh ? 2u * (LCAP2 - 1) : 0u);
    const uint32_t step = hh ? (uint32_t)-2 : 2u;
    const uint32_t fE0 = lb0, fO0 = lb0 + 1u, fE1 = lb1 + 1u, fO1 = lb1;     // first byte of a pair: parity of the target
    const uint32_t tA = smem_u32(S.T4) + 8u * (2u * g + (k & 1));
    const bool ld = k < 2;
    int total = 0;
    for (int r = 0; r < reps; ++r) {
        const float4* T = TU + (blockIdx.x * 7 + r) % 64 * 32;
        // B fragment: my operand target is 8 t8 + g
        const float4 pb = T[8 * t8 + g];
        const __half2 hxy = __floats2half2_rn(pb.x, pb.y);
        const __half hz = __float2half_rn(pb.z);
        const float2 fxy = __half22float2(hxy);
        const float fz = __half2float(hz);
        const float nt = fxy.x * fxy.x + fxy.y * fxy.y + fz * fz;
        uint32_t b0 = 0u;
        if (k == 0) b0 = h2_bits(__floats2half2_rn(-2.0f * fxy.x, -2.0f * fxy.y));
        if (k == 1) b0 = h2_bits(__floats2half2_rn(-2.0f * fz, 1.0f));
        // thresholds of my result targets 2k, 2k+1 of the octet: owned by lanes 4 (2k), 4 (2k + 1)
        const float c_t0 = __shfl_sync(0xffffffffu, nt, 8 * k) - HC2, c_t1 = __shfl_sync(0xffffffffu, nt, 8 * k + 4) - HC2;
        uint32_t pE0 = fE0, pO0 = fO0, pE1 = fE1, pO1 = fO1;
        const int kp_lo = hh * (NPAIR / 2), kp_hi = kp_lo + NPAIR / 2;
        for (int kp0 = kp_lo; kp0 < kp_hi; kp0 += 8) {
            uint32_t ae[8], ao[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ae[u] = 0u; ao[u] = 0u;
                if (ld) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ae[u]), "=r"(ao[u]) : "r"(tA + 128u * (kp0 + u)) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float d0, d1, d2, d3;
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
                             : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3)
                             : "r"(ae[u]), "r"(ao[u]), "r"(b0), "f"(c_t0), "f"(c_t1), "f"(c_t0), "f"(c_t1));
                const uint32_t mE = 2u * (kp0 + u), mO = mE + 1u;
                push(pE0, d0, mE, step); push(pE1, d1, mE, step); push(pO0, d2, mO, step); push(pO1, d3, mO, step);
            }
            const uint32_t fill = hh ? max(max(fE0 - pE0, fO0 - pO0), max(fE1 - pE1, fO1 - pO1))
                                     : max(max(pE0 - fE0, pO0 - fO0), max(pE1 - fE1, pO1 - fO1));
            if (__any_sync(0xffffffffu, fill > 2u * (LCAP2 - 8))) break;        // (mma.sync: the warp leaves together)
        }
        total += hh ? (int)(((fE0 - pE0) + (fO0 - pO0) + (fE1 - pE1) + (fO1 - pO1)) >> 1)
                    : (int)(((pE0 - fE0) + (pO0 - fO0) + (pE1 - fE1) + (pO1 - fO1)) >> 1);
        __syncthreads();
        __syncthreads();
    }
    atomicAdd(out, total);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    printf("start\n");
    const int NT = 64;                               // tiles / target sets
    float4* hU = (float4*)malloc(sizeof(float4) * NT * NPAIR * 16);
    float4* hT = (float4*)malloc(sizeof(float4) * NT * 32);
    srand(7);
    auto rnd = []() { return (float)rand() / (float)RAND_MAX; };
    for (int i = 0; i < NT * NPAIR * 16; ++i) hU[i] = make_float4(3.f * rnd() - 1.5f, 3.f * rnd() - 1.5f, 3.f * rnd() - 1.5f, 0.f);
    for (int i = 0; i < NT * 32; ++i) hT[i] = make_float4(rnd() - 0.5f, rnd() - 0.5f, rnd() - 0.5f, 0.f);
    float4 *U, *T;
    int* out;
    cudaMalloc(&U, sizeof(float4) * NT * NPAIR * 16); cudaMalloc(&T, sizeof(float4) * NT * 32); cudaMalloc(&out, 8);
    cudaMemcpy(U, hU, sizeof(float4) * NT * NPAIR * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(T, hT, sizeof(float4) * NT * 32, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_half, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    int occ_h = 0, occ_m = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_h, k_half, THREADS, sizeof(Smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, k_mma, THREADS, sizeof(Smem));
    printf("smem %zu B, CTAs/SM: half %d, mma %d\n", sizeof(Smem), occ_h, occ_m); fflush(stdout);
    const int reps = 2000, grid = 148 * 3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(out, 0, 8);
            cudaEventRecord(e0);
            if (v == 0) k_half<<<grid, THREADS, sizeof(Smem)>>>(U, T, reps, out);
            else k_mma<<<grid, THREADS, sizeof(Smem)>>>(U, T, reps, out);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            int h = 0;
            cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
            const double pairs = (double)grid * reps * 32.0 * NPAIR * 16.0;
            printf("%s: %s  %.3f ms  %.1f Gpairs/s  hits/pair %.4f  (%.2f us per pass per SM-resident CTA)\n", v ? "mma " : "half",
                   cudaGetErrorString(err), ms, pairs / ms * 1e-6, (double)h / pairs, ms * 1e3 / reps); fflush(stdout);
        }
    }
    return 0;
}

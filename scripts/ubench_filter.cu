// ubench_filter.cu -- developer microbenchmark (not part of the product): cost of the exact
// cutoff test of the neighbour walk on sm_100a, scalar vs packed f32x2 arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_filter scripts/ubench_filter.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int TILE = 1728;      // candidates of one 27-cell walk at the reference spacing
constexpr int THREADS = 256;

__device__ __forceinline__ float d2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}


// packed f32x2 with explicit .rn (ptxas must not contract these into fma.f32x2)
__device__ __forceinline__ unsigned long long pk(float2 a) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r) {
    float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}

// V0: float4 tile, scalar exact arithmetic, hit mask per 32 candidates
__global__ void __launch_bounds__(THREADS) v0(const float4* P, int reps, float cut, int* out) {
    __shared__ float4 tile[TILE];
    for (int e = threadIdx.x; e < TILE; e += THREADS) tile[e] = P[blockIdx.x * TILE + e];
    __syncthreads();
    float4 pi = P[blockIdx.x * TILE + threadIdx.x];
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
        for (int cb = 0; cb < TILE; cb += 32) {
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float4 c = tile[cb + k];
                float d2 = d2_exact(pi.x - c.x, pi.y - c.y, pi.z - c.z);
                if (d2 < cut) m |= 1u << k;
            }
            hits += __popc(m);
        }
        pi.x += 1e-7f;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = hits;
}

// V1: pair-SoA tile of NEGATED coordinates {-xa,-xb,-ya,-yb} + {-za,-zb}; packed exact arithmetic
__global__ void __launch_bounds__(THREADS) v1(const float4* P, int reps, float cut, int* out) {
    __shared__ float4 txy[TILE / 2];
    __shared__ float2 tz[TILE / 2];
    for (int e = threadIdx.x; e < TILE / 2; e += THREADS) {
        float4 a = P[blockIdx.x * TILE + 2 * e], b = P[blockIdx.x * TILE + 2 * e + 1];
        txy[e] = make_float4(-a.x, -b.x, -a.y, -b.y);
        tz[e] = make_float2(-a.z, -b.z);
    }
    __syncthreads();
    float4 pi = P[blockIdx.x * TILE + threadIdx.x];
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
        float2 xi = make_float2(pi.x, pi.x), yi = make_float2(pi.y, pi.y), zi = make_float2(pi.z, pi.z);
        for (int cb = 0; cb < TILE / 2; cb += 16) {
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float4 c = txy[cb + k];
                float2 cz = tz[cb + k];
                unsigned long long dx = add2(pk(xi), pk(make_float2(c.x, c.y)));
                unsigned long long dy = add2(pk(yi), pk(make_float2(c.z, c.w)));
                unsigned long long dz = add2(pk(zi), pk(cz));
                float2 s = upk(add2(add2(mul2(dx, dx), mul2(dy, dy)), mul2(dz, dz)));
                if (s.x < cut) m |= 1u << (2 * k);
                if (s.y < cut) m |= 2u << (2 * k);
            }
            hits += __popc(m);
        }
        pi.x += 1e-7f;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = hits;
}

// V2: as V1 but fused multiply-add (conservative pre-filter; exact test would follow on hits)
__global__ void __launch_bounds__(THREADS) v2(const float4* P, int reps, float cut, int* out) {
    __shared__ float4 txy[TILE / 2];
    __shared__ float2 tz[TILE / 2];
    for (int e = threadIdx.x; e < TILE / 2; e += THREADS) {
        float4 a = P[blockIdx.x * TILE + 2 * e], b = P[blockIdx.x * TILE + 2 * e + 1];
        txy[e] = make_float4(-a.x, -b.x, -a.y, -b.y);
        tz[e] = make_float2(-a.z, -b.z);
    }
    __syncthreads();
    float4 pi = P[blockIdx.x * TILE + threadIdx.x];
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
        float2 xi = make_float2(pi.x, pi.x), yi = make_float2(pi.y, pi.y), zi = make_float2(pi.z, pi.z);
        for (int cb = 0; cb < TILE / 2; cb += 16) {
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float4 c = txy[cb + k];
                float2 cz = tz[cb + k];
                float2 dx = __fadd2_rn(xi, make_float2(c.x, c.y));
                float2 dy = __fadd2_rn(yi, make_float2(c.z, c.w));
                float2 dz = __fadd2_rn(zi, cz);
                float2 s = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                if (s.x < cut) m |= 1u << (2 * k);
                if (s.y < cut) m |= 2u << (2 * k);
            }
            hits += __popc(m);
        }
        pi.x += 1e-7f;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = hits;
}

// V3: scalar with FMA (conservative pre-filter)
__global__ void __launch_bounds__(THREADS) v3(const float4* P, int reps, float cut, int* out) {
    __shared__ float4 tile[TILE];
    for (int e = threadIdx.x; e < TILE; e += THREADS) tile[e] = P[blockIdx.x * TILE + e];
    __syncthreads();
    float4 pi = P[blockIdx.x * TILE + threadIdx.x];
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
        for (int cb = 0; cb < TILE; cb += 32) {
            unsigned m = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float4 c = tile[cb + k];
                float dx = pi.x - c.x, dy = pi.y - c.y, dz = pi.z - c.z;
                float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d2 < cut) m |= 1u << k;
            }
            hits += __popc(m);
        }
        pi.x += 1e-7f;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = hits;
}

// V4: scalar exact + append of the hits to a per-thread smem list (current product scheme)
__global__ void __launch_bounds__(THREADS) v4(const float4* P, int reps, float cut, int* out) {
    __shared__ float4 tile[TILE];
    __shared__ unsigned short L[32 * THREADS];
    for (int e = threadIdx.x; e < TILE; e += THREADS) tile[e] = P[blockIdx.x * TILE + e];
    __syncthreads();
    float4 pi = P[blockIdx.x * TILE + threadIdx.x];
    int hits = 0;
    unsigned short* myL = L + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        for (int cb = 0; cb < TILE; cb += 32) {
            int pend = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float4 c = tile[cb + k];
                float d2 = d2_exact(pi.x - c.x, pi.y - c.y, pi.z - c.z);
                if (d2 < cut) { myL[pend * THREADS] = (unsigned short)(cb + k); ++pend; }
            }
            hits += pend;
        }
        pi.x += 1e-7f;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = hits + myL[0];
}

template <typename K>
static void run(const char* name, K kern, const float4* P, int* out, int grid, int reps, float cut) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<grid, THREADS>>>(P, 2, cut, out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    kern<<<grid, THREADS>>>(P, reps, cut, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    int h0 = 0;
    cudaMemcpy(&h0, out + 5, 4, cudaMemcpyDeviceToHost);
    double tests = (double)grid * THREADS * TILE * reps;
    printf("%-28s %8.3f ms  %7.2f Gtests/s  (%.2f SM-cycles per warp-candidate at 1.9 GHz)  hits[5]=%d  %s\n", name, ms,
           tests / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 * 4 / (tests / 32), h0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int grid = 148 * 8, reps = 40;
    size_t n = (size_t)grid * TILE;
    float4* h = (float4*)malloc(n * sizeof(float4));
    srand(1);
    for (size_t i = 0; i < n; ++i)
        h[i] = make_float4(0.12f * rand() / RAND_MAX, 0.12f * rand() / RAND_MAX, 0.12f * rand() / RAND_MAX, 1.f);
    float4* P; int* out;
    cudaMalloc(&P, n * sizeof(float4));
    cudaMalloc(&out, (size_t)grid * THREADS * 4);
    cudaMemcpy(P, h, n * sizeof(float4), cudaMemcpyHostToDevice);
    float cut = 0.04f * 0.04f;
    run("v0 scalar exact, mask", v0, P, out, grid, reps, cut);
    run("v1 f32x2 exact, mask", v1, P, out, grid, reps, cut);
    run("v2 f32x2 fma, mask", v2, P, out, grid, reps, cut);
    run("v3 scalar fma, mask", v3, P, out, grid, reps, cut);
    run("v4 scalar exact, smem list", v4, P, out, grid, reps, cut);
    return 0;
}

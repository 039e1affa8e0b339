// ubench_g8.cu -- developer microbenchmark (not part of the product): the hardware facts the
// round-2 neighbour walks are designed on, measured on sm_100a.
//   1. issue cost of scalar vs packed f32x2 FMA (is FFMA2 one or two FMA-pipe slots?)
//   2. the filter's instruction mix (packed subtract/FMA + funnel shift into a bit mask)
//   3. MUFU throughput
//   4. shared-memory wavefronts of 16-byte gathers: random slots vs slots whose residue mod 8 is
//      the lane's index inside its quarter-warp (then the 8 lanes of every LDS.128 phase hit 8
//      different bank groups)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_g8 scripts/ubench_g8.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int THREADS = 256;
constexpr int ITER = 4096;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r) {
    float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}

struct Res { long long cycles; long long ops; };

__global__ void __launch_bounds__(THREADS) k_ffma(float a, float b, float* out, long long* cyc) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3f + k;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(a), "f"(b));
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(THREADS) k_ffma2(float a, float b, float* out, long long* cyc) {
    unsigned long long x[8];
    const unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = pk(threadIdx.x * 1e-3f + k, k);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma2(x[k], A, B);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { float2 v = upk(x[k]); s += v.x + v.y; }
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// filter mix per candidate pair: MODE 0: 3 add2 + 3 fma2 + 2 funnel shifts; MODE 1: 1 add2 + 3 fma2 + 2 shifts
template <int MODE>
__global__ void __launch_bounds__(THREADS) k_mix(float a, float b, float* out, long long* cyc) {
    unsigned long long X = pk(a, a), Y = pk(b, b), Z = pk(a + b, a - b), C = pk(-b, -b);
    unsigned long long c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = pk(threadIdx.x * 1e-3f + k, k * 0.5f);
    unsigned m = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned long long s;
            if (MODE == 0) {
                unsigned long long dx = add2(X, c[k]), dy = add2(Y, c[k]), dz = add2(Z, c[k]);
                s = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, C)));
            } else {
                unsigned long long n = add2(c[k], C);
                s = fma2(X, c[k], fma2(Y, c[k], fma2(Z, c[k], n)));
            }
            float2 v = upk(s);
            m = __funnelshift_l(__float_as_uint(v.x), m, 1);
            m = __funnelshift_l(__float_as_uint(v.y), m, 1);
            c[k] = s;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = __uint_as_float(m);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(THREADS) k_mufu(float a, float* out, long long* cyc) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3f + k + a;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Shared-memory gathers.  2048 slots of 16 bytes.  PATTERN 0: every lane walks random slots;
// PATTERN 1: lane's slots are always == (lane & 7) mod 8; PATTERN 2: all lanes the same slot
// (broadcast); PATTERN 3: 8 consecutive slots per quarter-warp, the same for all 4 quarters (multicast).
// WIDTH 16: LDS.128, 4: LDS.32 (word index = slot, i.e. bank = slot mod 32)
template <int PATTERN, int WIDTH>
__global__ void __launch_bounds__(THREADS) k_lds(unsigned seed, float* out, long long* cyc) {
    extern __shared__ float4 tile[];
    for (int e = threadIdx.x; e < 2048; e += THREADS) tile[e] = make_float4(e, e + 1, e + 2, e + 3);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    unsigned s[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned h = (threadIdx.x * 2654435761u + k * 40503u + seed) * 2246822519u;
        h ^= h >> 15;
        if (PATTERN == 0) { s[k] = h & 2047u; d[k] = ((h >> 11) & 2047u) | 1u; }
        if (PATTERN == 1) { s[k] = ((h & 255u) << 3) | (lane & 7u); d[k] = (((h >> 11) & 255u) | 1u) << 3; }
        if (PATTERN == 2) { s[k] = (k * 37u + blockIdx.x) & 2047u; d[k] = 8u * (2 * k + 1); }
        if (PATTERN == 3) { s[k] = ((k * 37u) << 3 | (lane & 7u)) & 2047u; d[k] = 8u * (2 * k + 1); }
    }
    const unsigned base = (unsigned)__cvta_generic_to_shared(tile);
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s[k] = (s[k] + d[k]) & 2047u;
            if (WIDTH == 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + 16u * s[k]));
                acc += v.x + v.w;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * s[k]));
                acc += v;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int ctas_per_sm, double ops_per_thread_iter, int sms) {
    int grid = sms * ctas_per_sm;
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)grid * THREADS * 4);
    cudaMalloc(&cyc, (size_t)grid * 8);
    launch(grid, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch(grid, out, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc((size_t)grid * 8);
    cudaMemcpy(h, cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    // warp-instructions per SMSP: ctas_per_sm * 8 warps / 4 SMSPs * ITER * ops
    double winstr_smsp = ctas_per_sm * (THREADS / 32) / 4.0 * ITER * ops_per_thread_iter;
    printf("%-34s ctas/SM %d  %.3f ms  avg cycles %.0f  cycles per warp-op per SMSP %.3f  (per SM %.3f)  %s\n", name, ctas_per_sm,
           ms, avg, avg / winstr_smsp, avg / (winstr_smsp * 4), err == cudaSuccess ? "" : cudaGetErrorString(err));
    free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int c : {1, 2, 4}) {
        run("ffma (8 chains)", [&](int g, float* o, long long* cy) { k_ffma<<<g, THREADS>>>(1.0001f, 1e-3f, o, cy); }, c, 8, sms);
        run("ffma2 (8 chains)", [&](int g, float* o, long long* cy) { k_ffma2<<<g, THREADS>>>(1.0001f, 1e-3f, o, cy); }, c, 8, sms);
        run("filter mix 3add2+3fma2+2shf (x4)", [&](int g, float* o, long long* cy) { k_mix<0><<<g, THREADS>>>(1.0001f, 1e-3f, o, cy); }, c, 4 * 8, sms);
        run("filter mix 1add2+3fma2+2shf (x4)", [&](int g, float* o, long long* cy) { k_mix<1><<<g, THREADS>>>(1.0001f, 1e-3f, o, cy); }, c, 4 * 6, sms);
        run("mufu.rsq (8 chains)", [&](int g, float* o, long long* cy) { k_mufu<<<g, THREADS>>>(1.0001f, o, cy); }, c, 8, sms);
    }
    const int SM = 2048 * 16;
    for (int c : {2, 4}) {
        run("lds.128 random slots", [&](int g, float* o, long long* cy) { k_lds<0, 16><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
        run("lds.128 slot%8 == lane%8", [&](int g, float* o, long long* cy) { k_lds<1, 16><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
        run("lds.128 broadcast", [&](int g, float* o, long long* cy) { k_lds<2, 16><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
        run("lds.128 8-lane multicast", [&](int g, float* o, long long* cy) { k_lds<3, 16><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
        run("lds.32 random words", [&](int g, float* o, long long* cy) { k_lds<0, 4><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
        run("lds.32 word%8 == lane%8", [&](int g, float* o, long long* cy) { k_lds<1, 4><<<g, THREADS, SM>>>(7u, o, cy); }, c, 4, sms);
    }
    return 0;
}

// ubench_g8b.cu -- second developer microbenchmark: operand forms of packed FMA, ALU-pipe ops and
// which lanes share an LDS.128 phase on sm_100a.  Cycle counts come from one CTA per SM (clock64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/ubench_g8b scripts/ubench_g8b.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int ITER = 4096;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r) {
    float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a;
}

// MODE 0 ffma 3 regs; 1 ffma2 3 pairs; 2 ffma2 x*x+z; 3 fadd2; 4 fmul2; 5 ffma + shf interleaved;
// 6 lop3; 7 fmnmx; 8 fsetp+sel; 9 ffma2 + fmnmx interleaved; 10 ffma (x*x+z)
template <int MODE>
__global__ void k_alu(float a, float* out, long long* cyc) {
    float x[8], y[8], z[8];
    unsigned long long X[8], Y[8], Z[8];
    unsigned u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = threadIdx.x * 1e-3f + k; y[k] = 1.0f + a * k; z[k] = a * (k + 1);
        X[k] = pk(x[k], x[k] + 1); Y[k] = pk(y[k], y[k]); Z[k] = pk(z[k], -z[k]);
        u[k] = threadIdx.x * 77u + k;
    }
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(y[k]), "f"(z[k]));
            if (MODE == 10) asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(x[k]) : "f"(z[k]));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[k]) : "l"(Y[k]), "l"(Z[k]));
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %0, %1;" : "+l"(X[k]) : "l"(Z[k]));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[k]) : "l"(Y[k]));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(X[k]) : "l"(Y[k]));
            if (MODE == 5) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(y[k]), "f"(z[k]));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]));
            }
            if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(u[(k + 2) & 7]));
            if (MODE == 7) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(y[k]));
            if (MODE == 8) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(x[k]) : "f"(y[k]), "f"(z[k]));
            if (MODE == 9) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[k]) : "l"(Y[k]), "l"(Z[k]));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(y[k]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { float2 v = upk(X[k]); s += x[k] + v.x + v.y + __uint_as_float(u[k]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// lane -> residue maps: which 8 lanes form one LDS.128 phase?
//  0: random slots           1: residue = lane & 7          2: residue = lane >> 2
//  3: residue = (lane&3) | ((lane>>4)<<2)                   4: linear (slot = base + lane)
//  5: slot = 32 r + lane     6: residue = (lane & 1) | ((lane >> 3) << 1)     7: residue = lane & 7, LDS.64 on 8-byte slots
template <int PATTERN, int WIDTH>
__global__ void k_lds(unsigned seed, float* out, long long* cyc) {
    extern __shared__ float4 tile[];
    for (int e = threadIdx.x; e < 2048; e += blockDim.x) tile[e] = make_float4(e, e + 1, e + 2, e + 3);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    unsigned s[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned h = (threadIdx.x * 2654435761u + k * 40503u + seed) * 2246822519u;
        h ^= h >> 15;
        unsigned res = 0;
        if (PATTERN == 1 || PATTERN == 7) res = lane & 7u;
        if (PATTERN == 2) res = lane >> 2;
        if (PATTERN == 3) res = (lane & 3u) | ((lane >> 4) << 2);
        if (PATTERN == 6) res = (lane & 1u) | ((lane >> 3) << 1);
        s[k] = ((h & 255u) << 3) | res; d[k] = (((h >> 11) & 255u) | 1u) << 3;
        if (PATTERN == 0) { s[k] = h & 2047u; d[k] = ((h >> 11) & 2047u) | 1u; }
        if (PATTERN == 4) { s[k] = (k * 512u + lane) & 2047u; d[k] = 32u * (2 * k + 1); }
        if (PATTERN == 5) { s[k] = ((h & 63u) << 5) | lane; d[k] = (((h >> 11) & 63u) | 1u) << 5; }
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
            } else if (WIDTH == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + 8u * s[k]));
                acc += v.x + v.y;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * s[k]));
                acc += v;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int threads, double ops, int sms) {
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * threads * 4);
    cudaMalloc(&cyc, (size_t)sms * 8);
    for (int r = 0; r < 3; ++r) launch(sms, out, cyc);
    cudaError_t err = cudaDeviceSynchronize();
    long long* h = (long long*)malloc((size_t)sms * 8);
    cudaMemcpy(h, cyc, (size_t)sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    double per_smsp = (threads / 32) / 4.0 * ITER * ops;
    printf("%-44s %4d thr/SM  cycles/warp-op: per SMSP %.3f  per SM %.3f %s\n", name, threads, avg / per_smsp,
           avg / (per_smsp * 4), err == cudaSuccess ? "" : cudaGetErrorString(err));
    free(h); cudaFree(out); cudaFree(cyc);
}

#define ALU(MODE, NAME, OPS) for (int t : {256, 1024}) run(NAME, [&](int g, float* o, long long* cy) { k_alu<MODE><<<g, t>>>(1e-3f, o, cy); }, t, OPS, sms)
#define LDS(P, W, NAME) for (int t : {512, 1024}) run(NAME, [&](int g, float* o, long long* cy) { k_lds<P, W><<<g, t, 2048 * 16>>>(7u, o, cy); }, t, 4, sms)

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    ALU(0, "ffma x*y+z (3 regs)", 8);
    ALU(10, "ffma x*x+z", 8);
    ALU(1, "ffma2 X*Y+Z (3 pairs)", 8);
    ALU(2, "ffma2 X*X+Z", 8);
    ALU(3, "fadd2", 8);
    ALU(4, "fmul2", 8);
    ALU(5, "ffma + shf (pairs of ops)", 16);
    ALU(6, "lop3", 8);
    ALU(7, "fmnmx", 8);
    ALU(8, "fsetp+sel (pairs)", 16);
    ALU(9, "ffma2 + fmnmx (pairs)", 16);
    LDS(0, 16, "lds.128 random");
    LDS(1, 16, "lds.128 residue = lane&7");
    LDS(2, 16, "lds.128 residue = lane>>2");
    LDS(3, 16, "lds.128 residue = lane&3 | (lane>>4)<<2");
    LDS(6, 16, "lds.128 residue = lane&1 | (lane>>3)<<1");
    LDS(4, 16, "lds.128 linear");
    LDS(5, 16, "lds.128 slot = 32r + lane");
    LDS(0, 8, "lds.64 random");
    LDS(7, 8, "lds.64 residue(8B slots) = lane&7");
    LDS(4, 8, "lds.64 linear");
    LDS(0, 4, "lds.32 random");
    LDS(5, 4, "lds.32 word = 32r + lane");
    LDS(4, 4, "lds.32 linear");
    return 0;
}

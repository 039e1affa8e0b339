// tisph_voxel.cuh -- mesh -> boundary-particle sampler: surface voxelisation + interior fill.
//
// Replaces the trimesh call of load_rigid_body (partice_systemv4.py:276-277:
// mesh.voxelized(pitch=particle_diameter).fill().points).  trimesh is a host library that is not
// part of the reference tree (and not installable here), so its published behaviour is restated:
// voxel centres lie on the world-aligned lattice k * pitch; a voxel belongs to the surface when
// the triangle touches the cube of edge `pitch` centred there (trimesh reaches the same set by
// subdividing the triangles below the pitch and rounding the vertices); fill() = every voxel that
// is not connected to the outside through empty voxels (6-connectivity, scipy binary_fill_holes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tisph {

struct VoxGrid {
    int lo[3];      // lattice index of voxel (0,0,0)
    int dims[3];    // voxels per axis (the outermost layer is guaranteed empty padding)
    float pitch;
};

__device__ __forceinline__ size_t vox_index(const VoxGrid& g, int x, int y, int z) {
    return ((size_t)x * g.dims[1] + y) * g.dims[2] + z;
}

// separating-axis test of a triangle (relative to the box centre) against the box [-hw, hw]^3
__device__ __forceinline__ bool axis_sep(float ax, float ay, float az, const float v0[3], const float v1[3],
                                         const float v2[3], float hw) {
    float p0 = ax * v0[0] + ay * v0[1] + az * v0[2];
    float p1 = ax * v1[0] + ay * v1[1] + az * v1[2];
    float p2 = ax * v2[0] + ay * v2[1] + az * v2[2];
    float r = hw * (fabsf(ax) + fabsf(ay) + fabsf(az));
    return fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r;
}

__device__ bool tri_box_overlap(const float c[3], float hw, const float a[3], const float b[3], const float d[3]) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3];
    for (int k = 0; k < 3; ++k) { v0[k] = a[k] - c[k]; v1[k] = b[k] - c[k]; v2[k] = d[k] - c[k]; }
    for (int k = 0; k < 3; ++k) { e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v1[k]; e2[k] = v0[k] - v2[k]; }
    // box axes
    for (int k = 0; k < 3; ++k)
        if (fminf(v0[k], fminf(v1[k], v2[k])) > hw || fmaxf(v0[k], fmaxf(v1[k], v2[k])) < -hw) return false;
    // triangle normal
    float nx = e0[1] * e1[2] - e0[2] * e1[1], ny = e0[2] * e1[0] - e0[0] * e1[2], nz = e0[0] * e1[1] - e0[1] * e1[0];
    if (axis_sep(nx, ny, nz, v0, v1, v2, hw)) return false;
    // 9 cross products of box axes and edges
    const float* es[3] = {e0, e1, e2};
    for (int k = 0; k < 3; ++k) {
        const float* e = es[k];
        if (axis_sep(0.f, -e[2], e[1], v0, v1, v2, hw)) return false;
        if (axis_sep(e[2], 0.f, -e[0], v0, v1, v2, hw)) return false;
        if (axis_sep(-e[1], e[0], 0.f, v0, v1, v2, hw)) return false;
    }
    return true;
}

// one thread per triangle: mark every voxel of the triangle's bounding box that it touches
__global__ void __launch_bounds__(128)
k_vox_surface(VoxGrid g, const float* __restrict__ vert, const int* __restrict__ faces, int nf,
              unsigned char* __restrict__ occ) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    float a[3], b[3], d[3];
    for (int k = 0; k < 3; ++k) {
        a[k] = vert[3 * (size_t)faces[3 * f] + k];
        b[k] = vert[3 * (size_t)faces[3 * f + 1] + k];
        d[k] = vert[3 * (size_t)faces[3 * f + 2] + k];
    }
    int lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        float mn = fminf(a[k], fminf(b[k], d[k])), mx = fmaxf(a[k], fmaxf(b[k], d[k]));
        lo[k] = max((int)floorf(mn / g.pitch + 0.5f) - g.lo[k], 0);
        hi[k] = min((int)floorf(mx / g.pitch + 0.5f) - g.lo[k], g.dims[k] - 1);
    }
    const float hw = 0.5f * g.pitch;
    for (int x = lo[0]; x <= hi[0]; ++x)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int z = lo[2]; z <= hi[2]; ++z) {
                float c[3] = {(x + g.lo[0]) * g.pitch, (y + g.lo[1]) * g.pitch, (z + g.lo[2]) * g.pitch};
                if (tri_box_overlap(c, hw, a, b, d)) occ[vox_index(g, x, y, z)] = 1;
            }
}

// outside flood: one thread per (x, y) column sweeps z up and down; repeated until nothing changes
__global__ void __launch_bounds__(128)
k_vox_flood(VoxGrid g, const unsigned char* __restrict__ occ, unsigned char* __restrict__ outside,
            int* __restrict__ changed) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= g.dims[0] * g.dims[1]) return;
    int x = col / g.dims[1], y = col % g.dims[1];
    bool edge = x == 0 || y == 0 || x == g.dims[0] - 1 || y == g.dims[1] - 1;
    bool any = false;
    for (int pass = 0; pass < 2; ++pass) {
        bool prev = false;
        for (int t = 0; t < g.dims[2]; ++t) {
            int z = pass == 0 ? t : g.dims[2] - 1 - t;
            size_t i = vox_index(g, x, y, z);
            if (occ[i]) { prev = false; continue; }
            bool out = outside[i];
            if (!out) {
                out = edge || z == 0 || z == g.dims[2] - 1 || prev;
                if (!out) {
                    out = outside[vox_index(g, x - 1, y, z)] || outside[vox_index(g, x + 1, y, z)] ||
                          outside[vox_index(g, x, y - 1, z)] || outside[vox_index(g, x, y + 1, z)];
                }
                if (out) { outside[i] = 1; any = true; }
            }
            prev = out;
        }
    }
    if (any) *changed = 1;
}

__global__ void __launch_bounds__(256)
k_vox_fill(size_t n, unsigned char* __restrict__ occ, const unsigned char* __restrict__ outside) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !outside[i]) occ[i] = 1;
}

}  // namespace tisph

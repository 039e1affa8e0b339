/*
 * sph_oracle.c -- CPU restatement of Ti-SPH's per-step WCSPH loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ti_sph_b200/, core/,
 * utils/) may import, link or execute this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker / reported CPU baseline.
 *
 * It restates, loop for loop, the reference's Taichi kernels (paths relative to
 * the reference checkout):
 *   gen-2 (3D):  core/partice_system/partice_systemv4.py, core/sph/sph_basev2.py,
 *                core/sph/wcsphv2.py
 *   gen-1 (2D):  core/partice_system/partice_system.py, core/sph/sph_base.py,
 *                core/sph/wcsph.py, core/const.py
 * in IEEE binary32 arithmetic, evaluated in the reference's expression order,
 * with NO fused multiply-add (compile with -ffp-contract=off) and with the
 * serial (single-thread) semantics of its atomics (stable counting sort).
 *
 * Parity pin: the reference ships no golden vectors and Taichi is not
 * installable here, so this oracle is pinned against the tests/golden npz files,
 * which are produced by executing the reference's own, unmodified Python
 * sources under a minimal Taichi emulator (tests/golden/make_golden.py).
 *
 * Documented deviations from the literal reference:
 *   - neighbour cells outside the grid are empty (the reference reads out of
 *     bounds for cx outside [0,nx), partice_systemv4.py:341-343);
 *   - density_mode 1 ("summed") and volume_mode 1 ("akinci") are extensions;
 *     mode 0 is the literal reference behaviour.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t dim;            /* 3 for gen-2, 2 for gen-1 */
    int32_t grid_num[3];    /* partice_systemv4.py:59 */
    float h;                /* support_length = 4 r   (:34) == grid_size (:58) */
    float domain_size[3];   /* (:22) */
    float padding;          /* (:35) */
    float wall_hi[3];       /* (float)(domain_size - padding), sph_basev2.py:164 */
    float dt;               /* sph_basev2.py:14-15 */
    float g[3];             /* sph_basev2.py:16 (gen-2) ; const.py:2 on last axis (gen-1) */
    float c_s;              /* wcsphv2.py:16 */
    float rho0;             /* solver density_0 = 1000, sph_basev2.py:13 */
    float ps_density0;      /* ps.density0 from JSON, partice_systemv4.py:15 */
    float stiffness;        /* wcsphv2.py:11 */
    float exponent;         /* wcsphv2.py:10 */
    float k_w;              /* (float)(k / h^dim),   sph_basev2.py:22-30 */
    float k_dw;             /* (float)(6 k / h^dim), sph_basev2.py:42-50 */
    float visc_fluid_c;     /* (float)(2*0.05*h*c_s),   wcsphv2.py:69 */
    float visc_bound_c;     /* (float)(0.08*h*c_s),     wcsphv2.py:76 */
    float eps_h2;           /* (float)(0.01*h^2),       wcsphv2.py:72 */
    float m_V;              /* gen-1: ps.m_V (partice_system.py:23) */
    float g1_visc_c;        /* gen-1: (float)(2*(dim+2)*0.05), sph_base.py:81 */
    float g1_mass;          /* gen-1: m_V*density_0, sph_base.py:16 */
    float g1_press_c;       /* gen-1: (float)(-density_0*m_V), sph_base.py:68 */
    int32_t density_mode;   /* 0 reference (wcsphv2.py:32-34 discards the sum), 1 summed */
    int32_t volume_mode;    /* 0 reference (by-value accumulator lost), 1 akinci */
    int32_t max_per_cell;   /* gen-1: 100 (partice_system.py:25) */
    int32_t max_neighbors;  /* gen-1: 100 (partice_system.py:26) */
} ora_config;

enum { MAT_BOUNDARY = 0, MAT_FLUID = 1 };   /* partice_systemv4.py:24-25 */

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ora_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------- kernels */

/* How `q ** 3` and `(1 - q) ** 3` (sph_basev2.py:33,35) are evaluated.
 *   0 (default): repeated multiplication -- Taichi's alg_simp pass lowers pow(x, small
 *                integer constant) to multiplications;
 *   1          : libm powf (also for `** 2`), which is how numpy evaluates `**` when the reference sources run
 *                on the Taichi stand-in that produced tests/golden/ (bit-comparable mode). */
static int g_pow_mode = 0;
void ora_set_pow_mode(int m) { g_pow_mode = m; }
/* volatile exponents: gcc would otherwise fold powf(x, 2.0f) into x * x, which differs from
 * libm's powf in the last bit for ~0.07 % of the arguments */
static volatile float g_two = 2.0f, g_three = 3.0f;
static inline float cube(float x) { return g_pow_mode ? powf(x, g_three) : (x * x) * x; }
static inline float sq(float x) { return g_pow_mode ? powf(x, g_two) : x * x; }

/* sph_basev2.py:19-36 / sph_base.py:18-35 */
static inline float cubic_kernel(const ora_config *c, float r_norm) {
    float res = 0.0f;
    float q = r_norm / c->h;
    if (q <= 1.0f) {
        if (q <= 0.5f) {
            float q2 = sq(q);
            float q3 = cube(q);
            res = c->k_w * (6.0f * (q3 - q2) + 1.0f);
        } else {
            float f = 1.0f - q;
            res = (c->k_w * 2.0f) * cube(f);
        }
    }
    return res;
}

static inline float vnorm(const float *r, int dim) {
    float s = r[0] * r[0] + r[1] * r[1];
    if (dim == 3) s = s + r[2] * r[2];
    return sqrtf(s);
}
static inline float vdot(const float *a, const float *b, int dim) {
    float s = a[0] * b[0] + a[1] * b[1];
    if (dim == 3) s = s + a[2] * b[2];
    return s;
}

/* sph_basev2.py:38-61 / sph_base.py:37-60 */
static inline void cubic_kernel_derivative(const ora_config *c, const float *r, float *res) {
    int dim = c->dim;
    float r_norm = vnorm(r, dim);
    float q = r_norm / c->h;
    for (int a = 0; a < dim; ++a) res[a] = 0.0f;
    if (r_norm > 1e-5f && q <= 1.0f) {
        float den = r_norm * c->h;
        float s;
        if (q <= 0.5f) {
            s = (c->k_dw * q) * (3.0f * q - 2.0f);
        } else {
            float f = 1.0f - q;
            s = c->k_dw * (-f * f);
        }
        for (int a = 0; a < dim; ++a) res[a] = s * (r[a] / den);
    }
}

/* partice_systemv4.py:86-100 : cell = (int)(x / grid_size), key x-major */
static inline int32_t cell_key3(const ora_config *c, const float *x, int32_t *cell) {
    cell[0] = (int32_t)(x[0] / c->h);
    cell[1] = (int32_t)(x[1] / c->h);
    cell[2] = (int32_t)(x[2] / c->h);
    return cell[0] * c->grid_num[1] * c->grid_num[2] + cell[1] * c->grid_num[2] + cell[2];
}

/* -------------------------------------------------------- gen-2 : update() */

/* partice_systemv4.py:206-215 (update_gird_id) + :255 (inclusive scan).
 * keys[n], counts[ncell] (histogram), scan[ncell] (inclusive).
 * returns number of particles whose key is outside [0,ncell) (reference UB). */
int ora_bin_count(const ora_config *c, int n, const float *x, int32_t *keys,
                  int32_t *counts, int32_t *scan) {
    int64_t ncell = (int64_t)c->grid_num[0] * c->grid_num[1] * c->grid_num[2];
    int bad = 0;
    memset(counts, 0, sizeof(int32_t) * ncell);
    for (int i = 0; i < n; ++i) {
        int32_t cell[3];
        int32_t k = cell_key3(c, x + 3 * (size_t)i, cell);
        keys[i] = k;
        if (k < 0 || k >= ncell) { ++bad; continue; }
        counts[k] += 1;
    }
    int32_t run = 0;
    for (int64_t k = 0; k < ncell; ++k) { run += counts[k]; scan[k] = run; }
    return bad;
}

/* partice_systemv4.py:219-224 with serial atomics: stable counting sort.
 * new_index[i] = destination slot of particle i. */
void ora_sort_rank(const ora_config *c, int n, const int32_t *keys, const int32_t *scan,
                   int32_t *new_index) {
    int64_t ncell = (int64_t)c->grid_num[0] * c->grid_num[1] * c->grid_num[2];
    int32_t *temp = (int32_t *)malloc(sizeof(int32_t) * ncell);
    for (int64_t k = 0; k < ncell; ++k) temp[k] = scan[k] - (k > 0 ? scan[k - 1] : 0);
    for (int i = 0; i < n; ++i) {
        int j = n - 1 - i;
        int32_t k = keys[j];
        int32_t offset = 0;
        if (k - 1 >= 0) offset = scan[k - 1];
        int32_t old = temp[k];
        temp[k] = old - 1;                 /* ti.atomic_sub returns the old value */
        new_index[j] = old - 1 + offset;
    }
    free(temp);
}

/* partice_systemv4.py:226-249: scatter `words` 4-byte words per particle */
void ora_reorder(int n, int words, const int32_t *new_index, const void *src, void *dst) {
    const int32_t *s = (const int32_t *)src;
    int32_t *d = (int32_t *)dst;
    for (int i = 0; i < n; ++i)
        memcpy(d + (size_t)new_index[i] * words, s + (size_t)i * words, 4 * (size_t)words);
}

/* ------------------------------------------- gen-2 : neighbour iteration */

/* partice_systemv4.py:331-345.  Visits neighbours in the reference order:
 * offsets row-major (x slowest, z fastest), ascending sorted index inside a cell.
 * Cell range is [scan[max(0,c-1)], scan[c]) -- so cell 0 is never visited (Q3).
 * Cells outside the grid are empty (documented deviation, Q4). */
#define FOR_ALL_NEIGHBORS3(c, scan, x, i, J, ...)                                      \
    do {                                                                              \
        int32_t cc_[3];                                                               \
        cell_key3((c), (x) + 3 * (size_t)(i), cc_);                                   \
        for (int ox_ = -1; ox_ <= 1; ++ox_)                                           \
            for (int oy_ = -1; oy_ <= 1; ++oy_)                                       \
                for (int oz_ = -1; oz_ <= 1; ++oz_) {                                 \
                    int32_t cx_ = cc_[0] + ox_, cy_ = cc_[1] + oy_, cz_ = cc_[2] + oz_; \
                    if (cx_ < 0 || cx_ >= (c)->grid_num[0] || cy_ < 0 ||              \
                        cy_ >= (c)->grid_num[1] || cz_ < 0 || cz_ >= (c)->grid_num[2]) \
                        continue;                                                     \
                    int32_t g_ = cx_ * (c)->grid_num[1] * (c)->grid_num[2] +          \
                                 cy_ * (c)->grid_num[2] + cz_;                        \
                    int32_t b_ = (scan)[g_ - 1 > 0 ? g_ - 1 : 0], e_ = (scan)[g_];    \
                    for (int32_t J = b_; J < e_; ++J) {                               \
                        if (J == (i)) continue;                                       \
                        float r_[3] = {(x)[3 * (size_t)(i)] - (x)[3 * (size_t)J],     \
                                       (x)[3 * (size_t)(i) + 1] - (x)[3 * (size_t)J + 1], \
                                       (x)[3 * (size_t)(i) + 2] - (x)[3 * (size_t)J + 2]}; \
                        float rn_ = vnorm(r_, 3);                                     \
                        if (!(rn_ < (c)->h)) continue;                                \
                        __VA_ARGS__                                                   \
                    }                                                                 \
                }                                                                     \
    } while (0)

/* neighbour count per particle under the reference's strict `norm < h` rule */
void ora_neighbor_count(const ora_config *c, int n, const float *x, const int32_t *scan,
                        int32_t *count) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        int32_t cnt = 0;
        FOR_ALL_NEIGHBORS3(c, scan, x, i, j, { (void)r_; cnt++; });
        count[i] = cnt;
    }
}

/* sph_basev2.py:190-201 */
void ora_boundary_volume(const ora_config *c, int n, const float *x, const int32_t *material,
                         const int32_t *scan, float *volume) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_BOUNDARY) continue;
        float delta = cubic_kernel(c, 0.0f);
        if (c->volume_mode == 1) {
            FOR_ALL_NEIGHBORS3(c, scan, x, i, j, {
                if (material[j] == MAT_BOUNDARY) delta += cubic_kernel(c, rn_);
            });
        }
        /* volume_mode 0: compute_boundary_volume_task takes delta_bi by value
         * (no ti.template() annotation, sph_basev2.py:191) -> accumulation lost. */
        volume[i] = 1.0f / delta;
    }
}

/* wcsphv2.py:18-34.  S[i] = sum_j mass_i W(|x_ij|) (what the walk accumulates),
 * density[i] = mass_i W(0) in reference mode (line 34 overwrites the sum). */
void ora_density(const ora_config *c, int n, const float *x, const float *mass,
                 const int32_t *material, const int32_t *scan, float *density, float *S) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) { if (S) S[i] = 0.0f; continue; }
        float self = mass[i] * cubic_kernel(c, 0.0f);
        float s = 0.0f, acc = self;
        FOR_ALL_NEIGHBORS3(c, scan, x, i, j, {
            float t = mass[i] * cubic_kernel(c, rn_);   /* Q2: mass[p_i], any material of j */
            s += t;
            acc += t;
        });
        if (S) S[i] = s;
        density[i] = (c->density_mode == 1) ? acc : self;
    }
}

/* wcsphv2.py:56-93 */
void ora_non_pressure(const ora_config *c, int n, const float *x, const float *v,
                      const float *mass, const float *volume, const float *density,
                      const int32_t *material, const int32_t *scan, float *dvel) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        float a[3] = {c->g[0], c->g[1], c->g[2]};
        FOR_ALL_NEIGHBORS3(c, scan, x, i, j, {
            float vij[3] = {v[3 * (size_t)i] - v[3 * (size_t)j],
                            v[3 * (size_t)i + 1] - v[3 * (size_t)j + 1],
                            v[3 * (size_t)i + 2] - v[3 * (size_t)j + 2]};
            float gw[3];
            if (material[j] == MAT_FLUID) {
                /* :61-65 cohesion-like term */
                float s = 0.01f / mass[i] * mass[j];
                float w = cubic_kernel(c, rn_);
                for (int k = 0; k < 3; ++k) a[k] -= (s * r_[k]) * w;
                /* :68-73 artificial viscosity, fluid-fluid */
                float nu = c->visc_fluid_c / (density[i] + density[j]);
                float pi = -nu * fminf(0.0f, vdot(vij, r_, 3)) / (vdot(r_, r_, 3) + c->eps_h2);
                cubic_kernel_derivative(c, r_, gw);
                float s2 = mass[j] * pi;
                for (int k = 0; k < 3; ++k) a[k] -= s2 * gw[k];
            } else {
                /* :74-80 boundary neighbour */
                float nu = c->visc_bound_c / (2.0f * density[i]);
                float pi = -nu * fminf(vdot(vij, r_, 3), 0.0f) / (vdot(r_, r_, 3) + c->eps_h2);
                cubic_kernel_derivative(c, r_, gw);
                float s2 = c->ps_density0 * volume[j] * pi;
                for (int k = 0; k < 3; ++k) a[k] -= s2 * gw[k];
            }
        });
        dvel[3 * (size_t)i] = a[0];
        dvel[3 * (size_t)i + 1] = a[1];
        dvel[3 * (size_t)i + 2] = a[2];
    }
}

/* Test support (not a reference kernel): per fluid particle the sum of the MAGNITUDES of the terms that
 * ora_non_pressure (|g| + every pair's cohesion and viscosity vector) and ora_pressure_force (every
 * pair's pressure vector) add up, in double.  An f32 sum of N terms carries a rounding error bounded
 * by a multiple of eps * sum |term|, so this -- not |a| itself, which is what is left after the terms
 * cancel -- is the scale a relative tolerance on the accelerations refers to.
 * `density` / `pressure` are the post-EOS values (after ora_eos). */
void ora_force_magnitudes(const ora_config *c, int n, const float *x, const float *v, const float *mass,
                          const float *volume, const float *density_pre, const float *density,
                          const float *pressure, const int32_t *material, const int32_t *scan,
                          float *mag_np, float *mag_p) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        mag_np[i] = 0.0f;
        mag_p[i] = 0.0f;
        if (material[i] != MAT_FLUID) continue;
        double snp = sqrt((double)c->g[0] * c->g[0] + (double)c->g[1] * c->g[1] + (double)c->g[2] * c->g[2]);
        double sp = 0.0;
        FOR_ALL_NEIGHBORS3(c, scan, x, i, j, {
            float vij[3] = {v[3 * (size_t)i] - v[3 * (size_t)j],
                            v[3 * (size_t)i + 1] - v[3 * (size_t)j + 1],
                            v[3 * (size_t)i + 2] - v[3 * (size_t)j + 2]};
            float gw[3];
            cubic_kernel_derivative(c, r_, gw);
            double gn = sqrt((double)gw[0] * gw[0] + (double)gw[1] * gw[1] + (double)gw[2] * gw[2]);
            double mn = fmin(0.0, (double)vdot(vij, r_, 3)) / ((double)vdot(r_, r_, 3) + c->eps_h2);
            if (material[j] == MAT_FLUID) {
                snp += fabs(0.01 / mass[i] * mass[j]) * rn_ * cubic_kernel(c, rn_);
                snp += fabs(mass[j] * (c->visc_fluid_c / ((double)density_pre[i] + density_pre[j])) * mn) * gn;
                sp += fabs(mass[j] * ((double)pressure[i] / ((double)density[i] * density[i]) +
                                      (double)pressure[j] / ((double)density[j] * density[j]))) * gn;
            } else {
                snp += fabs(c->ps_density0 * volume[j] * (c->visc_bound_c / (2.0 * density_pre[i])) * mn) * gn;
                sp += fabs(c->rho0 * volume[j] * ((double)pressure[i] / ((double)density[i] * density[i]))) * gn;
            }
        });
        mag_np[i] = (float)snp;
        mag_p[i] = (float)sp;
    }
}

/* wcsphv2.py:45-48 (launch A): clamp + Tait EOS for every particle */
void ora_eos(const ora_config *c, int n, float *density, float *pressure) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        density[i] = fmaxf(density[i], c->rho0);
        pressure[i] = c->stiffness * (powf(density[i] / c->rho0, c->exponent) - 1.0f);
    }
}

/* wcsphv2.py:49-54 (launch B) + sph_basev2.py:64-78.  dvel += sum; also returns the
 * bare pressure sum in `apress` when non-NULL. */
void ora_pressure_force(const ora_config *c, int n, const float *x, const float *mass,
                        const float *volume, const float *density, const float *pressure,
                        const int32_t *material, const int32_t *scan, float *dvel,
                        float *apress) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        float d[3] = {0.0f, 0.0f, 0.0f};
        float p_rho_i = pressure[i] / sq(density[i]);
        FOR_ALL_NEIGHBORS3(c, scan, x, i, j, {
            float gw[3];
            cubic_kernel_derivative(c, r_, gw);
            if (material[j] == MAT_FLUID) {
                float s = -mass[j] * (pressure[i] / sq(density[i]) +
                                      pressure[j] / sq(density[j]));
                for (int k = 0; k < 3; ++k) d[k] += s * gw[k];
            } else if (material[j] == MAT_BOUNDARY) {
                float s = -c->rho0 * volume[j] * p_rho_i;
                for (int k = 0; k < 3; ++k) d[k] += s * gw[k];
            }
        });
        for (int k = 0; k < 3; ++k) {
            dvel[3 * (size_t)i + k] += d[k];
            if (apress) apress[3 * (size_t)i + k] = d[k];
        }
    }
}

/* wcsphv2.py:95-100 ; works for dim 2 and 3 */
void ora_advect(const ora_config *c, int n, float *x, float *v, const float *dvel,
                const int32_t *material) {
    int dim = c->dim;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        for (int k = 0; k < dim; ++k) {
            v[dim * (size_t)i + k] += c->dt * dvel[dim * (size_t)i + k];
            x[dim * (size_t)i + k] += c->dt * v[dim * (size_t)i + k];
        }
    }
}

/* sph_basev2.py:151-189, 204-208 (gen-2 only; gen-1's enforce_boundary is a no-op,
 * sph_base.py:161-166) */
void ora_enforce_boundary(const ora_config *c, int n, float *x, float *v,
                          const int32_t *material) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        float pos[3] = {x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]};
        float nrm[3] = {0.0f, 0.0f, 0.0f};
        for (int k = 0; k < 3; ++k) {
            if (pos[k] > c->wall_hi[k]) { nrm[k] += 1.0f; x[3 * (size_t)i + k] = c->wall_hi[k]; }
            if (pos[k] <= c->padding) { nrm[k] += -1.0f; x[3 * (size_t)i + k] = c->padding; }
        }
        float len = vnorm(nrm, 3);
        if (len > 1e-6f) {
            float u[3] = {nrm[0] / len, nrm[1] / len, nrm[2] / len};
            float *vi = v + 3 * (size_t)i;
            float s = (1.0f + 0.5f) * vdot(vi, u, 3);
            for (int k = 0; k < 3; ++k) vi[k] -= s * u[k];
        }
    }
}

/* One full gen-2 step (sph_basev2.py:210-214) on natural-layout SoA arrays.
 * All arrays are reordered in place like the reference (x,v,mass,volume,density,
 * pressure,material,color,m,grid_ids).  work = caller scratch:
 *   keys[n], new_index[n], counts[ncell], scan[ncell] (scan survives as the
 *   reference's grid_particles_num).  Optional outputs may be NULL. */
int ora_step_gen2(const ora_config *c, int n, float *x, float *v, float *mass, float *volume,
                  float *density, float *pressure, int32_t *material, int32_t *color,
                  float *m_unused, int32_t *keys, int32_t *new_index, int32_t *counts,
                  int32_t *scan, float *dvel, float *S_out) {
    int bad = ora_bin_count(c, n, x, keys, counts, scan);
    if (bad) return -bad;
    ora_sort_rank(c, n, keys, scan, new_index);
    size_t n3 = 3 * (size_t)n;
    float *tmp = (float *)malloc(sizeof(float) * n3);
#define REORDER(ptr, words)                                  \
    do {                                                     \
        if (ptr) {                                           \
            ora_reorder(n, (words), new_index, (ptr), tmp);  \
            memcpy((ptr), tmp, 4 * (size_t)(words) * n);     \
        }                                                    \
    } while (0)
    REORDER(x, 3); REORDER(v, 3); REORDER(mass, 1); REORDER(volume, 1);
    REORDER(density, 1); REORDER(pressure, 1); REORDER(material, 1);
    REORDER(color, 3); REORDER(m_unused, 1); REORDER(keys, 1);
#undef REORDER
    free(tmp);
    ora_boundary_volume(c, n, x, material, scan, volume);
    ora_density(c, n, x, mass, material, scan, density, S_out);
    ora_non_pressure(c, n, x, v, mass, volume, density, material, scan, dvel);
    ora_eos(c, n, density, pressure);
    ora_pressure_force(c, n, x, mass, volume, density, pressure, material, scan, dvel, NULL);
    ora_advect(c, n, x, v, dvel, material);
    ora_enforce_boundary(c, n, x, v, material);
    return 0;
}

/* ------------------------------------------------------------ gen-1 (2D) */

/* partice_system.py:91-93, 127-132 (serial atomics => cell lists in ascending i)
 * and :102-121 (search_neighbors; `break` on the first invalid cell, Q8).
 * The reference's dense grid (2560x2560 cells x 100 slots) is replaced by a sort
 * of (cell,i) pairs -- identical lists, no 2.6 GB table.
 * neighbors is [n][max_neighbors]; returns the number of list overflows
 * (cells > max_per_cell or lists > max_neighbors: reference UB). */
typedef struct { int64_t key; int32_t idx; } g1_pair;
static int g1_cmp(const void *a, const void *b) {
    const g1_pair *p = (const g1_pair *)a, *q = (const g1_pair *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}
static inline int64_t g1_find(const g1_pair *s, int n, int64_t key) { /* lower bound */
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) / 2; if (s[mid].key < key) lo = mid + 1; else hi = mid; }
    return lo;
}

int ora_g1_search_neighbors(const ora_config *c, int n, const float *x, const int32_t *material,
                            int32_t *neighbors, int32_t *neighbors_num) {
    int overflow = 0;
    int64_t ny = c->grid_num[1];
    g1_pair *s = (g1_pair *)malloc(sizeof(g1_pair) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        int32_t cx = (int32_t)(x[2 * (size_t)i] / c->h), cy = (int32_t)(x[2 * (size_t)i + 1] / c->h);
        s[i].key = (int64_t)cx * ny + cy;
        s[i].idx = i;
    }
    qsort(s, n, sizeof(g1_pair), g1_cmp);
    memset(neighbors, 0, sizeof(int32_t) * (size_t)n * c->max_neighbors);   /* :213 */
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : overflow)
    for (int i = 0; i < n; ++i) {
        if (material[i] == MAT_BOUNDARY) continue;
        int32_t cx = (int32_t)(x[2 * (size_t)i] / c->h), cy = (int32_t)(x[2 * (size_t)i + 1] / c->h);
        int cnt = 0, stop = 0;
        for (int ox = -1; ox <= 1 && !stop; ++ox)
            for (int oy = -1; oy <= 1; ++oy) {
                int32_t ax = cx + ox, ay = cy + oy;
                if (ax < 0 || ax >= c->grid_num[0] || ay < 0 || ay >= c->grid_num[1]) {
                    stop = 1;          /* `break` leaves the whole grouped loop (Q8) */
                    break;
                }
                int64_t key = (int64_t)ax * ny + ay;
                int64_t b = g1_find(s, n, key);
                int in_cell = 0;
                for (int64_t t = b; t < n && s[t].key == key; ++t, ++in_cell) {
                    int32_t j = s[t].idx;
                    if (in_cell >= c->max_per_cell) { ++overflow; break; }
                    if (j == i) continue;
                    float r[2] = {x[2 * (size_t)i] - x[2 * (size_t)j],
                                  x[2 * (size_t)i + 1] - x[2 * (size_t)j + 1]};
                    if (vnorm(r, 2) >= c->h) continue;
                    if (cnt >= c->max_neighbors) { ++overflow; continue; }
                    neighbors[(size_t)i * c->max_neighbors + cnt] = j;
                    cnt++;
                }
            }
        neighbors_num[i] = cnt;
    }
    free(s);
    return overflow;
}

/* wcsph.py:18-32 */
void ora_g1_density(const ora_config *c, int n, const float *x, const int32_t *material,
                    const int32_t *neighbors, const int32_t *neighbors_num, float *density) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        float rho = 0.0f;
        for (int t = 0; t < neighbors_num[i]; ++t) {
            int32_t j = neighbors[(size_t)i * c->max_neighbors + t];
            if (material[j] == MAT_FLUID) {
                float r[2] = {x[2 * (size_t)i] - x[2 * (size_t)j], x[2 * (size_t)i + 1] - x[2 * (size_t)j + 1]};
                rho += c->m_V * cubic_kernel(c, vnorm(r, 2));
            }
        }
        density[i] = rho * c->rho0;
    }
}

/* wcsph.py:52-65 + sph_base.py:77-84 */
void ora_g1_non_pressure(const ora_config *c, int n, const float *x, const float *v,
                         const float *density, const int32_t *material, const int32_t *neighbors,
                         const int32_t *neighbors_num, float *dvel) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        float a[2] = {0.0f, c->g[1]};            /* d_v[dim-1] = const.g */
        for (int t = 0; t < neighbors_num[i]; ++t) {
            int32_t j = neighbors[(size_t)i * c->max_neighbors + t];
            float r[2] = {x[2 * (size_t)i] - x[2 * (size_t)j], x[2 * (size_t)i + 1] - x[2 * (size_t)j + 1]};
            float vij[2] = {v[2 * (size_t)i] - v[2 * (size_t)j], v[2 * (size_t)i + 1] - v[2 * (size_t)j + 1]};
            float v_xy = vdot(vij, r, 2);
            float rn = vnorm(r, 2);
            float gw[2];
            cubic_kernel_derivative(c, r, gw);
            float s = c->g1_visc_c * (c->g1_mass / density[j]) * v_xy / (sq(rn) + c->eps_h2);
            a[0] += s * gw[0];
            a[1] += s * gw[1];
        }
        dvel[2 * (size_t)i] = a[0];
        dvel[2 * (size_t)i + 1] = a[1];
    }
}

/* wcsph.py:41-49 + sph_base.py:63-74 (boundary arm, Q9, is unreachable: skipped) */
void ora_g1_pressure_force(const ora_config *c, int n, const float *x, const float *density,
                           const float *pressure, const int32_t *material,
                           const int32_t *neighbors, const int32_t *neighbors_num, float *dvel) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        if (material[i] != MAT_FLUID) continue;
        float d[2] = {0.0f, 0.0f};
        for (int t = 0; t < neighbors_num[i]; ++t) {
            int32_t j = neighbors[(size_t)i * c->max_neighbors + t];
            if (material[j] != MAT_FLUID) continue;
            float r[2] = {x[2 * (size_t)i] - x[2 * (size_t)j], x[2 * (size_t)i + 1] - x[2 * (size_t)j + 1]};
            float gw[2];
            cubic_kernel_derivative(c, r, gw);
            float s = c->g1_press_c * (pressure[i] / sq(density[i]) +
                                       pressure[j] / sq(density[j]));
            d[0] += s * gw[0];
            d[1] += s * gw[1];
        }
        dvel[2 * (size_t)i] += d[0];
        dvel[2 * (size_t)i + 1] += d[1];
    }
}

/* Test support, gen-1 twin of ora_force_magnitudes: sums of the magnitudes of the terms that
 * ora_g1_non_pressure (|g| + viscosity pairs) and ora_g1_pressure_force add up, in double. */
void ora_g1_force_magnitudes(const ora_config *c, int n, const float *x, const float *v,
                             const float *density_pre, const float *density, const float *pressure,
                             const int32_t *material, const int32_t *neighbors,
                             const int32_t *neighbors_num, float *mag_np, float *mag_p) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        mag_np[i] = 0.0f;
        mag_p[i] = 0.0f;
        if (material[i] != MAT_FLUID) continue;
        double snp = fabs((double)c->g[1]), sp = 0.0;
        for (int t = 0; t < neighbors_num[i]; ++t) {
            int32_t j = neighbors[(size_t)i * c->max_neighbors + t];
            float r[2] = {x[2 * (size_t)i] - x[2 * (size_t)j], x[2 * (size_t)i + 1] - x[2 * (size_t)j + 1]};
            float vij[2] = {v[2 * (size_t)i] - v[2 * (size_t)j], v[2 * (size_t)i + 1] - v[2 * (size_t)j + 1]};
            float gw[2];
            cubic_kernel_derivative(c, r, gw);
            double gn = sqrt((double)gw[0] * gw[0] + (double)gw[1] * gw[1]);
            double rn = vnorm(r, 2);
            snp += fabs(c->g1_visc_c * ((double)c->g1_mass / density_pre[j]) * vdot(vij, r, 2) /
                        (rn * rn + c->eps_h2)) * gn;
            if (material[j] == MAT_FLUID)
                sp += fabs(c->g1_press_c * ((double)pressure[i] / ((double)density[i] * density[i]) +
                                            (double)pressure[j] / ((double)density[j] * density[j]))) * gn;
        }
        mag_np[i] = (float)snp;
        mag_p[i] = (float)sp;
    }
}

/* One full gen-1 step (sph_base.py:168-172).  neighbors/neighbors_num are caller
 * scratch ([n][max_neighbors], [n]) and hold the reference's neighbour table on return. */
int ora_step_gen1(const ora_config *c, int n, float *x, float *v, float *density,
                  float *pressure, const int32_t *material, int32_t *neighbors,
                  int32_t *neighbors_num, float *dvel) {
    int ov = ora_g1_search_neighbors(c, n, x, material, neighbors, neighbors_num);
    ora_g1_density(c, n, x, material, neighbors, neighbors_num, density);
    ora_g1_non_pressure(c, n, x, v, density, material, neighbors, neighbors_num, dvel);
    ora_eos(c, n, density, pressure);
    ora_g1_pressure_force(c, n, x, density, pressure, material, neighbors, neighbors_num, dvel);
    ora_advect(c, n, x, v, dvel, material);
    return ov;
}

size_t ora_config_size(void) { return sizeof(ora_config); }

/* TEST INFRASTRUCTURE — CPU restatement ("port") of the reference MPM substep.
 *
 * Restates, in plain C with O(N) memory, the arithmetic of brocbyte/realtime-deformations'
 * live MPM path (realtime-deformations/material_point_method.cpp; every function below cites the
 * lines it follows). It exists so that the CUDA path can be checked on scenes the real class cannot
 * hold (its WeightStorage allocates I*J*K*N floats, material_point_method.hpp:149-156) and on a box
 * where /root/reference does not exist. It is NOT shipped and NOT measured as the product.
 *
 * Parity status: PINNED against the reference itself. oracle/_ref/ref_mpm (the unmodified reference
 * class built by oracle/Makefile) produced tests/golden/*.npz; tests/test_oracle_golden.py checks
 * this file against them:
 *   - bit-exact: weights, cell indices, P2G mass/velocity, grid velocity update, collisions,
 *     F-update (Eigen-convention Jacobi SVD + clamp + FP), G2P, advection;
 *   - NOT bit-exact, by construction: the rotation R used by computeExplicitGridForces. The reference
 *     calls a 1.8 kLoC Higham-Noferini polar decomposition in fp32 (include/polar_decomposition_3x3*.h);
 *     here R is the limit of the Newton iteration R <- (R + R^-T)/2 in fp64, rounded to fp32
 *     (max |dR| vs the reference's own function ~1e-6, see tests). Trajectories therefore agree with
 *     the reference to within its own FMA-contraction noise floor, not bit-for-bit.
 *
 * Conventions: 3x3 matrices are float[9] in glm column-major order, m[c*3+r] == glm m[c][r].
 * All products/sums are written in the association glm 0.9.7.1 / Eigen 3.4.90 use, and this file
 * must be compiled with -ffp-contract=off (oracle/Makefile does). */
#include "mpm_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float mass, vel[3], volume, pos[3], FE[9], FP[9], B[9]; } OParticle;  /* 35 floats */
typedef struct { float mass, force[3], vel[3]; } OCell;

struct Oracle {
    int I, J, K, n, threads;
    OracleParams prm;
    float Dinv[9];         /* DpInverse, material_point_method.hpp:177 */
    OParticle* p;
    int* cell;             /* n x 3, cached by rasterize like Particle::neighs (cpp:96-98) */
    unsigned char* valid;  /* 5^3 neighbourhood inside the grid */
    OCell* g;
    int* used; int nused;  /* used_cells in lexicographic order (cpp:105-110) */
};

/* ---------------- glm 0.9.7.1 value-type arithmetic, same association ---------------- */
/* detail/type_mat3x3.inl:519-553 */
static void m3mul(float* R, const float* A, const float* B) {
    float t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            t[c * 3 + r] = A[0 * 3 + r] * B[c * 3 + 0] + A[1 * 3 + r] * B[c * 3 + 1] + A[2 * 3 + r] * B[c * 3 + 2];
    memcpy(R, t, sizeof t);
}
/* detail/type_mat3x3.inl:501-508 */
static void m3vec(float* o, const float* m, const float* v) {
    float t0 = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
    float t1 = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
    float t2 = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static void m3scale(float* R, const float* A, float s) { for (int i = 0; i < 9; ++i) R[i] = A[i] * s; }
static void m3add(float* R, const float* A, const float* B) { for (int i = 0; i < 9; ++i) R[i] = A[i] + B[i]; }
static void m3sub(float* R, const float* A, const float* B) { for (int i = 0; i < 9; ++i) R[i] = A[i] - B[i]; }
static void m3transpose(float* R, const float* A) {
    float t[9];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) t[c * 3 + r] = A[r * 3 + c];
    memcpy(R, t, sizeof t);
}
#define G(m, c, r) ((m)[(c) * 3 + (r)])
/* detail/func_matrix.inl compute_determinant<tmat3x3> */
static float m3det(const float* m) {
    return + G(m,0,0) * (G(m,1,1) * G(m,2,2) - G(m,2,1) * G(m,1,2))
           - G(m,1,0) * (G(m,0,1) * G(m,2,2) - G(m,2,1) * G(m,0,2))
           + G(m,2,0) * (G(m,0,1) * G(m,1,2) - G(m,1,1) * G(m,0,2));
}
/* detail/type_mat3x3.inl:37-56 compute_inverse */
static void m3inverse(float* R, const float* m) {
    float ood = 1.0f / (
        + G(m,0,0) * (G(m,1,1) * G(m,2,2) - G(m,2,1) * G(m,1,2))
        - G(m,1,0) * (G(m,0,1) * G(m,2,2) - G(m,2,1) * G(m,0,2))
        + G(m,2,0) * (G(m,0,1) * G(m,1,2) - G(m,1,1) * G(m,0,2)));
    float t[9];
    G(t,0,0) = + (G(m,1,1) * G(m,2,2) - G(m,2,1) * G(m,1,2)) * ood;
    G(t,1,0) = - (G(m,1,0) * G(m,2,2) - G(m,2,0) * G(m,1,2)) * ood;
    G(t,2,0) = + (G(m,1,0) * G(m,2,1) - G(m,2,0) * G(m,1,1)) * ood;
    G(t,0,1) = - (G(m,0,1) * G(m,2,2) - G(m,2,1) * G(m,0,2)) * ood;
    G(t,1,1) = + (G(m,0,0) * G(m,2,2) - G(m,2,0) * G(m,0,2)) * ood;
    G(t,2,1) = - (G(m,0,0) * G(m,2,1) - G(m,2,0) * G(m,0,1)) * ood;
    G(t,0,2) = + (G(m,0,1) * G(m,1,2) - G(m,1,1) * G(m,0,2)) * ood;
    G(t,1,2) = - (G(m,0,0) * G(m,1,2) - G(m,1,0) * G(m,0,2)) * ood;
    G(t,2,2) = + (G(m,0,0) * G(m,1,1) - G(m,1,0) * G(m,0,1)) * ood;
    memcpy(R, t, sizeof t);
}
static const float ID3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };

/* ---------------- weights: material_point_method.hpp:20-31 ---------------- */
float oracle_weight(float x) {
    const float modx = fabsf(x);
    const float modx2 = modx * modx;
    const float modx3 = modx * modx * modx;
    if ((double)modx < 1.0) return (float)(0.5 * (double)modx3 - (double)modx2 + 2.0 / 3.0);
    if ((double)modx < 2.0) {
        const float a = 2 - modx;   /* int - float -> float, three times the same value */
        return (float)((1.0 / 6.0) * (double)a * (double)a * (double)a);
    }
    return 0.0f;
}

/* quadratic B-spline (stencil = 1; not in the reference, same style: fp64 polynomial, float result):
 * N(x) = 3/4 - x^2 (|x| < 1/2), (3/2 - |x|)^2 / 2 (|x| < 3/2), 0 otherwise */
float oracle_weight_quadratic(float x) {
    const float modx = fabsf(x);
    if ((double)modx < 0.5) return (float)(0.75 - (double)modx * (double)modx);
    if ((double)modx < 1.5) { const double a = 1.5 - (double)modx; return (float)(0.5 * a * a); }
    return 0.0f;
}

/* per-axis weights of the 5 enumerated nodes cell-2..cell+2 (cpp:84-91, hpp:53-58) */
static void axis_weights(float pos, float h, int cell, float w[5], int stencil) {
    for (int d = 0; d < 5; ++d) {
        const int idx = cell + d - 2;
        const float comp = pos / h - (float)idx;
        w[d] = stencil == 1 ? oracle_weight_quadratic(comp) : oracle_weight(comp);
    }
}

void oracle_default_params(OracleParams* p) {
    p->h = 0.05f; p->E = 1.4e5f; p->nu = 0.2f; p->xi = 10.0f;
    p->theta_c = 2.5f * 1e-2; p->theta_s = 5.0f * 1e-3;
    p->gravity[0] = 0.0f; p->gravity[1] = (float)-9.8; p->gravity[2] = 0.0f;
    p->friction = 0.5f;
    p->stencil = 0;
}

Oracle* oracle_create(int I, int J, int K, int n, const OracleParams* prm) {
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->I = I; o->J = J; o->K = K; o->n = n; o->threads = 1;
    if (prm) o->prm = *prm; else oracle_default_params(&o->prm);
    /* hpp:177: inverse(mat3(1.0) * (1.0f/3.0f) * h * h) */
    float d[9];
    m3scale(d, ID3, o->prm.stencil == 1 ? 1.0f / 4.0f : 1.0f / 3.0f); m3scale(d, d, o->prm.h); m3scale(d, d, o->prm.h);     /* (quadratic: D = h^2/4) */
    m3inverse(o->Dinv, d);
    o->p = (OParticle*)calloc((size_t)n, sizeof(OParticle));
    for (int i = 0; i < n; ++i) { memcpy(o->p[i].FE, ID3, sizeof ID3); memcpy(o->p[i].FP, ID3, sizeof ID3); }
    o->cell = (int*)calloc((size_t)n * 3, sizeof(int));
    o->valid = (unsigned char*)calloc((size_t)n, 1);
    o->g = (OCell*)calloc((size_t)I * J * K, sizeof(OCell));
    o->used = (int*)malloc((size_t)I * J * K * sizeof(int));
    return o;
}
void oracle_destroy(Oracle* o) { if (!o) return; free(o->p); free(o->cell); free(o->valid); free(o->g); free(o->used); free(o); }
void oracle_set_threads(Oracle* o, int t) { o->threads = t < 1 ? 1 : t; }
void oracle_set_particles(Oracle* o, const float* s) { memcpy(o->p, s, (size_t)o->n * sizeof(OParticle)); }
void oracle_get_particles(const Oracle* o, float* s) { memcpy(s, o->p, (size_t)o->n * sizeof(OParticle)); }
void oracle_get_grid(const Oracle* o, float* g7) { memcpy(g7, o->g, (size_t)o->I * o->J * o->K * sizeof(OCell)); }
int oracle_num_used_cells(const Oracle* o) { return o->nused; }
/* inject a grid (stage-isolation tests); rebuilds used_cells by the reference's rule mass != 0 (cpp:105-110)
 * and the cached neighbourhoods from the current positions */
static void cache_neighbourhood(Oracle* o);
void oracle_set_grid(Oracle* o, const float* g7) {
    const size_t ncell = (size_t)o->I * o->J * o->K;
    memcpy(o->g, g7, ncell * sizeof(OCell));
    o->nused = 0;
    for (size_t c = 0; c < ncell; ++c) if (o->g[c].mass != 0.0f) o->used[o->nused++] = (int)c;
    cache_neighbourhood(o);
}
void oracle_cell_indices(const Oracle* o, int* c3) {
    for (int i = 0; i < o->n; ++i)
        for (int a = 0; a < 3; ++a) c3[i * 3 + a] = (int)(o->p[i].pos[a] / o->prm.h);
}
int oracle_num_out_of_grid(const Oracle* o) { int c = 0; for (int i = 0; i < o->n; ++i) c += !o->valid[i]; return c; }

static inline size_t nidx(const Oracle* o, int i, int j, int k) { return ((size_t)i * o->J + j) * o->K + k; }

static void cache_neighbourhood(Oracle* o) {   /* getParticleNeighs, cpp:80-93 */
    for (int i = 0; i < o->n; ++i) {
        int ok = 1;
        const int dims[3] = { o->I, o->J, o->K };
        for (int a = 0; a < 3; ++a) {
            const int c = (int)(o->p[i].pos[a] / o->prm.h);   /* glm::ivec3(pos / h): IEEE divide, truncate */
            o->cell[i * 3 + a] = c;
            if (c - 2 < 0 || c + 2 > dims[a] - 1) ok = 0;
        }
        o->valid[i] = (unsigned char)ok;
    }
}

/* scatter helper: runs body(particle, node index, w, Xi - pos) over the 125 nodes in the reference's
 * dx,dy,dz order (cpp:84-91); nodes whose weight is exactly 0 contribute +-0 and are skipped */
#define FOR_NEIGHBOURS(o, pi, ...)                                                                    \
    do {                                                                                              \
        const OParticle* P_ = &(o)->p[pi];                                                            \
        const int* c_ = &(o)->cell[(pi) * 3];                                                         \
        float wx_[5], wy_[5], wz_[5];                                                                 \
        axis_weights(P_->pos[0], (o)->prm.h, c_[0], wx_, (o)->prm.stencil);                                             \
        axis_weights(P_->pos[1], (o)->prm.h, c_[1], wy_, (o)->prm.stencil);                                             \
        axis_weights(P_->pos[2], (o)->prm.h, c_[2], wz_, (o)->prm.stencil);                                             \
        for (int dx_ = 0; dx_ < 5; ++dx_) for (int dy_ = 0; dy_ < 5; ++dy_) for (int dz_ = 0; dz_ < 5; ++dz_) { \
            const float w = wx_[dx_] * wy_[dy_] * wz_[dz_];                                           \
            if (w == 0.0f) continue;                                                                  \
            const int ni_ = c_[0] + dx_ - 2, nj_ = c_[1] + dy_ - 2, nk_ = c_[2] + dz_ - 2;            \
            const size_t node = nidx((o), ni_, nj_, nk_);                                             \
            float dxi[3];  /* Xi - pos with Xi = vec3(neigh) * h (cpp:114) */                         \
            dxi[0] = (float)ni_ * (o)->prm.h - P_->pos[0];                                            \
            dxi[1] = (float)nj_ * (o)->prm.h - P_->pos[1];                                            \
            dxi[2] = (float)nk_ * (o)->prm.h - P_->pos[2];                                            \
            __VA_ARGS__                                                                                    \
        }                                                                                             \
    } while (0)

/* ---------------- P2G: material_point_method.cpp:94-129 ---------------- */
/* threads > 1 (timing runs only): every thread scatters a contiguous range of particles into a private grid and the
 * private grids are summed in thread order. Deterministic, but NOT the reference's summation order; parity tests and
 * golden checks use threads == 1. */
static void scatter_mass(Oracle* o, OCell* g, int i0, int i1) {
    for (int i = i0; i < i1; ++i) {                              /* cpp:99-103 */
        if (!o->valid[i]) continue;
        FOR_NEIGHBOURS(o, i, { (void)dxi; g[node].mass += P_->mass * w; });
    }
}
static void scatter_momentum(Oracle* o, OCell* g, int i0, int i1) {
    for (int i = i0; i < i1; ++i) {                              /* cpp:112-117 */
        if (!o->valid[i]) continue;
        float BD[9];
        m3mul(BD, o->p[i].B, o->Dinv);
        FOR_NEIGHBOURS(o, i, {
            float a[3];
            m3vec(a, BD, dxi);
            const float wm = w * P_->mass;
            g[node].vel[0] += wm * (P_->vel[0] + a[0]);
            g[node].vel[1] += wm * (P_->vel[1] + a[1]);
            g[node].vel[2] += wm * (P_->vel[2] + a[2]);
        });
    }
}
typedef void (*scatter_fn)(Oracle*, OCell*, int, int);
static void run_scatter(Oracle* o, scatter_fn fn) {
    const int T = o->threads;
    if (T <= 1) { fn(o, o->g, 0, o->n); return; }
    const size_t ncell = (size_t)o->I * o->J * o->K;
    OCell* priv = (OCell*)calloc(ncell * (size_t)T, sizeof(OCell));
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        const int i0 = (int)((long long)o->n * t / T), i1 = (int)((long long)o->n * (t + 1) / T);
        fn(o, priv + ncell * (size_t)t, i0, i1);
    }
#pragma omp parallel for num_threads(T) schedule(static)
    for (long long c = 0; c < (long long)ncell; ++c)
        for (int t = 0; t < T; ++t) {
            const OCell* q = &priv[ncell * (size_t)t + (size_t)c];
            o->g[c].mass += q->mass;
            for (int a = 0; a < 3; ++a) { o->g[c].vel[a] += q->vel[a]; o->g[c].force[a] += q->force[a]; }
        }
    free(priv);
}

void oracle_rasterize_particles_to_grid(Oracle* o) {
    const size_t ncell = (size_t)o->I * o->J * o->K;
    memset(o->g, 0, ncell * sizeof(OCell));                      /* grid.clear(), cpp:95 */
    cache_neighbourhood(o);
    OCell* g = o->g;
    run_scatter(o, scatter_mass);
    o->nused = 0;                                                /* cpp:105-110 */
    for (size_t c = 0; c < ncell; ++c) if (g[c].mass != 0.0f) o->used[o->nused++] = (int)c;
    run_scatter(o, scatter_momentum);
    for (int u = 0; u < o->nused; ++u) {                         /* cpp:118-121 */
        OCell* c = &g[o->used[u]];
        c->vel[0] /= c->mass; c->vel[1] /= c->mass; c->vel[2] /= c->mass;
    }
}

/* ---------------- volumes: cpp:131-142 ---------------- */
void oracle_compute_particle_volumes_and_densities(Oracle* o) {
    const float h = o->prm.h;
    for (int i = 0; i < o->n; ++i) {
        float density = 0.0f;
        if (o->valid[i]) FOR_NEIGHBOURS(o, i, { (void)dxi; density += o->g[node].mass * w; });
        density /= (h * h * h);
        o->p[i].volume = density != 0.0f ? (o->p[i].mass / density) : 0;
    }
}

/* rotation factor of the polar decomposition (see header comment: not the reference's algorithm) */
void oracle_polar_rotation(const float F[9], float R[9]) {
    double X[9], Y[9];
    for (int i = 0; i < 9; ++i) X[i] = F[i];
    for (int it = 0; it < 60; ++it) {
        /* Y = X^-T via cofactors */
        const double a = X[0], b = X[3], c = X[6], d = X[1], e = X[4], f = X[7], g = X[2], hh = X[5], k = X[8];
        /* math matrix rows: [a b c; d e f; g hh k] (X is column-major) */
        const double c00 = e * k - f * hh, c01 = -(d * k - f * g), c02 = d * hh - e * g;
        const double c10 = -(b * k - c * hh), c11 = a * k - c * g, c12 = -(a * hh - b * g);
        const double c20 = b * f - c * e, c21 = -(a * f - c * d), c22 = a * e - b * d;
        const double det = a * c00 + b * c01 + c * c02;
        if (det == 0.0 || !isfinite(det)) break;
        /* inverse^T (row r, col cc) = cofactor(r, cc) / det ; store column-major */
        const double cof[3][3] = { { c00, c01, c02 }, { c10, c11, c12 }, { c20, c21, c22 } };
        double diff = 0.0;
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) {
            const double v = 0.5 * (X[cc * 3 + r] + cof[r][cc] / det);
            diff = fmax(diff, fabs(v - X[cc * 3 + r]));
            Y[cc * 3 + r] = v;
        }
        memcpy(X, Y, sizeof X);
        if (diff < 1e-15) break;
    }
    for (int i = 0; i < 9; ++i) R[i] = (float)X[i];
}

/* ---------------- forces: cpp:235-254 ---------------- */
static void scatter_forces(Oracle* o, OCell* g, int i0, int i1) {
    const OracleParams* q = &o->prm;
    for (int i = i0; i < i1; ++i) {
        if (!o->valid[i]) continue;
        const OParticle* P = &o->p[i];
        const float poisson = q->nu, E = q->E;
        const float mu0 = E / (2.0f * (1 + poisson));
        const float mu = mu0 * expf(q->xi * (1 - m3det(P->FP)));
        const float lambda0 = (E * poisson) / ((1 + poisson) * (1 - 2 * poisson));
        const float lambda = lambda0 * expf(q->xi * (1 - m3det(P->FP)));
        float R[9], FT[9], d1[9], d2[9], dpsi[9], M[9], Ft[9];
        oracle_polar_rotation(P->FE, R);
        const float Jd = m3det(P->FE);
        m3inverse(FT, P->FE); m3transpose(FT, FT);
        m3sub(d1, P->FE, R); m3scale(d1, d1, 2 * mu);            /* 2 * mu * (F - R) */
        m3scale(d2, FT, lambda * (Jd - 1) * Jd);                  /* lambda * (J - 1) * J * FT */
        m3add(dpsi, d1, d2);
        m3scale(M, o->Dinv, P->volume);                           /* p.volume * DpInverse */
        m3mul(M, M, dpsi);
        m3transpose(Ft, P->FE);
        m3mul(M, M, Ft);
        FOR_NEIGHBOURS(o, i, {
            float Mw[9], f[3];
            m3scale(Mw, M, w);
            m3vec(f, Mw, dxi);
            g[node].force[0] -= f[0]; g[node].force[1] -= f[1]; g[node].force[2] -= f[2];
        });
    }
}
void oracle_compute_explicit_grid_forces(Oracle* o) { run_scatter(o, scatter_forces); }

/* ---------------- grid velocities: cpp:256-262 ---------------- */
void oracle_grid_velocities_update(Oracle* o, float dt) {
    for (int u = 0; u < o->nused; ++u) {
        OCell* c = &o->g[o->used[u]];
        for (int a = 0; a < 3; ++a) c->vel[a] += dt * (c->force[a] / c->mass + o->prm.gravity[a]);
    }
}

/* ---------------- collisions: hpp:79-86, mathy.hpp:40-56, cpp:264-304 ---------------- */
float oracle_box_sdf(const OracleBoxCollider* c, const float pos[3]) {
    const float* m = c->world_to_local;   /* glm mat4 column-major: m[col*4+row] */
    float p[3];
    for (int r = 0; r < 3; ++r)           /* detail/type_mat4x4.inl: (m0*x + m1*y) + (m2*z + m3*w), w = 1 */
        p[r] = (m[0 * 4 + r] * pos[0] + m[1 * 4 + r] * pos[1]) + (m[2 * 4 + r] * pos[2] + m[3 * 4 + r] * 1.0f);
    if (c->half_extent[1] < 0.0f) {   /* sphere marker of mpm_sphere_collider (not a reference shape): |p_local| - radius */
        const float s2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        return sqrtf(s2) - c->half_extent[0];
    }
    const float qx = fabsf(p[0]) - c->half_extent[0], qy = fabsf(p[1]) - c->half_extent[1], qz = fabsf(p[2]) - c->half_extent[2];
    float mx = qx; if (mx < qy) mx = qy; if (mx < qz) mx = qz; if (mx < 0.0f) mx = 0.0f;   /* std::max({q.x,q.y,q.z,0}) */
    float in = qy < qz ? qz : qy; in = qx < in ? in : qx; in = 0.0f < in ? 0.0f : in;      /* std::min({max(..), 0}) */
    return fabsf(mx) + in;
}

void oracle_body_collision(const float pos[3], const float vel[3], const OracleBoxCollider* cs, int nc, float friction, float out[3]) {
    int all_out = 1;
    for (int k = 0; k < nc; ++k) if (!(oracle_box_sdf(&cs[k], pos) > 0)) { all_out = 0; break; }
    out[0] = vel[0]; out[1] = vel[1]; out[2] = vel[2];
    if (all_out) return;
    const float delta = 0.001f;
    for (int k = 0; k < nc; ++k) {
        if (oracle_box_sdf(&cs[k], pos) > 0) continue;
        float n[3];
        for (int a = 0; a < 3; ++a) {
            float lo[3], hi[3];
            for (int b = 0; b < 3; ++b) {
                const float step = (a == b ? 1.0f : 0.0f) * delta;
                lo[b] = pos[b] - step; hi[b] = pos[b] + step;
            }
            const float f1 = oracle_box_sdf(&cs[k], lo), f2 = oracle_box_sdf(&cs[k], hi);
            n[a] = (f2 - f1) / (2.0f * delta);
        }
        float rel[3];
        for (int a = 0; a < 3; ++a) rel[a] = out[a] - cs[k].velocity[a];
        const float vn = rel[0] * n[0] + rel[1] * n[1] + rel[2] * n[2];
        if (vn >= 0) continue;
        float vt[3], vrel[3] = { 0.0f, 0.0f, 0.0f };
        for (int a = 0; a < 3; ++a) vt[a] = rel[a] - n[a] * vn;
        /* cpp:290-291: glm::vec3::length() is the component count 3, not the norm */
        if (3 > -friction * vn) {
            const float s = friction * vn / 3;
            for (int a = 0; a < 3; ++a) vrel[a] = vt[a] + vt[a] * s;
        }
        for (int a = 0; a < 3; ++a) out[a] = vrel[a] + cs[k].velocity[a];
    }
}

void oracle_grid_based_collisions(Oracle* o, float dt, const OracleBoxCollider* cs, int nc) {
    (void)dt;
    const int JK = o->J * o->K;
    for (int u = 0; u < o->nused; ++u) {
        const int c = o->used[u];
        const int i = c / JK, j = (c % JK) / o->K, k = c % o->K;
        const float pos[3] = { (float)i * o->prm.h, (float)j * o->prm.h, (float)k * o->prm.h };
        float v[3];
        oracle_body_collision(pos, o->g[c].vel, cs, nc, o->prm.friction, v);
        o->g[c].vel[0] = v[0]; o->g[c].vel[1] = v[1]; o->g[c].vel[2] = v[2];
    }
}

/* ---------------- Eigen 3.4.90 JacobiSVD<MatrixXf, FullU|FullV> on a 3x3 ----------------
 * external/Eigen/src/SVD/JacobiSVD.h:689-817, misc/RealSvd2x2.h:21-51, Jacobi/Jacobi.h:96-126,326-337.
 * Matrices here are ROW-major (a[r*3+c] == Eigen m(r,c)). Returns 0 on success, 1 on non-finite input. */
int oracle_svd3(const float A[9], float U[9], float S[3], float V[9]) {
    float W[9];
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) {                         /* maxCoeff<PropagateNaN>: a NaN, once seen, sticks */
        const float a = fabsf(A[i]);
        if (scale != scale) break;
        if (a != a || a > scale) scale = a;
    }
    if (!isfinite(scale)) return 1;
    if (scale == 0.0f) scale = 1.0f;
    for (int i = 0; i < 9; ++i) { W[i] = A[i] / scale; U[i] = ID3[i]; V[i] = ID3[i]; }
    const float precision = 2.0f * FLT_EPSILON, considerAsZero = FLT_MIN;
    float maxDiag = fabsf(W[0]);
    if (fabsf(W[4]) > maxDiag) maxDiag = fabsf(W[4]);
    if (fabsf(W[8]) > maxDiag) maxDiag = fabsf(W[8]);
    int finished = 0;
    while (!finished) {
        finished = 1;
        for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {
            const float pm = precision * maxDiag;
            const float threshold = considerAsZero < pm ? pm : considerAsZero;
            if (fabsf(W[p * 3 + q]) > threshold || fabsf(W[q * 3 + p]) > threshold) {
                finished = 0;
                /* real_2x2_jacobi_svd */
                float m00 = W[p * 3 + p], m01 = W[p * 3 + q], m10 = W[q * 3 + p], m11 = W[q * 3 + q];
                float c1, s1;
                const float t = m00 + m11, d = m10 - m01;
                if (fabsf(d) < FLT_MIN) { s1 = 0.0f; c1 = 1.0f; }
                else { const float u = t / d; const float tmp = sqrtf(1.0f + u * u); s1 = 1.0f / tmp; c1 = u / tmp; }
                if (!(c1 == 1.0f && s1 == 0.0f)) {       /* m.applyOnTheLeft(0,1,rot1) */
                    const float a0 = c1 * m00 + s1 * m10, b0 = -s1 * m00 + c1 * m10;
                    const float a1 = c1 * m01 + s1 * m11, b1 = -s1 * m01 + c1 * m11;
                    m00 = a0; m10 = b0; m01 = a1; m11 = b1;
                }
                float cr, sr;                             /* j_right.makeJacobi(m00, m01, m11) */
                const float deno = 2.0f * fabsf(m01);
                if (deno < FLT_MIN) { cr = 1.0f; sr = 0.0f; }
                else {
                    const float tau = (m00 - m11) / deno;
                    const float w = sqrtf(tau * tau + 1.0f);
                    const float tt = tau > 0.0f ? 1.0f / (tau + w) : 1.0f / (tau - w);
                    const float sign_t = tt > 0.0f ? 1.0f : -1.0f;
                    const float nn = 1.0f / sqrtf(tt * tt + 1.0f);
                    sr = -sign_t * (m01 / fabsf(m01)) * fabsf(tt) * nn;
                    cr = nn;
                }
                /* j_left = rot1 * j_right.transpose(), Jacobi.h:55-61 with other = (cr, -sr) */
                const float cl = c1 * cr - s1 * (-sr);
                const float sl = c1 * (-sr) + s1 * cr;
                /* W.applyOnTheLeft(p,q,j_left): rows p,q */
                if (!(cl == 1.0f && sl == 0.0f)) for (int c = 0; c < 3; ++c) {
                    const float x = W[p * 3 + c], y = W[q * 3 + c];
                    W[p * 3 + c] = cl * x + sl * y; W[q * 3 + c] = -sl * x + cl * y;
                }
                /* U.applyOnTheRight(p,q,j_left.transpose()): columns p,q with rotation (cl, sl) */
                if (!(cl == 1.0f && sl == 0.0f)) for (int r = 0; r < 3; ++r) {
                    const float x = U[r * 3 + p], y = U[r * 3 + q];
                    U[r * 3 + p] = cl * x + sl * y; U[r * 3 + q] = -sl * x + cl * y;
                }
                /* W.applyOnTheRight(p,q,j_right), V.applyOnTheRight(p,q,j_right): rotation (cr, -sr) */
                if (!(cr == 1.0f && -sr == 0.0f)) {
                    const float s = -sr;
                    for (int r = 0; r < 3; ++r) {
                        const float x = W[r * 3 + p], y = W[r * 3 + q];
                        W[r * 3 + p] = cr * x + s * y; W[r * 3 + q] = -s * x + cr * y;
                    }
                    for (int r = 0; r < 3; ++r) {
                        const float x = V[r * 3 + p], y = V[r * 3 + q];
                        V[r * 3 + p] = cr * x + s * y; V[r * 3 + q] = -s * x + cr * y;
                    }
                }
                float dm = fabsf(W[p * 3 + p]); if (dm < fabsf(W[q * 3 + q])) dm = fabsf(W[q * 3 + q]);
                if (maxDiag < dm) maxDiag = dm;
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        const float a = W[i * 3 + i];
        S[i] = fabsf(a);
        if (a < 0.0f) for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
    for (int i = 0; i < 3; ++i) {                         /* selection sort, first maximum wins */
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (S[k] > S[pos]) pos = k;
        if (S[pos] == 0.0f) break;
        if (pos != i) {
            float t = S[i]; S[i] = S[pos]; S[pos] = t;
            for (int r = 0; r < 3; ++r) {
                t = U[r * 3 + i]; U[r * 3 + i] = U[r * 3 + pos]; U[r * 3 + pos] = t;
                t = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + pos]; V[r * 3 + pos] = t;
            }
        }
    }
    return 0;
}

/* ---------------- F-update: cpp:306-330, utils.h:15-33 ---------------- */
int oracle_update_deformation_gradient(Oracle* o, float dt) {
    /* cpp:320 clamps to (float)(1 - 2.5f*1e-2), (float)(1 + 5.0f*1e-3) (double arithmetic, then rounded);
     * with theta as a float parameter the same rule is 1 -/+ theta in double, rounded once */
    const float clo = (float)(1.0 - (double)o->prm.theta_c);
    const float chi = (float)(1.0 + (double)o->prm.theta_s);
    int failed = 0;
#pragma omp parallel for num_threads(o->threads) schedule(static) if (o->threads > 1)
    for (int i = 0; i < o->n; ++i) {
        if (failed) continue;
        OParticle* P = &o->p[i];
        float T[9], Fh[9], FPinv[9], U[9], S[3], V[9], Sg[9], Vt[9], FEinv[9];
        m3mul(T, P->B, o->Dinv); m3scale(T, T, dt); m3add(T, ID3, T);    /* m3t(1.0) + B * DpInverse * dt */
        m3mul(T, T, P->FE); m3mul(T, T, P->FP);                           /* FPn1 */
        m3inverse(FPinv, P->FP);
        m3mul(Fh, T, FPinv);                                             /* FEpKryshka */
        /* glmToEigen: Eigen m(i,j) = glm mat[i][j] -> the glm array read as a row-major matrix;
         * eigenToGlm maps back the same way, so U/V arrays are used as glm matrices unchanged */
        if (oracle_svd3(Fh, U, S, V)) { failed = 1; continue; }          /* early return, cpp:313-316 (sequential order when threads == 1) */
        for (int k = 0; k < 3; ++k) { float s = S[k]; if (s < clo) s = clo; if (chi < s) s = chi; S[k] = s; }
        memset(Sg, 0, sizeof Sg); Sg[0] = S[0]; Sg[4] = S[1]; Sg[8] = S[2];
        m3transpose(Vt, V);
        m3mul(P->FE, U, Sg); m3mul(P->FE, P->FE, Vt);                     /* U * S * transpose(V) */
        m3inverse(FEinv, P->FE);
        m3mul(P->FP, FEinv, T);
    }
    return failed;
}

/* ---------------- G2P: cpp:332-342 ---------------- */
void oracle_update_particle_velocities(Oracle* o) {
    const OCell* g = o->g;
#pragma omp parallel for num_threads(o->threads) schedule(static) if (o->threads > 1)
    for (int i = 0; i < o->n; ++i) {
        OParticle* Pw = &o->p[i];
        float v[3] = { 0, 0, 0 }, B[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        if (o->valid[i]) FOR_NEIGHBOURS(o, i, {
            const float* gv = g[node].vel;
            v[0] += gv[0] * w; v[1] += gv[1] * w; v[2] += gv[2] * w;
            /* w * outerProduct(gv, dxi): column c = gv * dxi[c] */
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) B[c * 3 + r] += w * (gv[r] * dxi[c]);
        });
        memcpy(Pw->vel, v, sizeof v); memcpy(Pw->B, B, sizeof B);
    }
}

/* ---------------- advect + clamp: cpp:344-350, 381-388 ---------------- */
void oracle_update_particle_positions(Oracle* o, float dt) {
    const float h = o->prm.h;
    const float lo = (float)(3 * h);
    const float hi[3] = { (float)((o->I - 3) * h), (float)((o->J - 3) * h), (float)((o->K - 3) * h) };
    for (int i = 0; i < o->n; ++i) {
        OParticle* P = &o->p[i];
        for (int a = 0; a < 3; ++a) {
            float x = P->pos[a] + P->vel[a] * dt;
            if (x < lo) x = lo;
            if (hi[a] < x) x = hi[a];
            P->pos[a] = x;
        }
    }
}

/* main.cpp:192-218 */
void oracle_substep(Oracle* o, float dt, const OracleBoxCollider* c, int nc, int nsteps) {
    for (int s = 0; s < nsteps; ++s) {
        oracle_rasterize_particles_to_grid(o);
        oracle_compute_explicit_grid_forces(o);
        oracle_grid_velocities_update(o, dt);
        oracle_grid_based_collisions(o, dt, c, nc);
        if (oracle_update_deformation_gradient(o, dt)) { /* reference logs and carries on with later stages */ }
        oracle_update_particle_velocities(o);
        oracle_update_particle_positions(o, dt);
    }
}

/* =====================================================================================================
 * Implicit (optimisation-based) time integration: material_point_method.cpp:160-233 + mathy.hpp:10-38 +
 * the vendored mcloptlib (external/mcloptlib/include/MCL: LBFGS.hpp, Backtracking.hpp, Problem.hpp).
 * The reference never calls timeIntegration (README.md:17 lists it as a TODO); its objective and its
 * optimiser settings are restated here as written, including the placeholders mu0 = lambda0 = 1
 * (hpp:212-214) and the hardening factor exp(xi*1 - det FP) (cpp:187-191).
 * Not restated as written: the gradient. The library differentiates Energy by central differences with
 * eps = 2.2204e-6 in float (Problem.hpp:87-107), which is below the resolution of a float velocity above
 * 16 m/s and below the rounding noise of a float Energy everywhere, so the reference's own search
 * direction is noise; gradient = 1 reproduces that rule for completeness, gradient = 0 (the checker for the
 * CUDA path) is the analytic derivative of the same objective, evaluated in double.
 * ===================================================================================================== */
float oracle_weight_derivative(float x) {          /* hpp:32-52 */
    const float modx = fabsf(x);
    if (modx < 1.0f) {
        if (x >= 0) return (float)(+3.0 / 2.0 * (double)x * (double)x - (double)(2 * x));
        return (float)(-3.0 / 2.0 * (double)x * (double)x - (double)(2 * x));
    } else if ((double)modx < 2.0) {
        if (x >= 0) { const float a = 2 - x; return (float)(-0.5 * (double)a * (double)a); }
        { const float a = 2 + x; return (float)(0.5 * (double)a * (double)a); }
    }
    return 0.0f;
}
void oracle_default_implicit_params(OracleImplicitParams* q) {
    q->mu0 = 1.0f; q->lambda0 = 1.0f; q->xi = 10.0f;      /* hpp:212-214 */
    q->hardening = 0;
    q->max_iters = 50;                                    /* LBFGS.hpp:48 */
    q->ls_decrease = 1e-4f; q->ls_tau = 0.7f; q->ls_max_iters = 100000;   /* Minimizer.hpp:62-63, Backtracking.hpp:46 */
    q->tol_grad = 1e-2f; q->tol_step = 1e-2f;             /* mathy.hpp:31-35 */
    q->gradient = 0;
}
void oracle_used_cells(const Oracle* o, int* idx) { memcpy(idx, o->used, (size_t)o->nused * sizeof(int)); }

/* wipGrad, hpp:59-71: comp = (pos - idx*h)/h (NOT pos/h - idx as in wipHost) */
static void wip_grad(const Oracle* o, const float pos[3], int ni, int nj, int nk, float g[3]) {
    const float h = o->prm.h;
    const float xc = (pos[0] - (float)ni * h) / h, yc = (pos[1] - (float)nj * h) / h, zc = (pos[2] - (float)nk * h) / h;
    const float wx = oracle_weight(xc), wy = oracle_weight(yc), wz = oracle_weight(zc);
    const double ih = 1.0 / (double)h;
    g[0] = (float)(ih * (double)oracle_weight_derivative(xc) * (double)wy * (double)wz);
    g[1] = (float)(ih * (double)wx * (double)oracle_weight_derivative(yc) * (double)wz);
    g[2] = (float)(ih * (double)wx * (double)wy * (double)oracle_weight_derivative(zc));
}
static float hardening_factor(const OracleImplicitParams* q, float detFP) {
    return q->hardening == 0 ? expf(q->xi * 1 - detFP) : expf(q->xi * (1 - detFP));
}
/* dense map node -> position in used_cells (or -1), built per call */
static int* used_rank(const Oracle* o) {
    const size_t N = (size_t)o->I * o->J * o->K;
    int* r = (int*)malloc(N * sizeof(int));
    for (size_t i = 0; i < N; ++i) r[i] = -1;
    for (int u = 0; u < o->nused; ++u) r[o->used[u]] = u;
    return r;
}
/* trial elastic deformation gradient of one particle, cpp:176-185: (I + sum_cells outer(v dt, wipGrad)) * FE.
 * The reference sums over ALL used cells in lexicographic order; cells outside the 4^3 support add +-0. */
static void trial_FE(const Oracle* o, const int* rank, int pi, const float* vel, float dt, float FEt[9]) {
    const OParticle* P = &o->p[pi];
    float A[9];
    memcpy(A, ID3, sizeof A);
    const int* c = &o->cell[pi * 3];
    if (o->valid[pi])
        for (int di = -2; di <= 2; ++di) for (int dj = -2; dj <= 2; ++dj) for (int dk = -2; dk <= 2; ++dk) {      /* lexicographic = used_cells order */
            const int ni = c[0] + di, nj = c[1] + dj, nk = c[2] + dk;
            const int u = rank[nidx(o, ni, nj, nk)];
            if (u < 0) continue;
            float g[3];
            wip_grad(o, P->pos, ni, nj, nk, g);
            const float vt[3] = { vel[u * 3 + 0] * dt, vel[u * 3 + 1] * dt, vel[u * 3 + 2] * dt };
            for (int col = 0; col < 3; ++col) for (int r = 0; r < 3; ++r) A[col * 3 + r] += vt[r] * g[col];       /* outerProduct(c, r): m[col][row] = c[row] * r[col] */
        }
    m3mul(FEt, A, P->FE);
}
/* ElasticPlasticEnergyDensity, cpp:186-209 */
static float energy_density(const OracleImplicitParams* q, const float FE[9], const float FP[9]) {
    const float e = hardening_factor(q, m3det(FP));
    const float mu = q->mu0 * e, lambda = q->lambda0 * e;
    float R[9], D[9];
    oracle_polar_rotation(FE, R);
    m3sub(D, FE, R);
    double val = 0.0;
    for (int i = 0; i < 9; ++i) val += D[i] * D[i];
    const float fn2 = (float)val;
    const float JE = m3det(FE);
    return mu * fn2 + lambda / 2 * (JE - 1) * (JE - 1);
}
float oracle_energy(Oracle* o, const float* vel, float dt, const OracleImplicitParams* q) {      /* cpp:160-185 */
    double energy = 0.0;
    for (int u = 0; u < o->nused; ++u) {
        const OCell* c = &o->g[o->used[u]];
        const float d[3] = { vel[u * 3] - c->vel[0], vel[u * 3 + 1] - c->vel[1], vel[u * 3 + 2] - c->vel[2] };
        const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        energy += 0.5 * c->mass * n * n;
    }
    int* rank = used_rank(o);
    float acc = 0.0f;
    for (int i = 0; i < o->n; ++i) {
        float FEt[9];
        trial_FE(o, rank, i, vel, dt, FEt);
        acc += o->p[i].volume * energy_density(q, FEt, o->p[i].FP);
    }
    free(rank);
    energy += acc;
    return (float)energy;
}
/* analytic gradient in double: dE/dv_i = m_i (v_i - v*_i) + dt sum_p V_p (2 mu (F - R) + lambda (J - 1) J F^-T) FE_p^T grad w_ip */
void oracle_energy_gradient(Oracle* o, const float* vel, float dt, const OracleImplicitParams* q, double* grad) {
    for (int u = 0; u < o->nused; ++u) {
        const OCell* c = &o->g[o->used[u]];
        for (int a = 0; a < 3; ++a) grad[u * 3 + a] = (double)c->mass * ((double)vel[u * 3 + a] - (double)c->vel[a]);
    }
    int* rank = used_rank(o);
    for (int i = 0; i < o->n; ++i) {
        if (!o->valid[i]) continue;
        const OParticle* P = &o->p[i];
        float FEt[9], R[9], Finv[9];
        trial_FE(o, rank, i, vel, dt, FEt);
        const double e = (double)hardening_factor(q, m3det(P->FP));
        const double mu = q->mu0 * e, lambda = q->lambda0 * e;
        oracle_polar_rotation(FEt, R);
        m3inverse(Finv, FEt);
        const double J = m3det(FEt);
        double Pk[9], G[9];                                   /* column-major like everything else */
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
            Pk[c * 3 + r] = 2.0 * mu * ((double)FEt[c * 3 + r] - (double)R[c * 3 + r]) + lambda * (J - 1.0) * J * (double)Finv[r * 3 + c];   /* F^-T (r,c) = F^-1 (c,r) */
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {                                                                           /* G = V Pk FE^T */
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += Pk[k * 3 + r] * (double)P->FE[k * 3 + c];
            G[c * 3 + r] = (double)P->volume * s;
        }
        const int* c3 = &o->cell[i * 3];
        for (int di = -2; di <= 2; ++di) for (int dj = -2; dj <= 2; ++dj) for (int dk = -2; dk <= 2; ++dk) {
            const int ni = c3[0] + di, nj = c3[1] + dj, nk = c3[2] + dk;
            const int u = rank[nidx(o, ni, nj, nk)];
            if (u < 0) continue;
            float g[3];
            wip_grad(o, P->pos, ni, nj, nk, g);
            for (int r = 0; r < 3; ++r) grad[u * 3 + r] += (double)dt * (G[0 + r] * g[0] + G[3 + r] * g[1] + G[6 + r] * g[2]);
        }
    }
    free(rank);
}
/* Problem::gradient as the optimiser sees it: value + gradient (Problem.hpp:49-52, 87-107) */
static float objective_gradient(Oracle* o, const float* x, float dt, const OracleImplicitParams* q, float* grad, int n, oracle_objective_fn fn, void* user) {
    if (fn) return fn(user, x, grad, n);
    if (q->gradient == 1) {
        const float eps = 2.2204e-6f;
        float* xx = (float*)malloc((size_t)n * sizeof(float));
        for (int d = 0; d < n; ++d) {
            memcpy(xx, x, (size_t)n * sizeof(float));
            float gd = 0.0f;
            xx[d] = x[d] + 1 * eps; gd += 1 * oracle_energy(o, xx, dt, q);
            xx[d] = x[d] + -1 * eps; gd += -1 * oracle_energy(o, xx, dt, q);
            grad[d] = gd / (2 * eps);
        }
        free(xx);
    } else {
        double* g = (double*)malloc((size_t)n * sizeof(double));
        oracle_energy_gradient(o, x, dt, q, g);
        for (int d = 0; d < n; ++d) grad[d] = (float)g[d];
        free(g);
    }
    return oracle_energy(o, x, dt, q);
}
static float objective_value(Oracle* o, const float* x, float dt, const OracleImplicitParams* q, int n, oracle_objective_fn fn, void* user) {
    if (fn) return fn(user, x, NULL, n);
    return oracle_energy(o, x, dt, q);
}
static double vdot(const float* a, const float* b, int n) { double s = 0.0; for (int i = 0; i < n; ++i) s += (double)a[i] * (double)b[i]; return s; }
/* mcl::optlib::LBFGS<float, Dynamic, 8>::minimize (LBFGS.hpp:52-152) with Backtracking::search (Backtracking.hpp:38-69) and
 * Objective::converged (mathy.hpp:31-35). fn != NULL minimises a caller-supplied function instead of Energy (known-answer
 * tests of the optimiser restatement against the vendored library). Returns the iteration count, -1 on line-search failure. */
int oracle_lbfgs(Oracle* o, float* x, int n, float dt, const OracleImplicitParams* q, oracle_objective_fn fn, void* user, int* n_evals) {
    enum { M = 8 };
    float* s = (float*)calloc((size_t)n * M, sizeof(float));
    float* y = (float*)calloc((size_t)n * M, sizeof(float));
    float *grad = (float*)calloc(n, sizeof(float)), *qv = (float*)calloc(n, sizeof(float)), *grad_old = (float*)calloc(n, sizeof(float)),
          *x_old = (float*)calloc(n, sizeof(float)), *x_last = (float*)calloc(n, sizeof(float)), *tmp = (float*)calloc(n, sizeof(float)),
          *lsgrad = (float*)calloc(n, sizeof(float));
    float alpha[M] = { 0 }, rho[M] = { 0 };
    int evals = 0, result = 0;
    objective_gradient(o, x, dt, q, grad, n, fn, user); ++evals;
    float gamma_k = 1.0f;
    float alpha_init = 1.0f;
    int global_iter = 0, max_iters = q->max_iters;
    for (int k = 0; k < max_iters; ++k) {
        memcpy(x_old, x, (size_t)n * sizeof(float));
        memcpy(grad_old, grad, (size_t)n * sizeof(float));
        memcpy(qv, grad, (size_t)n * sizeof(float));
        global_iter++;
        const int iter = k < M ? k : M;
        for (int i = iter - 1; i >= 0; --i) {
            rho[i] = (float)(1.0 / vdot(s + (size_t)i * n, y + (size_t)i * n, n));
            alpha[i] = rho[i] * (float)vdot(s + (size_t)i * n, qv, n);
            for (int d = 0; d < n; ++d) qv[d] = qv[d] - alpha[i] * y[(size_t)i * n + d];
        }
        for (int d = 0; d < n; ++d) qv[d] = gamma_k * qv[d];
        for (int i = 0; i < iter; ++i) {
            const float beta = rho[i] * (float)vdot(qv, y + (size_t)i * n, n);
            for (int d = 0; d < n; ++d) qv[d] = qv[d] + (alpha[i] - beta) * s[(size_t)i * n + d];
        }
        const float dir = (float)vdot(qv, grad, n);
        if (dir <= 0) {
            memcpy(qv, grad, (size_t)n * sizeof(float));
            max_iters -= k;
            k = 0;
            float ginf = 0.0f;
            for (int d = 0; d < n; ++d) ginf = fmaxf(ginf, fabsf(grad[d]));
            alpha_init = (float)fmin(1.0, 1.0 / (double)ginf);
        }
        /* Backtracking::search(x, p = -q) */
        float rate;
        {
            double pn = 0.0;
            for (int d = 0; d < n; ++d) pn += (double)qv[d] * (double)qv[d];
            if ((float)sqrt(pn) <= FLT_EPSILON) rate = q->ls_decrease;
            else {
                float a = alpha_init;
                const float fx0 = objective_gradient(o, x, dt, q, lsgrad, n, fn, user); ++evals;
                const float gtp = -(float)vdot(lsgrad, qv, n);
                int it = 0;
                for (; it < q->ls_max_iters; ++it) {
                    for (int d = 0; d < n; ++d) tmp[d] = x[d] + a * -qv[d];
                    const float fxa = objective_value(o, tmp, dt, q, n, fn, user); ++evals;
                    const float bound = fx0 + a * q->ls_decrease * gtp;
                    if (fxa <= bound) break;
                    a *= q->ls_tau;
                }
                rate = it >= q->ls_max_iters ? -1.0f : a;
            }
        }
        if (rate <= 0) { result = -1; break; }
        memcpy(x_last, x, (size_t)n * sizeof(float));
        for (int d = 0; d < n; ++d) x[d] -= rate * qv[d];
        {   /* Objective::converged(x_last, x, grad) */
            double gn = 0.0, sn = 0.0;
            for (int d = 0; d < n; ++d) { gn += (double)grad[d] * grad[d]; const double dd = (double)x_last[d] - (double)x[d]; sn += dd * dd; }
            if ((float)sqrt(gn) < q->tol_grad || (float)sqrt(sn) < q->tol_step) { result = global_iter; break; }
        }
        objective_gradient(o, x, dt, q, grad, n, fn, user); ++evals;
        float *st = s + (size_t)(k < M ? k : M - 1) * n, *yt = y + (size_t)(k < M ? k : M - 1) * n;
        if (k >= M) {
            memmove(s, s + n, (size_t)n * (M - 1) * sizeof(float));
            memmove(y, y + n, (size_t)n * (M - 1) * sizeof(float));
        }
        for (int d = 0; d < n; ++d) { st[d] = x[d] - x_old[d]; yt[d] = grad[d] - grad_old[d]; }
        const float denom = (float)vdot(yt, yt, n);
        if (fabsf(denom) <= 0) { result = global_iter; break; }
        gamma_k = (float)vdot(st, yt, n) / denom;
        alpha_init = 1.0f;
        result = global_iter;
    }
    if (n_evals) *n_evals = evals;
    free(s); free(y); free(grad); free(qv); free(grad_old); free(x_old); free(x_last); free(tmp); free(lsgrad);
    return result;
}
/* timeIntegration, cpp:211-233: minimise Energy over the velocities of the used cells, starting from the grid's */
int oracle_time_integration(Oracle* o, float dt, const OracleImplicitParams* q, int* n_evals) {
    const int n = o->nused * 3;
    float* x = (float*)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    for (int u = 0; u < o->nused; ++u) for (int a = 0; a < 3; ++a) x[u * 3 + a] = o->g[o->used[u]].vel[a];
    const int it = oracle_lbfgs(o, x, n, dt, q, NULL, NULL, n_evals);
    for (int u = 0; u < o->nused; ++u) for (int a = 0; a < 3; ++a) o->g[o->used[u]].vel[a] = x[u * 3 + a];
    free(x);
    return it;
}

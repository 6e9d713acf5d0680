/* TEST INFRASTRUCTURE — CPU restatement of the reference MPM substep (see mpm_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (libmpm_b200.so) never does. */
#ifndef MPM_ORACLE_H
#define MPM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float h;            /* grid spacing, material_point_method.cpp:17 (0.05) */
    float E;            /* Young's modulus, material_point_method.cpp:238 (1.4e5) */
    float nu;           /* Poisson ratio, :237 (0.2) */
    float xi;           /* hardening, material_point_method.hpp:213 (10) */
    float theta_c;      /* critical compression, material_point_method.cpp:320 (2.5e-2) */
    float theta_s;      /* critical stretch, :320 (5.0e-3) */
    float gravity[3];   /* :260 (0,-9.8,0) */
    float friction;     /* :288 (0.5) */
    int stencil;        /* 0 = the reference's cubic B-spline (hpp:20-31); 1 = quadratic B-spline, NOT reference behaviour: the
                           checker for MpmParams.stencil = 1 (same loops, other weight function and D = h^2/4) */
} OracleParams;

/* Same POD as MpmBoxCollider in include/mpm_b200.h. world_to_local is the glm column-major
 * inverse(translate(t) * toMat4(q)) of material_point_method.hpp:82, computed by the caller. */
typedef struct {
    float world_to_local[16];
    float half_extent[3];
    float velocity[3];
} OracleBoxCollider;

typedef struct Oracle Oracle;

void oracle_default_params(OracleParams* p);
Oracle* oracle_create(int I, int J, int K, int n, const OracleParams* p);
void oracle_destroy(Oracle* o);
void oracle_set_threads(Oracle* o, int threads);   /* 1 = sequential, bit-faithful summation order */

/* particle state: n x 35 floats = mass, vel[3], volume, pos[3], FE[9], FP[9], B[9]; 3x3 blocks are
 * glm column-major (m[col][row]) exactly like oracle/ref_driver.cpp's dumps */
void oracle_set_particles(Oracle* o, const float* state35);
void oracle_get_particles(const Oracle* o, float* state35);
/* grid: I*J*K x 7 floats = mass, force[3], velocity[3] at node index i*J*K + j*K + k */
void oracle_get_grid(const Oracle* o, float* grid7);
void oracle_set_grid(Oracle* o, const float* grid7);   /* stage-isolation tests */
int oracle_num_used_cells(const Oracle* o);
/* per-particle cell index int(pos/h) per axis (n x 3 int32), material_point_method.cpp:83 */
void oracle_cell_indices(const Oracle* o, int* cells3);
/* returns the number of particles whose 5^3 neighbourhood leaves the grid (the reference would
 * index out of bounds, material_point_method.cpp:101); such particles are skipped by the oracle */
int oracle_num_out_of_grid(const Oracle* o);

/* one entry per reference stage, main.cpp:192-218 */
void oracle_rasterize_particles_to_grid(Oracle* o);
void oracle_compute_particle_volumes_and_densities(Oracle* o);
void oracle_compute_explicit_grid_forces(Oracle* o);
void oracle_grid_velocities_update(Oracle* o, float dt);
void oracle_grid_based_collisions(Oracle* o, float dt, const OracleBoxCollider* c, int nc);
int  oracle_update_deformation_gradient(Oracle* o, float dt);
void oracle_update_particle_velocities(Oracle* o);
void oracle_update_particle_positions(Oracle* o, float dt);
void oracle_substep(Oracle* o, float dt, const OracleBoxCollider* c, int nc, int nsteps);

/* function-level known-answer helpers */
float oracle_weight(float x);                                   /* material_point_method.hpp:20-31 */
int   oracle_svd3(const float A[9], float U[9], float S[3], float V[9]); /* Eigen JacobiSVD, row-major 3x3 */
void  oracle_polar_rotation(const float F[9], float R[9]);      /* glm column-major in/out */
float oracle_box_sdf(const OracleBoxCollider* c, const float pos[3]);     /* hpp:79-86 */
void  oracle_body_collision(const float pos[3], const float vel[3], const OracleBoxCollider* c, int nc,
                            float friction, float out[3]);      /* cpp:264-296 */

/* ---- implicit (optimisation-based) time integration: material_point_method.cpp:160-233, mathy.hpp:10-38, vendored
 * mcloptlib (LBFGS.hpp, Backtracking.hpp, Problem.hpp). Dead code in the reference (never called); see mpm_oracle.c. ---- */
typedef struct {
    float mu0, lambda0, xi;   /* hpp:212-214: 1, 10, 1 (placeholders in the reference) */
    int hardening;            /* 0 = exp(xi*1 - det FP) as written (cpp:187-191); 1 = exp(xi*(1 - det FP)) like the explicit forces (cpp:240) */
    int max_iters;            /* 50 (LBFGS.hpp:48) */
    float ls_decrease, ls_tau; int ls_max_iters;   /* 1e-4, 0.7, 100000 (Minimizer.hpp:62-63, Backtracking.hpp:46) */
    float tol_grad, tol_step; /* 1e-2 each (mathy.hpp:31-35) */
    int gradient;             /* 0 = analytic derivative (double); 1 = the library's float central differences, eps 2.2204e-6 */
} OracleImplicitParams;
void  oracle_default_implicit_params(OracleImplicitParams* q);
float oracle_weight_derivative(float x);                              /* hpp:32-52 */
void  oracle_used_cells(const Oracle* o, int* node_index);            /* used_cells (cpp:105-110) as i*J*K + j*K + k */
/* vel: 3 floats per used cell in used_cells order */
float oracle_energy(Oracle* o, const float* vel, float dt, const OracleImplicitParams* q);                     /* cpp:160-209 */
void  oracle_energy_gradient(Oracle* o, const float* vel, float dt, const OracleImplicitParams* q, double* grad);
/* value (and gradient when grad != NULL) of a caller-supplied objective, for known-answer tests of the optimiser */
typedef float (*oracle_objective_fn)(void* user, const float* x, float* grad, int n);
int   oracle_lbfgs(Oracle* o, float* x, int n, float dt, const OracleImplicitParams* q, oracle_objective_fn fn, void* user, int* n_evals);
int   oracle_time_integration(Oracle* o, float dt, const OracleImplicitParams* q, int* n_evals);              /* cpp:211-233 */

#ifdef __cplusplus
}
#endif
#endif

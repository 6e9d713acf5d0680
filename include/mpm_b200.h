/* mpm_b200.h — C ABI of libmpm_b200.so, the B200-native (sm_100a) MPM substep.
 *
 * Drop-in boundary for the hot path of brocbyte/realtime-deformations: every entry point below replaces one
 * member of MaterialPointMethod::LagrangeEulerView (reference realtime-deformations/material_point_method.hpp:
 * 174-233, definitions in material_point_method.cpp) as called by the viewer loop (main.cpp:48-54, 192-218).
 * Plain pointers and sizes only; no C++/torch types. All functions return MPM_OK (0) or a negative MpmStatus;
 * mpm_last_error() returns the message of the last failure on the calling thread. The reference itself has no
 * error convention (void methods that print "cudaMalloc error", hpp:125-127) — a C++ adapter that keeps those
 * signatures is in adapter/lagrange_euler_view_b200.cpp, the binding recipe in INTEGRATION.md.
 *
 * Conventions
 *   - 3x3 matrices cross the ABI as 9 floats in glm column-major order (m[c*3+r] == glm m[c][r]), i.e. exactly
 *     the bytes of the reference's glm::mat3 members (Particle::FElastic/FPlastic/B, hpp:101-103).
 *   - Grid nodes are addressed as in Grid::operator() (hpp:129-131): node = i*MAX_J*MAX_K + j*MAX_K + k.
 *   - One handle per GPU (the device current at mpm_create). Handles are not thread-safe; distinct handles are
 *     independent. All work is issued on a library-owned non-blocking stream (or the one given to
 *     mpm_set_stream); every download_* / get_* call synchronises that stream.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with MPM_ERR_CUDA.
 */
#ifndef MPM_B200_H
#define MPM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum MpmStatus {
    MPM_OK = 0,
    MPM_ERR_INVALID = -1,     /* bad argument / call order */
    MPM_ERR_CUDA = -2,        /* CUDA runtime failure (message has cudaGetErrorString) */
    MPM_ERR_CAPACITY = -3,    /* particle capacity exceeded (slab migration) */
    MPM_ERR_SVD = -4          /* non-finite deformation gradient met by the F-update (reference: cpp:313-316) */
} MpmStatus;

/* Material / scene constants the reference hard-codes; defaults (mpm_default_params) are its values. */
typedef struct MpmParams {
    float h;              /* grid spacing            material_point_method.cpp:17   (0.05)   */
    float youngs_modulus; /* E                       material_point_method.cpp:238  (1.4e5)  */
    float poisson_ratio;  /* nu                      material_point_method.cpp:237  (0.2)    */
    float hardening_xi;   /* xi                      material_point_method.hpp:213  (10)     */
    float theta_c;        /* critical compression    material_point_method.cpp:320  (2.5e-2) */
    float theta_s;        /* critical stretch        material_point_method.cpp:320  (5.0e-3) */
    float gravity[3];     /*                         material_point_method.cpp:260  (0,-9.8,0) */
    float friction_mu;    /*                         material_point_method.cpp:288  (0.5)    */
    int   p2g_variant;    /* 0 = auto: block-tile kernel (packed fp32 pairs); inside the fused mpm_substep it also runs the F-update,
                             1 = per-particle global atomics (debug / baseline), 2 = block-tile kernel with the F-update as a
                             kernel of its own (A/B), 9 = deterministic debug mode: one thread adds the contributions in
                             ascending particle-id order without atomics, so a run is bitwise reproducible (small scenes;
                             single-domain handles) */
    int   g2p_variant;    /* 0 = auto: warp-per-block gather from a TMA-staged linear tile with packed fp32 pairs,
                             1 = direct global gathers (debug / baseline) */
    int   fupdate_exact;  /* F-update (cpp:306-330) inside the fused mpm_substep: 0 = tolerance form (FMA contraction, MUFU
                             reciprocals / square roots, F^ = (I + dt C) FE taken directly; same Eigen Jacobi control flow;
                             held to 4x the reference's own FMA-contraction noise floor by the trajectory tests),
                             1 = the bit-faithful form everywhere, 2 = the tolerance form everywhere (tests). With 0 the staged
                             mpm_update_deformation_gradient stays bit-faithful. */
    int   stencil;        /* 0 = cubic B-spline, the reference's (hpp:20-31; the parity default); 1 = quadratic B-spline: three nodes
                             per axis instead of four, D = h^2/4 -- not reference behaviour, checked against
                             the oracle's own quadratic mode; the tile kernels run W = 3 instantiations */
    int   reserved[4];
} MpmParams;

/* Box collider = MeshCollider (hpp:74-93) reduced to what its sdf lambda uses (hpp:80-85):
 * world_to_local = inverse(translate(mesh.translation) * toMat4(mesh.rotation)) as a glm column-major mat4,
 * half_extent = mesh.scale, velocity = MeshCollider::velocity. The caller computes the matrix on the host
 * exactly as the reference does, so the device evaluates the same numbers. */
typedef struct MpmBoxCollider {
    float world_to_local[16];
    float half_extent[3];
    float velocity[3];
} MpmBoxCollider;
#define MPM_MAX_COLLIDERS 16

/* Scene front-end for hosts without glm: the pose a MeshCollider's sdf lambda reads (hpp:80-83) --
 * mesh.scale, mesh.rotation (glm::quat, stored w x y z here), mesh.translation -- and MeshCollider::velocity. */
typedef struct MpmBoxTransform {
    float scale[3];
    float rotation_wxyz[4];
    float translation[3];
    float velocity[3];
} MpmBoxTransform;

typedef struct MpmStats {
    int64_t n_particles;        /* live particles on this handle */
    int64_t n_out_of_grid;      /* particles whose stencil leaves the grid (reference would index out of bounds) */
    int64_t n_active_nodes;     /* nodes with mass != 0 == used_cells.size() (cpp:105-110), after the last P2G */
    int64_t n_particle_blocks;  /* occupied 4^3-cell particle blocks after the last binning */
    int64_t n_grid_blocks;      /* active 4^3-node grid blocks after the last binning */
    int64_t substeps_done;
    int64_t kernel_launches;    /* CUDA kernels launched by this handle since creation */
    float   last_ms[8];         /* device ms of the last mpm_substep call: bin, clear, p2g, grid, g2p (F-update + gather),
                                   time between begin/end calls, total, F-update alone (-1 if it ran on the side stream) */
    int32_t svd_failed;         /* 1 if the last F-update met a non-finite matrix */
    int32_t reserved[7];        /* [0] = 1 if the pos/h shortcut passed its exhaustive check for this h,
                                   [1] = 1 if a migration buffer overflowed (particles kept one more substep),
                                   [2] = 1 if a peer-memory halo wait gave up (mpm_substep_begin_peer): FATAL for the run -- the flag stays set,
                                       later waits return at once and mpm_sync_counts keeps failing with MPM_ERR_CUDA */
} MpmStats;

typedef struct mpm_sim mpm_t;

void        mpm_default_params(MpmParams* p);
const char* mpm_last_error(void);
int         mpm_device_count(void);                 /* 0 when no CUDA device is usable */

/* LagrangeEulerView::LagrangeEulerView(max_i, max_j, max_k, particlesNum)  (cpp:144-154, hpp:178) */
int mpm_create(const MpmParams* params, int max_i, int max_j, int max_k, int64_t n_particles, mpm_t** out);
/* Slab-decomposed variant (no reference analogue): this handle owns the particle-block layers
 * [block_lo, block_hi) along i (a layer = 4 cells) of the global max_i x max_j x max_k grid and
 * can hold up to `capacity` particles. mpm_create == mpm_create_slab(.., 0, all layers, n_particles). */
int mpm_create_slab(const MpmParams* params, int max_i, int max_j, int max_k, int block_lo, int block_hi,
                    int64_t n_particles, int64_t capacity, mpm_t** out);
/* LagrangeEulerView::~LagrangeEulerView()  (cpp:156-158) */
int mpm_destroy(mpm_t* s);
int mpm_set_stream(mpm_t* s, void* cuda_stream);   /* cudaStream_t; NULL restores the library-owned stream */
int mpm_set_params(mpm_t* s, const MpmParams* params);  /* material sweep (BASELINE config 4); h must not change */

/* Particle upload. Replaces initializeParticles()'s writes into std::vector<Particle> (cpp:42-53) and any
 * later host-side edit through getParticles() (hpp:181-183).
 *  - AoS: `particles` points at an array of the reference's own struct Particle (hpp:96-110), `stride` =
 *    sizeof(Particle); the off_* arguments are offsetof() of mass, velocity, volume, pos, FElastic, FPlastic, B.
 *  - SoA: n x 3 / n x 9 row arrays; FE/FP/B/volume may be NULL (identity / zero / computed by
 *    mpm_compute_particle_volumes_and_densities). All host pointers. */
int mpm_upload_particles_aos(mpm_t* s, const void* particles, int64_t n, size_t stride, size_t off_mass,
                             size_t off_velocity, size_t off_volume, size_t off_pos, size_t off_FE, size_t off_FP,
                             size_t off_B);
int mpm_upload_particles_soa(mpm_t* s, int64_t n, const float* pos, const float* vel, const float* mass,
                             const float* volume, const float* FE, const float* FP, const float* B);
int mpm_download_particles_aos(mpm_t* s, void* particles, int64_t n, size_t stride, size_t off_mass,
                               size_t off_velocity, size_t off_volume, size_t off_pos, size_t off_FE, size_t off_FP,
                               size_t off_B);
int mpm_download_particles_soa(mpm_t* s, int64_t n, float* pos, float* vel, float* mass, float* volume, float* FE,
                               float* FP, float* B);
/* What drawParticles() copies out of getParticles() each frame (main.cpp:257-271): xyz + size, rgba. Either
 * pointer may be NULL. Particle order = upload order. */
int mpm_download_render_buffers(mpm_t* s, int64_t n, float* xyzs, unsigned char* rgba, float size);

/* Pipelined variant for a render loop: enqueues the instance-buffer build behind the work already issued and its
 * device->host copy on a separate copy stream, and returns at once, so the copy of frame t overlaps the substeps of
 * frame t+1 (the copy engine and the SMs run concurrently). xyzs must be page-locked host memory and must not be
 * read before mpm_wait_render_buffers returns. */
int mpm_download_render_buffers_async(mpm_t* s, int64_t n, float* xyzs_pinned, float size);
int mpm_wait_render_buffers(mpm_t* s);

/* Viewer hand-off without the host round trip (replaces the copy loop + glBufferSubData of
 * main.cpp:255-286): writes the same instance data straight into DEVICE memory owned by the caller, e.g. the pointer
 * cudaGraphicsResourceGetMappedPointer returns for the viewer's GL instance buffers. d_xyzs: n float4 (x, y, z, size);
 * d_rgba: n uchar4 (255,255,255,255) or NULL. Asynchronous on the handle's stream (mpm_synchronize before unmapping). */
int mpm_write_render_buffers_device(mpm_t* s, int64_t n, void* d_xyzs, void* d_rgba, float size);

/* One entry per reference stage (same order and meaning as main.cpp:192-218). */
int mpm_rasterize_particles_to_grid(mpm_t* s);                   /* rasterizeParticlesToGrid        cpp:94-129  */
int mpm_compute_particle_volumes_and_densities(mpm_t* s);        /* computeParticleVolumesAndDensities cpp:131-142 */
int mpm_compute_explicit_grid_forces(mpm_t* s);                  /* computeExplicitGridForces       cpp:235-254 */
int mpm_grid_velocities_update(mpm_t* s, float dt);              /* gridVelocitiesUpdate            cpp:256-262 */
int mpm_grid_based_collisions(mpm_t* s, float dt, const MpmBoxCollider* colliders, int n_colliders);
                                                                 /* gridBasedCollisions + bodyCollision cpp:264-304 */
int mpm_update_deformation_gradient(mpm_t* s, float dt);         /* updateDeformationGradient       cpp:306-330 */
int mpm_update_particle_velocities(mpm_t* s);                    /* updateParticleVelocities        cpp:332-342 */
int mpm_update_particle_positions(mpm_t* s, float dt);           /* updateParticlePositions         cpp:344-350 */

/* ---- implicit (optimisation-based) time integration ----
 * LagrangeEulerView::timeIntegration (material_point_method.cpp:211-233): between mpm_grid_velocities_update and
 * mpm_grid_based_collisions, replace the grid velocities v* of the used cells by the minimiser of
 *     Energy(v) = sum_i 1/2 m_i |v_i - v*_i|^2 + sum_p V_p psi((I + dt sum_i v_i grad w_ip^T) FE_p, FP_p)      cpp:160-209
 * found by the optimiser mathy.hpp:10-38 configures (external/mcloptlib: L-BFGS, history 8, <= 50 iterations, Armijo
 * backtracking 0.7 / 1e-4, stop when |grad| < 1e-2 or |step| < 1e-2). The reference never calls this path
 * (README.md:17: TODO) and its defaults are placeholders (mu0 = lambda0 = 1, hardening exp(xi*1 - det FP)): they are the
 * defaults here too, for parity of the objective; pass E/(2(1+nu)), E nu/((1+nu)(1-2nu)) and hardening = 1 for the
 * material of the explicit path. Deviation: the reference differentiates its float objective by central differences of
 * 2.2e-6 (rounding noise; Problem.hpp:87-107); this library uses the analytic gradient of the same objective.
 * Single-domain handles, cubic stencil; needs a binned handle (mpm_rasterize_particles_to_grid). */
typedef struct MpmImplicitParams {
    float mu0, lambda0, xi;   /* material_point_method.hpp:212-214  (1, 1, 10) */
    int   hardening;          /* 0 = exp(xi*1 - det FP) as written (cpp:187-191); 1 = exp(xi*(1 - det FP)) as the explicit forces (cpp:240) */
    int   max_iters;          /* LBFGS.hpp:48 (50) */
    float ls_decrease, ls_tau; /* Minimizer.hpp:63 (1e-4), Backtracking.hpp:46 (0.7) */
    int   ls_max_iters;       /* Minimizer.hpp:62 (100000) */
    float tol_grad, tol_step; /* mathy.hpp:31-35 (1e-2, 1e-2) */
    int   reserved[4];
} MpmImplicitParams;
typedef struct MpmImplicitStats {
    int    iterations;        /* what LBFGS::minimize returns; -1 = line-search failure */
    int    evaluations;       /* objective evaluations (each one pass over the particles) */
    double energy_start, energy_end, grad_norm_end;
    int    reserved[4];
} MpmImplicitStats;
void mpm_default_implicit_params(MpmImplicitParams* q);
int mpm_time_integration(mpm_t* s, float dt, const MpmImplicitParams* q, MpmImplicitStats* stats /* may be NULL */);
/* Diagnostics (parity checks): Energy / its gradient at a trial velocity field given as a dense host array of
 * I*J*K x 3 floats in node order i*J*K + j*K + k (NULL = the grid's own velocities); relative != 0: the array is added to
 * the grid's velocities. energy = the whole objective, elastic (may be NULL) = the ElasticPotential part (cpp:173-185).
 * Parked (out-of-grid) particles are not part of the sums. gradient: I*J*K x 3 floats out, zero at nodes without mass. */
int mpm_energy(mpm_t* s, float dt, const MpmImplicitParams* q, const float* trial_velocity, int relative, double* energy, double* elastic);
int mpm_energy_gradient(mpm_t* s, float dt, const MpmImplicitParams* q, const float* trial_velocity, int relative, float* gradient);

/* The fused fast path: n_substeps repetitions of the seven stages above in main.cpp's order, as
 * bin/sort -> clear -> P2G (mass + APIC momentum + stress) -> grid update (velocity solve, gravity, collisions)
 * -> G2P (F-update with the previous B, plasticity, gather, advect, re-sort). */
int mpm_substep(mpm_t* s, float dt, const MpmBoxCollider* colliders, int n_colliders, int n_substeps);

/* Test / diagnostics access. grid7: max_i*max_j*max_k x 7 floats = mass, force[3], velocity[3] (struct Cell,
 * hpp:112-117, without nParticles). cells3: n x 3 int32 = ivec3(pos / h) (cpp:83), upload order. block_key: n
 * int32 = linear particle-block id the binning stage assigned (-1 = out of grid). */
int mpm_download_grid(mpm_t* s, float* grid7);
int mpm_upload_grid(mpm_t* s, const float* grid7);               /* stage-isolation tests */
int mpm_download_binning(mpm_t* s, int64_t n, int32_t* cells3, int32_t* block_key, int32_t* sorted_ids);
int mpm_get_stats(mpm_t* s, MpmStats* out);
int mpm_synchronize(mpm_t* s);

/* ---- scene front-end: host-only helpers, no handle, no device ----
 * mpm_box_collider_from_transform: world_to_local = inverse(translate(mat4(), translation) * toMat4(rotation)) with glm
 * 0.9.7.1's operation order (gtc/quaternion.inl mat3_cast, gtc/matrix_transform.inl translate, detail/type_mat4x4.inl
 * operator* and compute_inverse), half_extent = scale, velocity copied: bit for bit what the reference's sdf lambda
 * (hpp:80-83) computes from the same pose. mpm_box_transform_move: MeshCollider::move (hpp:90-92), i.e.
 * applyMatrix4(translate(dt * velocity)) followed by glm::decompose (mesh.hpp:20-23): translation += dt * velocity,
 * scale and rotation unchanged. mpm_box_transform_flip_velocity: the key_callback case (main.cpp:37-41,174-180).
 * (In the reference the boxes pushed into solidObjects are copies whose sdf lambdas still read the ORIGINAL objects, so
 * its collisions never see a move; through this ABI a collider is wherever the host's transform says it is.)
 * mpm_fill_ball: LagrangeEulerView::initializeParticles (cpp:18-63) with the ball radius as a parameter (the reference
 * hard-codes 0.2): every cell of the cube of +-int(radius/h) cells around ivec3(origin/h) gets 8 candidate sites
 * (cell + {1/4,3/4}^3 + generateRandomInsideUnitBall(0.25), utils.h:114-127) * h, kept if within `radius` of the origin.
 * Accepted positions are written in acceptance order (the reference stores the k-th one in slot nParticles-1-k); once
 * `capacity` positions are stored further candidates are only counted in *n_missing (its "k more!!!"), and -- as in the
 * reference -- three more random numbers are drawn per STORED particle (its discarded colour). rnd has libc rand()'s
 * contract (NULL = rand() itself: with the default seed this is the reference's start-up scene, particle for particle).
 * Several bodies = several calls into one position array. */
typedef int (*MpmRandFn)(void* user);
int mpm_fill_ball(const float origin[3], float radius, float h, MpmRandFn rnd, void* user,
                  float* pos_xyz, int64_t capacity, int64_t* n_written, int64_t* n_missing);
/* Bodies from triangle meshes (the reference ships common/objloader.hpp:4 loadOBJ and never calls it): mpm_load_obj reads the
 * "v" / "f" records of a Wavefront OBJ into n_tri x 9 floats (malloc'ed: release with mpm_free; polygons are fanned);
 * mpm_fill_mesh applies initializeParticles' fill rule (8 jittered sites per cell, the random stream of mpm_fill_ball) to the
 * cells of the mesh's bounding box and keeps the candidates inside the CLOSED mesh (ray parity). */
int mpm_load_obj(const char* path, float** tri_xyz, int64_t* n_tri);
void mpm_free(void* p);
int mpm_fill_mesh(const float* tri_xyz, int64_t n_tri, float h, MpmRandFn rnd, void* user,
                  float* pos_xyz, int64_t capacity, int64_t* n_written, int64_t* n_missing);
/* A second SDF shape through the same collider POD (MeshCollider::sdf is a std::function, hpp:87): sphere of `radius` about
 * `centre`: world_to_local = translate(-centre), half_extent = (radius, -1, -1) -- a negative half_extent[1] marks the sphere;
 * normal, friction rule and velocity handling are the reference's bodyCollision unchanged. */
int mpm_sphere_collider(const float centre[3], float radius, const float velocity[3], MpmBoxCollider* out);
int mpm_box_collider_from_transform(const MpmBoxTransform* t, MpmBoxCollider* out);
int mpm_box_transform_move(MpmBoxTransform* t, float time_delta);
int mpm_box_transform_flip_velocity(MpmBoxTransform* t);

/* ---- slab decomposition plumbing (multi-GPU; the exchange itself is done by the caller, e.g. NCCL) ----
 * Ghost layer: the handle's grid holds one extra block layer above block_hi. After P2G (inside
 * mpm_substep_begin) its partial sums must be added into the upper neighbour's first layer and vice versa. All
 * pointers here are DEVICE pointers owned by the caller; *_bytes() give the sizes. */
size_t mpm_halo_bytes(const mpm_t* s);                           /* one block layer of float4 nodes */
int mpm_halo_pack(mpm_t* s, int upper, void* dev_buf);           /* upper=1: ghost layer (block_hi); 0: first layer (block_lo) */
int mpm_halo_add(mpm_t* s, int upper, const void* dev_buf);      /* add a neighbour's partial sums into that layer */
/* split form of mpm_substep for the exchange points: begin = bin/clear/P2G, end = grid update/G2P */
int mpm_substep_begin(mpm_t* s, float dt);
int mpm_substep_end(mpm_t* s, float dt, const MpmBoxCollider* colliders, int n_colliders);
/* Migration: after mpm_substep_end, particles that left [block_lo, block_hi) sit packed (MPM_MIGRATE_FLOATS
 * floats each) in two device buffers; counts are returned after a stream sync. mpm_migrate_append adds received
 * particles. */
#define MPM_MIGRATE_FLOATS 44
int mpm_migrate_outgoing(mpm_t* s, int64_t* n_down, int64_t* n_up, const void** dev_down, const void** dev_up);
int mpm_migrate_append(mpm_t* s, const void* dev_buf, int64_t n);
/* Sync-free migration (what multi.py uses): mpm_migrate_pack fills two fixed-size device buffers of
 * mpm_migrate_buffer_bytes(s) bytes each = one 16-byte header (first int32 = record count) + capacity records, without
 * any host synchronisation; mpm_migrate_append_packed appends a neighbour's buffer, reading the count on the device.
 * mpm_sync_counts is the occasional host read-back that re-tightens the launch bound and reports overflow. */
int mpm_set_migrate_capacity(mpm_t* s, int64_t records);   /* must be IDENTICAL on neighbouring ranks (fixed-size messages) */
size_t mpm_migrate_buffer_bytes(const mpm_t* s);
int mpm_migrate_pack(mpm_t* s, const void** dev_down, const void** dev_up);
int mpm_migrate_append_packed(mpm_t* s, const void* dev_buf);
int mpm_sync_counts(mpm_t* s);
/* Peer-memory halo (the multi-GPU default of realtime-deformations_b200/multi.py; measured on 2 / 4 / 8 B200s): the
 * ghost-layer reduction done by P2G itself. Every
 * rank exports its grid allocation (mpm_peer_export: a cudaIpcMemHandle_t), the host exchanges the handles and each rank
 * opens its neighbours' (mpm_peer_connect; *_layers = the neighbour's block_hi - block_lo; NULL at the ends of the chain).
 * P2G then adds every tile node of a shared block layer to the local copy and, through NVLink, to the neighbour's copy;
 * mpm_substep_begin_peer(s, dt, 0), (.., 1), (.., 2) called back to back replace mpm_substep_begin + the halo exchange
 * (device-side flags between the phases, no host synchronisation, no message); mpm_substep_end follows unchanged.
 * mpm_peer_connect_ptr takes neighbour grids that live in THIS process (mpm_grid_device_ptr of another handle on the same
 * or a peer-enabled device): used by the single-process test of the protocol. */
#define MPM_IPC_HANDLE_BYTES 64
int mpm_peer_export(mpm_t* s, unsigned char* handle);
int mpm_peer_connect(mpm_t* s, const unsigned char* lower_handle, int lower_layers, const unsigned char* upper_handle, int upper_layers);
int mpm_peer_connect_ptr(mpm_t* s, void* lower_grid, int lower_layers, void* upper_grid, int upper_layers);
int mpm_grid_device_ptr(mpm_t* s, void** grid);
int mpm_substep_begin_peer(mpm_t* s, float dt, int phase);
/* The migration without messages (after mpm_peer_connect): every rank exports its two packed buffers, opens the lower
 * neighbour's UP and the upper neighbour's DOWN buffer, and mpm_migrate_peer(s, 0), (s, 1) called back to back replace
 * mpm_migrate_pack + the exchange + mpm_migrate_append_packed (pull over NVLink, flags "packed" / "consumed"). */
int mpm_peer_export_migration(mpm_t* s, unsigned char* handle_down, unsigned char* handle_up);
int mpm_peer_connect_migration(mpm_t* s, const unsigned char* lower_up_handle, const unsigned char* upper_down_handle);
int mpm_peer_connect_migration_ptr(mpm_t* s, const void* lower_up_buf, const void* upper_down_buf);
int mpm_migrate_peer(mpm_t* s, int phase);
/* Distributed bookkeeping: particle ids are upload indices + pid_base (set before an upload) so that they stay
 * unique across slabs; mpm_download_live_particles returns the handle's current particles in storage order as
 * 35-float rows (mass, vel[3], volume, pos[3], FE[9], FP[9], B[9]) with their ids. */
int mpm_set_pid_base(mpm_t* s, int64_t pid_base);
int mpm_download_live_particles(mpm_t* s, int64_t capacity, int64_t* n_out, float* state35, int32_t* pid);
/* Conserved quantities of this handle's live particles, reduced on the device (the reference's own self-check is the momentum
 * balance it prints, cpp:122-128): fsum5 = { sum m, sum m vx, sum m vy, sum m vz, sum m y } in fp64,
 * isum3 = { count, sum id, sum splitmix64(id) } mod 2^64 (ids = upload indices + pid base: a lost or duplicated particle
 * changes the checksums). Summing the arrays over the slabs of a decomposed run gives the single-domain values. */
int mpm_reduce_invariants(mpm_t* s, double* fsum5, uint64_t* isum3);
/* Tuning aid: clock64 ticks per phase of the block-tile P2G kernel summed over its CTAs; all zero unless the library was
 * built with -DMPM_P2G_PROFILE (tools/p2g_phase_profile.py). */
int mpm_debug_p2g_profile(mpm_t* s, int64_t* ticks8, int reset);
#ifdef __cplusplus
}
#endif
#endif /* MPM_B200_H */

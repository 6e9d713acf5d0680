// Kernels of the B200 MPM substep. See DESIGN.md for the data layout and the per-kernel rooflines.
//
// HBM layout
//   particles : 11 "planes" of float4, plane-major (plane k = cap consecutive float4), double-buffered.
//               P0 x y z m | P1 B0..3 | P2 B4..7 | P3 B8 vx vy vz | P4 tau0..3 | P5 tau4 tau5 - - |
//               P6 V0 pid FE0 FE1 | P7 FE2..5 | P8 FE6 FE7 FE8 FP0 | P9 FP1..4 | P10 FP5..8
//               (m < 0 marks a dead slot; pid = upload index; tau = V0*Dinv*dPsi*FE^T, symmetric, 6 floats)
//   grid      : float4 (mass, momentum|velocity xyz) per node, BLOCKED: 4x4x4-node grid blocks are contiguous
//               1 KB chunks, block index ((bi-lo)*nbj + bj)*nbk + bk, node-in-block ((i&3)*4 + (j&3))*4 + (k&3).
//   binning   : key[slot] (particle-block id), blk_count/blk_start/blk_cursor[n_pblocks+3], sorted_ids[rank] = slot.
// A particle with cell c = int(x/h) belongs to particle block (c-1)>>2 per axis; its 4^3 stencil c-1..c+2 then
// lies inside the 2x2x2 grid blocks starting at that block ("tile").
#pragma once
#include "mpm_math.cuh"
#include <cstdint>

namespace mpm {

constexpr int NPLANES = 11;
struct Planes { float4* p[NPLANES]; };

struct GridDims {
    int I, J, K;            // global node counts (MAX_I, MAX_J, MAX_K)
    int nbj, nbk;           // grid blocks along j,k (= ceil(J/4)+1, ceil(K/4)+1)
    int npbj, npbk;         // particle blocks along j,k (= nbj-1, nbk-1)
    int lo, hi;             // owned particle-block layers [lo, hi) along i
    int npbi_global;        // ceil(I/4)
    int n_pblocks;          // (hi-lo)*npbj*npbk
    int n_gblocks;          // (hi-lo+1)*nbj*nbk
};

struct SimConst {
    float h, dinv, E, nu, xi, clamp_lo, clamp_hi, friction;
    float mu0, lambda0;        // E / (2 (1 + nu)), E nu / ((1 + nu)(1 - 2 nu)): cpp:237-238, hoisted for the tolerance-form F-update
    float g[3];
    float pos_lo, pos_hi[3];   // clampPosition bounds (cpp:381-388)
    int p2g_rotate;            // k_p2g_tile: start each cell's record walk at a different record (bank-conflict fix), 0 = off
    PosDiv pd;                 // pos / h
};

// device-resident counters (one struct per handle); nothing here needs a host sync inside a substep
struct DevCounters {
    int n_slots;            // occupied slots in the current particle buffer (live + dead)
    int n_binned;           // particles in real blocks after the last binning (== blk_start[n_pblocks])
    int n_sorted;           // n_binned + parked (out-of-grid) particles: what the re-sort keeps
    int n_active_pblocks;
    int n_active_gblocks;
    int n_active_nodes;
    int n_out_of_grid;
    int n_out_down, n_out_up;
    int work_a, work_b, work_c;   // dynamic work counters for persistent kernels
    int svd_failed;
    int epoch;
    int n_mig[2];           // particles packed for the lower / upper slab neighbour by k_mark_outgoing
    int mig_overflow;
    int peer_timeout;       // peer-memory halo: a neighbour's flag did not arrive (k_peer_wait gave up)
    int pad[2];
    long long prof[8];      // MPM_P2G_PROFILE builds only: clock64 ticks per P2G phase, summed over CTAs (thread 0)
};

enum { KEY_DEAD = -2 };

// slab-local particle-block coordinates in one int (10 bits each; mpm_create refuses grids beyond 1024 blocks per axis)
constexpr int PB_COORD_BITS = 10, PB_COORD_MAX = 1 << PB_COORD_BITS;
__host__ __device__ inline int pack_block_coords(int pbi_local, int pbj, int pbk) { return (pbi_local << (2 * PB_COORD_BITS)) | (pbj << PB_COORD_BITS) | pbk; }

struct ColliderSet { BoxCollider c[16]; };

// exhaustive check of the pos/h shortcut: every fp32 bit pattern in [lo_bits, hi_bits] against the IEEE intrinsic
__global__ void k_validate_pos_div(PosDiv pd, unsigned lo_bits, unsigned hi_bits, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long b = (unsigned long long)lo_bits + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b <= hi_bits;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        const float q0 = mul_rn(x, pd.rh);
        const float q = __fmaf_rn(__fmaf_rn(-q0, pd.h, x), pd.rh, q0);
        bad += (__float_as_uint(q) != __float_as_uint(div_rn(x, pd.h)));
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------------
template <int Q = 2>
MPM_DI int particle_key(float4 xm, const GridDims& gd, const PosDiv& pd, int* cells /*3*/) {
    if (xm.w < 0.0f) return KEY_DEAD;
    const int cx = cell_of_t<Q>(xm.x, pd), cy = cell_of_t<Q>(xm.y, pd), cz = cell_of_t<Q>(xm.z, pd);
    cells[0] = cx; cells[1] = cy; cells[2] = cz;
    // the reference enumerates cell-2..cell+2 without bounds checks (cpp:84-91); outside that the particle is parked
    const bool ok = cx >= 2 && cy >= 2 && cz >= 2 && cx + 2 <= gd.I - 1 && cy + 2 <= gd.J - 1 && cz + 2 <= gd.K - 1;
    if (!ok) return gd.n_pblocks;                      // parked bucket
    const int pbi = (cx - 1) >> 2, pbj = (cy - 1) >> 2, pbk = (cz - 1) >> 2;
    if (pbi < gd.lo) return gd.n_pblocks + 1;          // leaves the slab downwards
    if (pbi >= gd.hi) return gd.n_pblocks + 2;         // ... upwards
    return ((pbi - gd.lo) * gd.npbj + pbj) * gd.npbk + pbk;
}

// Slab migration fused into the re-sorting gather: a particle whose new block layer left [lo, hi) (key n_pblocks+1 / +2) is
// appended to the packed buffer for that neighbour -- one float4 header whose first int is the record count (it IS the atomic
// cursor, so the receiver learns the count on the device), then 11 float4 per record -- and its slot is retired (m < 0,
// key KEY_DEAD). If the buffer is full the particle simply stays (alive, in its "leaving" bucket, frozen like a parked one)
// and is offered again by the next substep's k_copy_parked.
struct MigOut { float4* dn; float4* up; int cap; };
MPM_DI bool mig_try_pack(const MigOut& mo, bool upward, const Planes& D, int q, DevCounters* dc) {
    float4* buf = upward ? mo.up : mo.dn;
    if (!buf) return false;
    const int idx = atomicAdd(reinterpret_cast<int*>(buf), 1);
    if (idx >= mo.cap) { dc->mig_overflow = 1; return false; }      // header count may exceed cap; the receiver clamps it
    float4* o = buf + 1 + (size_t)idx * NPLANES;
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) o[k] = D.p[k][q];
    const float4 a0 = D.p[0][q];
    D.p[0][q] = make_float4(a0.x, a0.y, a0.z, -1.0f);               // dead slot: skipped by the binning, dropped by the next re-sort
    return true;
}

// Both binning kernels handle BIN_E elements per thread (warp-coalesced, stride = blockDim) so that several
// position loads / atomic round trips are in flight per thread: they are latency-bound, not bandwidth-bound.
constexpr int BIN_E = 4, BIN_T = 256;
__global__ void __launch_bounds__(BIN_T)
k_bin_count(const float4* __restrict__ P0, int n_bound, const DevCounters* __restrict__ dc,
            GridDims gd, PosDiv pd, int* __restrict__ key, int* __restrict__ blk_count) {
    const int n = min(n_bound, dc->n_slots);
    const int base = blockIdx.x * (BIN_T * BIN_E) + threadIdx.x;
    float4 xm[BIN_E];
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) {
        const int j = base + e * BIN_T;
        xm[e] = j < n ? P0[j] : make_float4(0.f, 0.f, 0.f, -1.f);
    }
    int k[BIN_E];
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) {
        const int j = base + e * BIN_T;
        int cells[3];
        k[e] = KEY_DEAD;
        if (j < n) { k[e] = particle_key(xm[e], gd, pd, cells); key[j] = k[e]; }
    }
    // warp-aggregated histogram: one atomic per distinct key per warp
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) {
        const unsigned peers = __match_any_sync(0xffffffffu, k[e]);
        if (k[e] >= 0 && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&blk_count[k[e]], __popc(peers));
    }
}

__global__ void __launch_bounds__(BIN_T)
k_bin_scatter(int n_bound, const DevCounters* __restrict__ dc, const int* __restrict__ key,
              int* __restrict__ blk_cursor, int* __restrict__ sorted_ids) {
    const int n = min(n_bound, dc->n_slots);
    const int base0 = blockIdx.x * (BIN_T * BIN_E) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int k[BIN_E], slot[BIN_E], leader[BIN_E];
    unsigned peers[BIN_E];
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) { const int j = base0 + e * BIN_T; k[e] = j < n ? key[j] : KEY_DEAD; }
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) {               // all BIN_E atomic round trips are issued before any is consumed
        peers[e] = __match_any_sync(0xffffffffu, k[e]);
        leader[e] = __ffs(peers[e]) - 1;
        slot[e] = 0;
        if (k[e] >= 0 && lane == leader[e]) slot[e] = atomicAdd(&blk_cursor[k[e]], __popc(peers[e]));
    }
#pragma unroll
    for (int e = 0; e < BIN_E; ++e) {
        const int b = __shfl_sync(0xffffffffu, slot[e], leader[e]);
        if (k[e] >= 0) sorted_ids[b + __popc(peers[e] & ((1u << lane) - 1u))] = base0 + e * BIN_T;
    }
}

// ---- exclusive scan over the per-block counts (+ compaction of the occupied blocks) ----
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_CHUNK = SCAN_T * SCAN_E;

MPM_DI int2 block_scan_excl(int2 v, int2* total, int2* smem /* 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int2 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc.x, o), b = __shfl_up_sync(0xffffffffu, inc.y, o);
        if (lane >= o) { inc.x += a; inc.y += b; }
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int2 w = lane < (int)(blockDim.x >> 5) ? smem[lane] : make_int2(0, 0);
        int2 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, wi.x, o), b = __shfl_up_sync(0xffffffffu, wi.y, o);
            if (lane >= o) { wi.x += a; wi.y += b; }
        }
        smem[lane] = make_int2(wi.x - w.x, wi.y - w.y);
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    const int2 off = smem[warp];
    __syncthreads();
    return make_int2(off.x + inc.x - v.x, off.y + inc.y - v.y);
}

// pass 1: per-chunk totals of (count, occupied flag)
__global__ void k_scan_reduce(const int* __restrict__ blk_count, int n, int n_real, int2* __restrict__ partial) {
    __shared__ int2 sm[32];
    __shared__ int2 tot;
    int2 v = make_int2(0, 0);
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_E;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        const int i = base + e;
        if (i < n) { const int c = blk_count[i]; v.x += c; v.y += (c > 0 && i < n_real) ? 1 : 0; }
    }
    block_scan_excl(v, &tot, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// pass 2: one CTA scans the chunk totals in place (exclusive)
__global__ void k_scan_partials(int2* partial, int n_chunks, DevCounters* dc, const int* __restrict__ blk_count, int n_real) {
    __shared__ int2 sm[32];
    __shared__ int2 tot;
    int2 carry = make_int2(0, 0);
    for (int base = 0; base < n_chunks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int2 v = i < n_chunks ? partial[i] : make_int2(0, 0);
        const int2 ex = block_scan_excl(v, &tot, sm);
        if (i < n_chunks) partial[i] = make_int2(ex.x + carry.x, ex.y + carry.y);
        carry.x += tot.x; carry.y += tot.y;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        dc->n_active_pblocks = carry.y;
        dc->n_active_gblocks = 0; dc->n_active_nodes = 0;
        dc->work_a = 0; dc->work_b = 0; dc->work_c = 0;
        dc->epoch += 1;
        // bucket counts of the three special buckets (parked, leaving down, leaving up)
        dc->n_out_of_grid = blk_count[n_real];
        dc->n_out_down = blk_count[n_real + 1];
        dc->n_out_up = blk_count[n_real + 2];
    }
}
// pass 3: write blk_start / blk_cursor, the occupied particle-block list, and mark the 2x2x2 grid blocks of
// every occupied particle block (appended once each to the active grid-block list through an epoch stamp)
__global__ void k_scan_apply(const int* __restrict__ blk_count, int n, int n_real, const int2* __restrict__ partial,
                             int* __restrict__ blk_start, int* __restrict__ blk_cursor, int4* __restrict__ pblock_list,
                             int* __restrict__ gflag, int* __restrict__ gblock_list, DevCounters* dc, GridDims gd) {
    __shared__ int2 sm[32];
    __shared__ int2 tot;
    int c[SCAN_E];
    int2 v = make_int2(0, 0);
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_E;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        const int i = base + e;
        c[e] = i < n ? blk_count[i] : 0;
        v.x += c[e]; v.y += (c[e] > 0 && i < n_real) ? 1 : 0;
    }
    int2 ex = block_scan_excl(v, &tot, sm);
    const int2 off = partial[blockIdx.x];
    ex.x += off.x; ex.y += off.y;
    const int epoch = dc->epoch;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        const int i = base + e;
        if (i < n) {
            blk_start[i] = ex.x; blk_cursor[i] = ex.x;
            if (i == n_real) dc->n_binned = ex.x;                 // start of the parked bucket
            // what a re-sort keeps besides the binned particles: parked ones AND the two "leaving the slab" buckets. The latter
            // are empty unless a migration buffer overflowed (k_mark_outgoing*: such a leaver stays alive here one more
            // substep, frozen like a parked particle, and is offered to the neighbour again after the next gather)
            if (i == n_real + 2) dc->n_sorted = ex.x + c[e];
            if (c[e] > 0 && i < n_real) {
                const int pbk = i % gd.npbk, pbj = (i / gd.npbk) % gd.npbj, pbi = i / (gd.npbk * gd.npbj);
                // work item: (block id, first sorted rank, count, slab-local block coordinates) -- the tile kernels read the
                // coordinates instead of dividing the block id again in every thread of every block
                pblock_list[ex.y] = make_int4(i, ex.x, c[e], pack_block_coords(pbi, pbj, pbk));
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const int gb = ((pbi + (d >> 2)) * gd.nbj + pbj + ((d >> 1) & 1)) * gd.nbk + pbk + (d & 1);
                    if (atomicExch(&gflag[gb], epoch) != epoch) gblock_list[atomicAdd(&dc->n_active_gblocks, 1)] = gb;
                }
            }
            ex.x += c[e]; ex.y += (c[e] > 0 && i < n_real) ? 1 : 0;
        }
    }
}
// slab exchange layers are always active (they are packed / added densely)
__global__ void k_mark_layer(int layer, GridDims gd, int* __restrict__ gflag, int* __restrict__ gblock_list, DevCounters* dc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= gd.nbj * gd.nbk) return;
    const int gb = layer * gd.nbj * gd.nbk + t;
    const int epoch = dc->epoch;
    if (atomicExch(&gflag[gb], epoch) != epoch) gblock_list[atomicAdd(&dc->n_active_gblocks, 1)] = gb;
}

// ------------------------------------------------------------------------------------------------------
// grid kernels (over the active grid-block list; 64 threads = one block of nodes)
// ------------------------------------------------------------------------------------------------------
__global__ void k_grid_clear(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc,
                             float4* __restrict__ grid, float4* __restrict__ gforce) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    for (int b = blockIdx.x * per + sub; b < nb; b += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[b] * 64 + t;
        grid[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gforce) gforce[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

enum { GU_NORMALIZE = 1, GU_FORCE = 2, GU_GRAVITY = 4, GU_COLLIDE = 8, GU_COUNT = 16 };
// FLAGS select the reference stages folded into this pass:
//  NORMALIZE  velocity = momentum / mass              (cpp:118-121)
//  FORCE      velocity += dt * force / mass           (cpp:260, staged path with a separate force array)
//  GRAVITY    velocity += dt * g                      (cpp:260)   [fused path: momentum already holds dt*f]
//  COLLIDE    bodyCollision                           (cpp:298-304)
template <int FLAGS>
__global__ void k_grid_update(const int* __restrict__ gblock_list, DevCounters* dc, float4* __restrict__ grid,
                              const float4* __restrict__ gforce, GridDims gd, SimConst sc, float dt,
                              const __grid_constant__ ColliderSet cs, int n_colliders) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    int used = 0;
    for (int b = blockIdx.x * per + sub; b < nb; b += gridDim.x * per) {
        const int gb = gblock_list[b];
        const size_t idx = (size_t)gb * 64 + t;
        float4 n = grid[idx];
        if (n.x == 0.0f) continue;                       // used_cells: mass != 0 (cpp:105-110)
        ++used;
        float v[3] = { n.y, n.z, n.w };
        if (FLAGS & GU_NORMALIZE) { v[0] = div_rn(v[0], n.x); v[1] = div_rn(v[1], n.x); v[2] = div_rn(v[2], n.x); }
        if ((FLAGS & GU_FORCE) && (FLAGS & GU_GRAVITY)) {
            const float4 f = gforce[idx];                // v += dt * (f/m + g), reference association
            v[0] = add_rn(v[0], mul_rn(dt, add_rn(div_rn(f.y, n.x), sc.g[0])));   // force array = (0, fx, fy, fz)
            v[1] = add_rn(v[1], mul_rn(dt, add_rn(div_rn(f.z, n.x), sc.g[1])));
            v[2] = add_rn(v[2], mul_rn(dt, add_rn(div_rn(f.w, n.x), sc.g[2])));
        } else if (FLAGS & GU_GRAVITY) {
            v[0] = add_rn(v[0], mul_rn(dt, sc.g[0])); v[1] = add_rn(v[1], mul_rn(dt, sc.g[1])); v[2] = add_rn(v[2], mul_rn(dt, sc.g[2]));
        }
        if (FLAGS & GU_COLLIDE) {
            const int gbk = gb % gd.nbk, gbj = (gb / gd.nbk) % gd.nbj, gbi = gb / (gd.nbk * gd.nbj) + gd.lo;
            const int i = gbi * 4 + (t >> 4), j = gbj * 4 + ((t >> 2) & 3), k = gbk * 4 + (t & 3);
            body_collision_rn(cs.c, n_colliders, sc.friction, mul_rn((float)i, sc.h), mul_rn((float)j, sc.h),
                              mul_rn((float)k, sc.h), v);
        }
        grid[idx] = make_float4(n.x, v[0], v[1], v[2]);
    }
    if (FLAGS & GU_COUNT) {
        used = __reduce_add_sync(0xffffffffu, used);
        if ((threadIdx.x & 31) == 0 && used) atomicAdd(&dc->n_active_nodes, used);
    }
}

// ------------------------------------------------------------------------------------------------------
// particle helpers
// ------------------------------------------------------------------------------------------------------
MPM_DI size_t node_index(const GridDims& gd, int i, int j, int k) {   // global node -> blocked storage index
    const int gb = (((i >> 2) - gd.lo) * gd.nbj + (j >> 2)) * gd.nbk + (k >> 2);
    return (size_t)gb * 64 + (((i & 3) * 4 + (j & 3)) * 4 + (k & 3));
}

enum { P2G_MOMENTUM = 0, P2G_FORCE = 1, P2G_FUSED = 2 };

// affine scatter coefficients of one particle: contribution to node x_i is
//   w * (mass_ch, a0 + A * (x_i - x_p))      with A row-major here: A[r*3+c]
struct P2GInG { float4 xm, b0, b1, b2, t0, t1; };      // the six planes P2G reads of one particle
template <int MODE>
MPM_DI P2GInG p2g_load_planes(const Planes& P, int gid) {
    P2GInG r;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    r.xm = P.p[0][gid];
    r.b0 = r.b1 = r.b2 = r.t0 = r.t1 = z;
    if (MODE != P2G_FORCE) { r.b0 = P.p[1][gid]; r.b1 = P.p[2][gid]; r.b2 = P.p[3][gid]; }
    if (MODE != P2G_MOMENTUM) { r.t0 = P.p[4][gid]; r.t1 = P.p[5][gid]; }
    return r;
}
template <int MODE, class In>
MPM_DI void p2g_coeffs(const In& in, float dinv, float dt, float& mass_ch, float (&a0)[3], float (&A)[9]) {
    const float m = in.xm.w;
    float Bm[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 }, v[3] = { 0, 0, 0 }, tau[6] = { 0, 0, 0, 0, 0, 0 };
    if (MODE != P2G_FORCE) {
        const float4 b0 = in.b0, b1 = in.b1, b2 = in.b2;
        Bm[0] = b0.x; Bm[1] = b0.y; Bm[2] = b0.z; Bm[3] = b0.w; Bm[4] = b1.x; Bm[5] = b1.y; Bm[6] = b1.z; Bm[7] = b1.w; Bm[8] = b2.x;
        v[0] = b2.y; v[1] = b2.z; v[2] = b2.w;
    }
    if (MODE != P2G_MOMENTUM) {
        const float4 t0 = in.t0, t1 = in.t1;
        tau[0] = t0.x; tau[1] = t0.y; tau[2] = t0.z; tau[3] = t0.w; tau[4] = t1.x; tau[5] = t1.y;
    }
    const float md = m * dinv;
    const float s = (MODE == P2G_FORCE) ? 1.0f : dt;        // force mode scatters -M, fused mode -dt*M
    // glm B[c*3+r] -> row-major A[r*3+c]; M symmetric from tau
    const float M[9] = { tau[0], tau[3], tau[4], tau[3], tau[1], tau[5], tau[4], tau[5], tau[2] };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = md * Bm[c * 3 + r] - s * M[r * 3 + c];
    a0[0] = m * v[0]; a0[1] = m * v[1]; a0[2] = m * v[2];
    mass_ch = (MODE == P2G_FORCE) ? 0.0f : m;
}

// ------------------------------------------------------------------------------------------------------
// P2G, variant 1 (baseline / debug): one thread per particle, one vector red.global per (particle, node)
// ------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void k_p2g_atomic(Planes P, const int* __restrict__ sorted_ids, const DevCounters* __restrict__ dc,
                             float4* __restrict__ grid, GridDims gd, SimConst sc, float dt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dc->n_binned) return;
    const int p = sorted_ids[j];
    float mch, a0[3], A[9];
    const P2GInG in = p2g_load_planes<MODE>(P, p);
    const float4 xm = in.xm;
    p2g_coeffs<MODE>(in, sc.dinv, dt, mch, a0, A);
    const int cx = cell_of(xm.x, sc.pd), cy = cell_of(xm.y, sc.pd), cz = cell_of(xm.z, sc.pd);
    float wx[4], wy[4], wz[4];
    axis_weights(xm.x, sc.pd, cx, wx); axis_weights(xm.y, sc.pd, cy, wy); axis_weights(xm.z, sc.pd, cz, wz);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float dx = (float)(cx - 1 + a) * sc.h - xm.x;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float dy = (float)(cy - 1 + b) * sc.h - xm.y;
            const float wxy = wx[a] * wy[b];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float dz = (float)(cz - 1 + c) * sc.h - xm.z;
                const float w = wxy * wz[c];
                if (w == 0.0f) continue;
                float4 val;
                val.x = w * mch;
                val.y = w * (a0[0] + A[0] * dx + A[1] * dy + A[2] * dz);
                val.z = w * (a0[1] + A[3] * dx + A[4] * dy + A[5] * dz);
                val.w = w * (a0[2] + A[6] * dx + A[7] * dy + A[8] * dz);
                atomicAdd(&grid[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)], val);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// per-particle kernels shared by the staged API and the fused path
// ------------------------------------------------------------------------------------------------------
struct ParticleRegs {
    float x[3], m, v[3], V0, B[9], FE[9], FP[9], tau[6];
    int pid;
};
MPM_DI void load_particle(const Planes& P, int p, ParticleRegs& r, bool need_v_tau) {
    const float4 a0 = P.p[0][p], a1 = P.p[1][p], a2 = P.p[2][p], a3 = P.p[3][p];
    const float4 a6 = P.p[6][p], a7 = P.p[7][p], a8 = P.p[8][p], a9 = P.p[9][p], a10 = P.p[10][p];
    r.x[0] = a0.x; r.x[1] = a0.y; r.x[2] = a0.z; r.m = a0.w;
    r.B[0] = a1.x; r.B[1] = a1.y; r.B[2] = a1.z; r.B[3] = a1.w; r.B[4] = a2.x; r.B[5] = a2.y; r.B[6] = a2.z; r.B[7] = a2.w; r.B[8] = a3.x;
    r.v[0] = a3.y; r.v[1] = a3.z; r.v[2] = a3.w;
    r.V0 = a6.x; r.pid = __float_as_int(a6.y);
    r.FE[0] = a6.z; r.FE[1] = a6.w; r.FE[2] = a7.x; r.FE[3] = a7.y; r.FE[4] = a7.z; r.FE[5] = a7.w; r.FE[6] = a8.x; r.FE[7] = a8.y; r.FE[8] = a8.z;
    r.FP[0] = a8.w; r.FP[1] = a9.x; r.FP[2] = a9.y; r.FP[3] = a9.z; r.FP[4] = a9.w; r.FP[5] = a10.x; r.FP[6] = a10.y; r.FP[7] = a10.z; r.FP[8] = a10.w;
    if (need_v_tau) {
        const float4 a4 = P.p[4][p], a5 = P.p[5][p];
        r.tau[0] = a4.x; r.tau[1] = a4.y; r.tau[2] = a4.z; r.tau[3] = a4.w; r.tau[4] = a5.x; r.tau[5] = a5.y;
    }
}
MPM_DI void store_particle(const Planes& P, int p, const ParticleRegs& r) {
    P.p[0][p] = make_float4(r.x[0], r.x[1], r.x[2], r.m);
    P.p[1][p] = make_float4(r.B[0], r.B[1], r.B[2], r.B[3]);
    P.p[2][p] = make_float4(r.B[4], r.B[5], r.B[6], r.B[7]);
    P.p[3][p] = make_float4(r.B[8], r.v[0], r.v[1], r.v[2]);
    P.p[4][p] = make_float4(r.tau[0], r.tau[1], r.tau[2], r.tau[3]);
    P.p[5][p] = make_float4(r.tau[4], r.tau[5], 0.0f, 0.0f);
    P.p[6][p] = make_float4(r.V0, __int_as_float(r.pid), r.FE[0], r.FE[1]);
    P.p[7][p] = make_float4(r.FE[2], r.FE[3], r.FE[4], r.FE[5]);
    P.p[8][p] = make_float4(r.FE[6], r.FE[7], r.FE[8], r.FP[0]);
    P.p[9][p] = make_float4(r.FP[1], r.FP[2], r.FP[3], r.FP[4]);
    P.p[10][p] = make_float4(r.FP[5], r.FP[6], r.FP[7], r.FP[8]);
}

// F-update + stress of one particle held in registers (cpp:306-330 then the per-particle part of cpp:235-248)
MPM_DI bool particle_f_update(ParticleRegs& r, const SimConst& sc, float dt) {
    float Ug[9], Sg[3];
    if (!f_update_rn(r.B, r.FE, r.FP, sc.dinv, dt, sc.clamp_lo, sc.clamp_hi, Ug, Sg)) return false;
    const float J = m3_det_rn(r.FE), dFP = m3_det_rn(r.FP);
    tau_from_factors(Ug, Sg, J, dFP, r.V0, sc.dinv, sc.E, sc.nu, sc.xi, r.tau);
    return true;
}

// recompute tau for arbitrary (uploaded) FE / FP / V0
__global__ void k_stress(Planes P, const DevCounters* __restrict__ dc, SimConst sc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    ParticleRegs r;
    load_particle(P, p, r, false);
    if (r.m < 0.0f) return;
    tau_general(r.FE, m3_det_rn(r.FP), r.V0, sc.dinv, sc.E, sc.nu, sc.xi, r.tau);
    P.p[4][p] = make_float4(r.tau[0], r.tau[1], r.tau[2], r.tau[3]);
    P.p[5][p] = make_float4(r.tau[4], r.tau[5], 0.0f, 0.0f);
}

// updateDeformationGradient (cpp:306-330) + next substep's stress, one thread per sorted slot. It touches only
// B (read), FE/FP/V0 (read) and FE/FP/tau (write), i.e. planes disjoint from what the gather kernel writes, so in the
// re-sorting fused path the two kernels split the particle record between them at no extra HBM traffic.
// The F-update + next substep's stress of one particle from its loaded planes (1..3: B, 6..10: V0, pid, FE, FP); results to
// planes 4..10 of D at slot q. FAST = tolerance form (f_update_fast, the fused substep's default), else bit-faithful.
struct FUpdIn { float4 a1, a2, a3, a6, a7, a8, a9, a10; };
MPM_DI FUpdIn fupd_load(const Planes& cur, int p) {
    FUpdIn r;
    r.a1 = cur.p[1][p]; r.a2 = cur.p[2][p]; r.a3 = cur.p[3][p];
    r.a6 = cur.p[6][p]; r.a7 = cur.p[7][p]; r.a8 = cur.p[8][p]; r.a9 = cur.p[9][p]; r.a10 = cur.p[10][p];
    return r;
}
template <bool FAST>
MPM_DI void fupd_compute_store(const FUpdIn& in, const Planes& D, int q, DevCounters* dc, const SimConst& sc, float dt) {
    float B[9] = { in.a1.x, in.a1.y, in.a1.z, in.a1.w, in.a2.x, in.a2.y, in.a2.z, in.a2.w, in.a3.x };
    float FE[9] = { in.a6.z, in.a6.w, in.a7.x, in.a7.y, in.a7.z, in.a7.w, in.a8.x, in.a8.y, in.a8.z };
    float FP[9] = { in.a8.w, in.a9.x, in.a9.y, in.a9.z, in.a9.w, in.a10.x, in.a10.y, in.a10.z, in.a10.w };
    float Ug[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, Sg[3] = { 1, 1, 1 }, tau[6];
    if (FAST) {
        if (!f_update_fast(B, FE, FP, sc.dinv * dt, sc.clamp_lo, sc.clamp_hi, Ug, Sg)) { dc->svd_failed = 1; }
        tau_from_factors_fast(Ug, Sg, m3_det_fast(FE), m3_det_fast(FP), in.a6.x, sc.dinv, sc.mu0, sc.lambda0, sc.xi, tau);
    } else {
        if (!f_update_rn(B, FE, FP, sc.dinv, dt, sc.clamp_lo, sc.clamp_hi, Ug, Sg)) { dc->svd_failed = 1; }
        tau_from_factors(Ug, Sg, m3_det_rn(FE), m3_det_rn(FP), in.a6.x, sc.dinv, sc.E, sc.nu, sc.xi, tau);
    }
    D.p[4][q] = make_float4(tau[0], tau[1], tau[2], tau[3]);
    D.p[5][q] = make_float4(tau[4], tau[5], 0.0f, 0.0f);
    D.p[6][q] = make_float4(in.a6.x, in.a6.y, FE[0], FE[1]);
    D.p[7][q] = make_float4(FE[2], FE[3], FE[4], FE[5]);
    D.p[8][q] = make_float4(FE[6], FE[7], FE[8], FP[0]);
    D.p[9][q] = make_float4(FP[1], FP[2], FP[3], FP[4]);
    D.p[10][q] = make_float4(FP[5], FP[6], FP[7], FP[8]);
}
template <bool REORDER, bool FAST = false>
__global__ void __launch_bounds__(256)
k_fupdate(Planes cur, Planes nxt, const int* __restrict__ sorted_ids, DevCounters* dc, SimConst sc, float dt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dc->n_binned) return;
    const int p = sorted_ids[j];
    const FUpdIn in = fupd_load(cur, p);
    fupd_compute_store<FAST>(in, REORDER ? nxt : cur, REORDER ? j : p, dc, sc, dt);
}

// computeParticleVolumesAndDensities (cpp:131-142): density = sum m_i w_ip / h^3, V0 = m / density
__global__ void k_volumes(Planes P, const int* __restrict__ sorted_ids, const DevCounters* __restrict__ dc,
                          const float4* __restrict__ grid, GridDims gd, SimConst sc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dc->n_sorted) return;
    const int p = sorted_ids[j];
    const float4 xm = P.p[0][p];
    float density = 0.0f;
    if (j < dc->n_binned) {
        const int cx = cell_of(xm.x, sc.pd), cy = cell_of(xm.y, sc.pd), cz = cell_of(xm.z, sc.pd);
        float wx[4], wy[4], wz[4];
        axis_weights(xm.x, sc.pd, cx, wx); axis_weights(xm.y, sc.pd, cy, wy); axis_weights(xm.z, sc.pd, cz, wz);
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) for (int c = 0; c < 4; ++c) {
            const float w = mul_rn(mul_rn(wx[a], wy[b]), wz[c]);
            density = add_rn(density, mul_rn(grid[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)].x, w));
        }
    }
    density = div_rn(density, mul_rn(mul_rn(sc.h, sc.h), sc.h));
    float4 a6 = P.p[6][p];
    a6.x = density != 0.0f ? div_rn(xm.w, density) : 0.0f;
    P.p[6][p] = a6;
}

enum { G2P_F = 1, G2P_GATHER = 2, G2P_ADVECT = 4, G2P_REORDER = 8, G2P_HIST = 16, G2P_GRADW = 32 };

MPM_DI void advect_rn(ParticleRegs& r, const SimConst& sc, float dt) {   // cpp:344-350, 381-388
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float x = add_rn(r.x[a], mul_rn(r.v[a], dt));
        if (x < sc.pos_lo) x = sc.pos_lo;
        if (sc.pos_hi[a] < x) x = sc.pos_hi[a];
        r.x[a] = x;
    }
}

// ------------------------------------------------------------------------------------------------------
// P2G, variant 9 (deterministic debug mode, SURVEY section 7 hard part 3): ONE thread adds the particles' contributions in
// ascending particle-id order with plain (non-atomic) additions, so the grid -- and with it the whole substep, whose other
// stages are pure per-particle / per-node functions -- is bitwise reproducible from run to run. Same per-node arithmetic
// as k_p2g_atomic. Minutes per substep at millions of particles: for bisecting differences on small scenes only.
// ------------------------------------------------------------------------------------------------------
__global__ void k_slot_of_pid(Planes P, const int* __restrict__ key, const DevCounters* __restrict__ dc, GridDims gd,
                              int* __restrict__ slot_of_pid, int n_pid) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= dc->n_slots) return;
    const int k = key[slot];
    const int pid = __float_as_int(P.p[6][slot].y);
    if (k >= 0 && k < gd.n_pblocks && pid >= 0 && pid < n_pid) slot_of_pid[pid] = slot;      // binned particles only
}
template <int MODE>
__global__ void k_p2g_serial(Planes P, const int* __restrict__ slot_of_pid, int n_pid, float4* __restrict__ grid, GridDims gd,
                             SimConst sc, float dt) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int pid = 0; pid < n_pid; ++pid) {
        const int p = slot_of_pid[pid];
        if (p < 0) continue;
        float mch, a0[3], A[9];
        const P2GInG in = p2g_load_planes<MODE>(P, p);
        const float4 xm = in.xm;
        p2g_coeffs<MODE>(in, sc.dinv, dt, mch, a0, A);
        const int cx = cell_of(xm.x, sc.pd), cy = cell_of(xm.y, sc.pd), cz = cell_of(xm.z, sc.pd);
        float wx[4], wy[4], wz[4];
        axis_weights(xm.x, sc.pd, cx, wx); axis_weights(xm.y, sc.pd, cy, wy); axis_weights(xm.z, sc.pd, cz, wz);
        for (int a = 0; a < 4; ++a) {
            const float dx = (float)(cx - 1 + a) * sc.h - xm.x;
            for (int b = 0; b < 4; ++b) {
                const float dy = (float)(cy - 1 + b) * sc.h - xm.y;
                const float wxy = wx[a] * wy[b];
                for (int c = 0; c < 4; ++c) {
                    const float dz = (float)(cz - 1 + c) * sc.h - xm.z;
                    const float w = wxy * wz[c];
                    if (w == 0.0f) continue;
                    float4& g = grid[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)];
                    g.x += w * mch;
                    g.y += w * (a0[0] + A[0] * dx + A[1] * dy + A[2] * dz);
                    g.z += w * (a0[1] + A[3] * dx + A[4] * dy + A[5] * dz);
                    g.w += w * (a0[2] + A[6] * dx + A[7] * dy + A[8] * dz);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// G2P, variant 1 (baseline / debug / staged API): one thread per sorted slot, nodes read straight from global
// ------------------------------------------------------------------------------------------------------
template <int FLAGS>
__global__ void k_g2p_direct(Planes cur, Planes nxt, const int* __restrict__ sorted_ids, DevCounters* dc,
                             const float4* __restrict__ grid, GridDims gd, SimConst sc, float dt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dc->n_sorted) return;
    const int p = sorted_ids[j];
    ParticleRegs r;
    load_particle(cur, p, r, true);
    if (j < dc->n_binned) {
        if (FLAGS & G2P_F) {
            if (!particle_f_update(r, sc, dt)) dc->svd_failed = 1;
        }
        if (FLAGS & G2P_GATHER) {
            const int cx = cell_of(r.x[0], sc.pd), cy = cell_of(r.x[1], sc.pd), cz = cell_of(r.x[2], sc.pd);
            float wx[4], wy[4], wz[4];
            axis_weights(r.x[0], sc.pd, cx, wx); axis_weights(r.x[1], sc.pd, cy, wy); axis_weights(r.x[2], sc.pd, cz, wz);
            float v[3] = { 0, 0, 0 }, Bn[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float dx = (float)(cx - 1 + a) * sc.h - r.x[0];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float dy = (float)(cy - 1 + b) * sc.h - r.x[1];
                    const float wxy = wx[a] * wy[b];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float dz = (float)(cz - 1 + c) * sc.h - r.x[2];
                        const float w = wxy * wz[c];
                        const float4 n = __ldg(&grid[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)]);
                        const float wv[3] = { w * n.y, w * n.z, w * n.w };
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            v[q] += wv[q];
                            Bn[0 + q] += wv[q] * dx; Bn[3 + q] += wv[q] * dy; Bn[6 + q] += wv[q] * dz;   // glm B[c][r] += w v[r] dx[c]
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) r.v[q] = v[q];
#pragma unroll
            for (int q = 0; q < 9; ++q) r.B[q] = Bn[q];
        }
        if (FLAGS & G2P_ADVECT) advect_rn(r, sc, dt);
    }
    store_particle((FLAGS & G2P_REORDER) ? nxt : cur, (FLAGS & G2P_REORDER) ? j : p, r);
}

// ------------------------------------------------------------------------------------------------------
// host <-> device format conversion
// ------------------------------------------------------------------------------------------------------
// scatter live particles back to upload order (by pid) into 11 staging planes
__global__ void k_unsort(Planes cur, Planes out, const DevCounters* __restrict__ dc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    const float4 a0 = cur.p[0][p];
    if (a0.w < 0.0f) return;
    const int pid = __float_as_int(cur.p[6][p].y);
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) out.p[k][pid] = cur.p[k][p];
}
__global__ void k_render(Planes cur, float4* __restrict__ xyzs, const DevCounters* __restrict__ dc, float size) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    const float4 a0 = cur.p[0][p];
    if (a0.w < 0.0f) return;
    const int pid = __float_as_int(cur.p[6][p].y);
    xyzs[pid] = make_float4(a0.x, a0.y, a0.z, size);
}
// slab mode: particles migrate between handles, so ids are not dense; render buffers come out in storage order and
// retired slots get size 0 (an invisible billboard)
__global__ void k_render_slots(Planes cur, float4* __restrict__ xyzs, const DevCounters* __restrict__ dc, float size, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 a0 = p < dc->n_slots ? cur.p[0][p] : make_float4(0.f, 0.f, 0.f, -1.f);
    xyzs[p] = make_float4(a0.x, a0.y, a0.z, a0.w < 0.0f ? 0.0f : size);
}
__global__ void k_binning_debug(Planes cur, const int* __restrict__ key, const DevCounters* __restrict__ dc, PosDiv h,
                                int* __restrict__ cells3, int* __restrict__ key_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    const float4 a0 = cur.p[0][p];
    if (a0.w < 0.0f) return;
    const int pid = __float_as_int(cur.p[6][p].y);
    cells3[pid * 3 + 0] = cell_of(a0.x, h); cells3[pid * 3 + 1] = cell_of(a0.y, h); cells3[pid * 3 + 2] = cell_of(a0.z, h);
    key_out[pid] = key[p];
}
// ---- slab migration: pack particles whose block layer left [lo, hi) and retire their slots ----
// record = 11 float4 (the particle's planes, particle-major) = MPM_MIGRATE_FLOATS floats
__global__ void k_mark_outgoing(Planes cur, DevCounters* dc, GridDims gd, PosDiv pd, float4* __restrict__ out_down,
                                float4* __restrict__ out_up, int cap) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    const float4 a0 = cur.p[0][p];
    if (a0.w < 0.0f) return;
    int cells[3];
    const int k = particle_key(a0, gd, pd, cells);
    if (k != gd.n_pblocks + 1 && k != gd.n_pblocks + 2) return;
    const int dir = k - (gd.n_pblocks + 1);
    const int idx = atomicAdd(&dc->n_mig[dir], 1);
    if (idx >= cap) { dc->mig_overflow = 1; atomicSub(&dc->n_mig[dir], 1); return; }   // stays here one more substep
    float4* o = (dir ? out_up : out_down) + (size_t)idx * NPLANES;
#pragma unroll
    for (int q = 0; q < NPLANES; ++q) o[q] = cur.p[q][p];
    cur.p[0][p] = make_float4(a0.x, a0.y, a0.z, -1.0f);     // dead slot: skipped by the binning, dropped by the next re-sort
}
__global__ void k_append_incoming(Planes cur, DevCounters* dc, const float4* __restrict__ in, int base, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) dc->n_slots = base + n;
    if (i >= n) return;
    const float4* r = in + (size_t)i * NPLANES;
#pragma unroll
    for (int q = 0; q < NPLANES; ++q) cur.p[q][base + i] = r[q];
}
// Sync-free variant: the packed buffers start with one float4 header whose first int is the record count (it IS the
// atomic cursor), so the receiving side learns the count on the device and the host never has to.
__global__ void k_mark_outgoing_hdr(Planes cur, DevCounters* dc, GridDims gd, PosDiv pd, float4* __restrict__ out_down,
                                    float4* __restrict__ out_up, int cap) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    const float4 a0 = cur.p[0][p];
    if (a0.w < 0.0f) return;
    int cells[3];
    const int k = particle_key(a0, gd, pd, cells);
    if (k != gd.n_pblocks + 1 && k != gd.n_pblocks + 2) return;
    float4* buf = (k == gd.n_pblocks + 2) ? out_up : out_down;
    const int idx = atomicAdd(reinterpret_cast<int*>(buf), 1);
    if (idx >= cap) { dc->mig_overflow = 1; return; }      // header count may exceed cap; the receiver clamps it
    float4* o = buf + 1 + (size_t)idx * NPLANES;
#pragma unroll
    for (int q = 0; q < NPLANES; ++q) o[q] = cur.p[q][p];
    cur.p[0][p] = make_float4(a0.x, a0.y, a0.z, -1.0f);
}
__global__ void k_append_incoming_hdr(Planes cur, const DevCounters* __restrict__ dc, const float4* __restrict__ in, int cap, int capacity,
                                      GridDims gd, PosDiv pd, int* __restrict__ key_out, int* __restrict__ blk_count) {
    const int n = min(*reinterpret_cast<const int*>(in), cap);
    const int base = dc->n_slots;          // advanced by k_bump_slots after this kernel
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || base + i >= capacity) return;
    const float4* r = in + 1 + (size_t)i * NPLANES;
#pragma unroll
    for (int q = 0; q < NPLANES; ++q) cur.p[q][base + i] = r[q];
    if (key_out) {                          // keep the gather's keys + histogram of the current buffer complete (fused binning)
        int cells[3];
        const int k = particle_key(r[0], gd, pd, cells);
        key_out[base + i] = k;
        if (k >= 0) atomicAdd(&blk_count[k], 1);
    }
}
__global__ void k_bump_slots(DevCounters* dc, const float4* __restrict__ in, int cap, int capacity) {
    const int n = min(*reinterpret_cast<const int*>(in), cap);
    if (dc->n_slots + n > capacity) { dc->mig_overflow = 2; dc->n_slots = capacity; }
    else dc->n_slots += n;
}
// live particles in slot order as 35-float rows (+ pid) for distributed downloads
__global__ void k_export_live(Planes cur, const DevCounters* __restrict__ dc, float* __restrict__ out35, int* __restrict__ pid_out,
                              int* __restrict__ counter, int cap) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dc->n_slots) return;
    ParticleRegs r;
    load_particle(cur, p, r, false);
    if (r.m < 0.0f) return;
    const int i = atomicAdd(counter, 1);
    if (i >= cap) return;
    float* o = out35 + (size_t)i * 35;
    o[0] = r.m; o[1] = r.v[0]; o[2] = r.v[1]; o[3] = r.v[2]; o[4] = r.V0; o[5] = r.x[0]; o[6] = r.x[1]; o[7] = r.x[2];
#pragma unroll
    for (int q = 0; q < 9; ++q) { o[8 + q] = r.FE[q]; o[17 + q] = r.FP[q]; o[26 + q] = r.B[q]; }
    pid_out[i] = r.pid;
}

// conserved quantities of the live particles, for run-time correctness checks (bench.py prints them at every GPU count):
// fsum = { sum m, sum m vx, sum m vy, sum m vz, sum m y } in fp64, isum = { count, sum pid, sum hash(pid) } mod 2^64
MPM_DI unsigned long long pid_hash(unsigned long long pid) {          // splitmix64 finaliser
    unsigned long long z = pid + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
k_invariants(Planes cur, const DevCounters* __restrict__ dc, double* __restrict__ fsum, unsigned long long* __restrict__ isum) {
    double f[5] = { 0, 0, 0, 0, 0 };
    unsigned long long c[3] = { 0, 0, 0 };
    const int n = dc->n_slots;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float4 a0 = cur.p[0][p];
        if (a0.w < 0.0f) continue;
        const float4 a3 = cur.p[3][p];
        const unsigned long long pid = (unsigned long long)(unsigned)__float_as_int(cur.p[6][p].y);
        const double m = (double)a0.w;
        f[0] += m; f[1] += m * (double)a3.y; f[2] += m * (double)a3.z; f[3] += m * (double)a3.w; f[4] += m * (double)a0.y;
        c[0] += 1ull; c[1] += pid; c[2] += pid_hash(pid);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) f[k] += __shfl_down_sync(0xffffffffu, f[k], o);
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] += __shfl_down_sync(0xffffffffu, c[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) atomicAdd(&fsum[k], f[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(&isum[k], c[k]);
    }
}

// blocked <-> linear grid (download_grid / upload_grid, tests only)
__global__ void k_grid_export(const float4* __restrict__ grid, const float4* __restrict__ gforce, GridDims gd, float* __restrict__ out7) {
    const size_t n = (size_t)gd.I * gd.J * gd.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % gd.K), j = (int)((t / gd.K) % gd.J), i = (int)(t / ((size_t)gd.K * gd.J));
    float* o = out7 + t * 7;
    const int bi = i >> 2;
    if (bi < gd.lo || bi > gd.hi) { for (int q = 0; q < 7; ++q) o[q] = 0.0f; return; }
    const size_t idx = node_index(gd, i, j, k);
    const float4 n4 = grid[idx];
    const float4 f = gforce ? gforce[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    o[0] = n4.x; o[1] = f.y; o[2] = f.z; o[3] = f.w; o[4] = n4.y; o[5] = n4.z; o[6] = n4.w;
}
__global__ void k_grid_import(float4* __restrict__ grid, float4* __restrict__ gforce, GridDims gd, const float* __restrict__ in7) {
    const size_t n = (size_t)gd.I * gd.J * gd.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % gd.K), j = (int)((t / gd.K) % gd.J), i = (int)(t / ((size_t)gd.K * gd.J));
    const int bi = i >> 2;
    if (bi < gd.lo || bi > gd.hi) return;
    const float* o = in7 + t * 7;
    const size_t idx = node_index(gd, i, j, k);
    grid[idx] = make_float4(o[0], o[4], o[5], o[6]);
    if (gforce) gforce[idx] = make_float4(0.0f, o[1], o[2], o[3]);
}
// make every grid block that holds a non-zero node active (after upload_grid)
__global__ void k_activate_all(int n_gblocks, int* __restrict__ gflag, int* __restrict__ gblock_list, DevCounters* dc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_gblocks) return;
    gflag[b] = dc->epoch;
    gblock_list[b] = b;
    if (b == 0) dc->n_active_gblocks = n_gblocks;
}

}  // namespace mpm

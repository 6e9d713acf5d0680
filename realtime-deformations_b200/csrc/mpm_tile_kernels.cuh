// Block-tile P2G / G2P kernels: one CTA (P2G) or one warp (G2P) works on one occupied 4^3-cell particle block at a time
// (persistent CTAs pulling blocks from a device-side work counter, so no host sync is needed to size the grid).
//
// P2G (k_p2g_tile):  no shared-memory float atomics at all (on sm_100a they are CAS loops, ATOMS.CAST.SPIN).
//   derive       : 512 particles at a time (2 per thread) are loaded as coalesced float4 planes; each thread derives its
//                  particles' 12 axis weights (branch-free fp32 form, <= 1 ulp from the reference's) and affine coefficients
//                  into shared memory and draws its rank inside its cell with one shared integer atomic.
//   sort         : every warp scans the 64 cell counts in its own registers (shuffles); the chunk's record indices and the
//                  block's segment of sorted_ids come out in cell order.
//   accumulate   : thread (cell, x-slab a) walks the particles of ITS cell and accumulates the 16 stencil nodes
//                  (a, b, c) x (mass, momentum) in 32 packed fp32 pairs (FFMA2) -> the cross-particle reduction happens in
//                  registers, never between lanes.
//   fold         : the four cells of a z-column are folded with warp shuffles, x/y from a padded shared array; ONE vector
//                  red.global.add.v4.f32 per tile node (7^3 = 343 per block instead of 64 per particle), also into the
//                  neighbour slab's grid for nodes of a shared layer (PEER).
//   F-update     : (fused substep) the block's particles, after the write-back, into planes 4..10 of the other buffer.
// G2P (k_g2p_tile):  each warp fetches the 2x2x2 grid blocks of its block's tile with 128 row-wise cp.async.bulk copies (TMA,
//   completion on the warp's own mbarrier) into a linear, padded layout; each lane then owns one particle: separable gather
//   of v and the APIC matrix on packed pairs, advection, next substep's block key + histogram (and migration packing on slab
//   handles), and the write into the other particle buffer at its sorted rank (the physical re-sort that keeps the next
//   substep's loads coalesced).
#pragma once
#include "mpm_kernels.cuh"

namespace mpm {

#ifndef MPM_HOST_EMU
#define MPM_DYN_SMEM(name, al) extern __shared__ __align__(al) unsigned char name[]
#define MPM_SMEM_PROBE(site, key, ptr, bytes)          /* nothing in the product build */
#define MPM_SMEM_EPOCH()
#else
#define MPM_DYN_SMEM(name, al) unsigned char* name = emu_dyn_smem()
// tests/emu only: records which shared-memory words the lanes of a warp touch in one instruction (site, key = loop
// iteration) so that the emulator can count wavefronts under a bank model
#define MPM_SMEM_PROBE(site, key, ptr, bytes) emu::smem_probe(site, key, ptr, bytes)
#define MPM_SMEM_EPOCH() emu::smem_epoch()
#endif

// (the packed fp32 pair helpers pack2 / lo2 / hi2 / ffma2 / fadd2 / fmul2 live in mpm_math.cuh)
constexpr int P2G_T = 256;         // threads per CTA = 64 cells x 4 x-slabs
constexpr int P2G_PPT = 2;         // particles derived per thread per chunk
constexpr int P2G_CH = P2G_T * P2G_PPT;   // particles per chunk (a full 8-ppc block is one chunk)
struct P2GSmem {
    union {
        struct {
            float4 wx[P2G_CH];            // read as scalar component [a] (4 slab lanes of a cell hit 4 consecutive banks)
            float4 wy[P2G_CH], wz[P2G_CH];
            float4 qc[P2G_CH];            // (mass channel, c0.x, c0.y, c0.z)
            float4 hA0[P2G_CH], hA1[P2G_CH];   // h*A row-major entries 0..3, 4..7
            float hA8[P2G_CH];
            unsigned short order[P2G_CH];
        } c;
        float4 t1[4][4][4][42];           // phase 2: z-folded partial sums [cx][cy][a][b*7 + k]; the a-stride is padded
                                          // from 28 to 42 float4 (= 2 mod 8 slots) so that a quarter-warp's stores
                                          // (a = 0..3, two consecutive k) land in 8 different 16-byte bank groups
    } u;
    int cell_cnt[64];
    int4 work;
};
typedef P2GInG P2GIn;
// L2 prefetch: no register, no shared memory, one instruction per 128-byte line. P2G uses it twice per block (measured at 64 Mi,
// profiles/r2_experiments.md: P2G + F-update 5.09 -> 4.84 ms): the five planes the in-kernel F-update will read are requested
// while the derive phase has the particle's id in hand, and the six planes of the NEXT block's first chunk right after this
// block's fold, when its ids (requested before the fold) have arrived.
#ifndef MPM_HOST_EMU
MPM_DI void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
MPM_DI void prefetch_l2(const void*) {}
#endif
// first-chunk ids of a work item (chunks interleave the block's segment: slot q of chunk 0 is rank start + q * n_chunks)
MPM_DI void p2g_first_chunk_ids(const int4& wk, const int* __restrict__ sorted_ids, int t, int (&gid)[P2G_PPT], int& nch0) {
    const int nck0 = (wk.z + P2G_CH - 1) / P2G_CH;
    nch0 = wk.x < 0 ? 0 : (nck0 <= 1 ? wk.z : (wk.z + nck0 - 1) / nck0);      // (one chunk: no division by a variable)
#pragma unroll
    for (int u = 0; u < P2G_PPT; ++u) { const int q = t + u * P2G_T; gid[u] = q < nch0 ? sorted_ids[wk.y + q * nck0] : 0; }
}

// Thread t = cell*4 + a with cell = (cx*4 + cy)*4 + cz: within a warp the four cells of a z-column sit at lane stride 4.
// FUPD (the fused substep's default): after a block's tile has been written back, the same CTA runs the F-update
// (cpp:306-330) of the block's particles: it depends only on particle state of the START of the substep (B of the previous
// gather, FE, FP), never on the grid, so it can run anywhere between two gathers. Inside P2G its HBM streaming
// (240 B/particle, 2.4 ms as a kernel of its own) shares the SM with the other CTA's accumulation. Results go to planes
// 4..10 of the OTHER buffer at the particle's sorted rank, exactly where k_fupdate<true> puts them; the gather then runs
// without an F-update launch. Measured at 64 Mi (profiles/r2_ab_64M.md): P2G 3.53 + k_fupdate 2.37 -> 5.08 ms.
// PEER (mpm_substep_begin_peer; the multi-GPU default, 3.89 / 2.10 / 1.19 ms at 2 / 4 / 8 GPUs): the ghost-layer reduction of the slab
// decomposition done by this kernel itself. A tile node that lies in a block layer shared with a neighbouring slab (my
// ghost layer = the upper neighbour's first layer; my first layer = the lower neighbour's ghost layer) is added to the
// local copy AND, with the same vector red, to the neighbour's copy through its peer-mapped grid (NVLink atomics execute
// at the owning GPU's L2), so that after both P2G kernels both copies hold the complete sums: no halo message, no pack /
// add kernels, and the remote reds of a block overlap the accumulation of the next ones.
struct PeerLayers { float4* dn; float4* up; };      // lower neighbour's ghost layer, upper neighbour's first layer (or null)
// The work list is in block order (lowest i first). With remote reds the blocks of BOTH boundary layers go first -- tickets
// alternate between the front and the back of the list -- so that the NVLink traffic overlaps the interior blocks and the
// neighbours' waits end early.
MPM_DI int peer_work_order(int ticket, int n_work) { return (ticket & 1) ? n_work - 1 - (ticket >> 1) : (ticket >> 1); }
template <int MODE, int FUPD = 0 /* 0 none, 1 bit-faithful, 2 tolerance form */, bool PEER = false, int W = 4 /* stencil: 4 = cubic (compile time), 3 = quadratic (compile time, zero column skipped), 0 = either (run time, 4-wide) */>
__global__ void __launch_bounds__(P2G_T, 2)
k_p2g_tile(Planes P, int* __restrict__ sorted_ids, const int4* __restrict__ pblock_list, DevCounters* dc,
           float4* __restrict__ grid, GridDims gd, SimConst sc, float dt, Planes Nx, PeerLayers peer = PeerLayers{ nullptr, nullptr }) {
    MPM_DYN_SMEM(smem_raw, 16);
    P2GSmem& S = *reinterpret_cast<P2GSmem*>(smem_raw);
    constexpr int SQ = W == 3 ? 1 : (W == 4 ? 0 : 2);      // stencil selection of the weight functions
    constexpr int WN = W == 3 ? 3 : 4;                     // stencil nodes per axis that can carry weight
    const int t = threadIdx.x, lane = t & 31;
    const int my_cell = t >> 2, my_a = t & 3;
    const int my_cx = my_cell >> 4, my_cy = (my_cell >> 2) & 3, my_cz = my_cell & 3;
    const float fa = (float)my_a;
    const int n_work = dc->n_active_pblocks;
    // work pipeline of thread 0: the atomic ticket is drawn two blocks ahead and its work item (block id, first rank,
    // count) one block ahead, so neither latency is ever waited for; everybody else pre-loads the ids of the next
    // block's first chunk during phase 2.
    int w_ticket = 0;
    int4 wk_reg = make_int4(-1, 0, 0, 0);
    const int4 no_work = make_int4(-1, 0, 0, 0);
    if (t == 0) {
        const int w0 = atomicAdd(&dc->work_a, 1);
        S.work = w0 < n_work ? pblock_list[PEER ? peer_work_order(w0, n_work) : w0] : no_work;
        w_ticket = atomicAdd(&dc->work_a, 1);
    }
    __syncthreads();
    int gid_pref[P2G_PPT];
    { int nch0; p2g_first_chunk_ids(S.work, sorted_ids, t, gid_pref, nch0); }
#ifdef MPM_P2G_PROFILE
    long long prof_t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, prof_c = clock64();
#define MPM_PROF(i) do { const long long c_ = clock64(); prof_t[i] += c_ - prof_c; prof_c = c_; } while (0)
#else
#define MPM_PROF(i)
#endif
    for (;;) {
        const int4 wk = S.work;
        if (t < 64) S.cell_cnt[t] = 0;
        __syncthreads();          // also: everyone has read S.work, and phase 2b of the previous block is done with t1
        MPM_PROF(0);              // top barrier
        if (wk.x < 0) break;
        const int start = wk.y, cnt = wk.z;
        const int pbk = wk.w & (PB_COORD_MAX - 1), pbj = (wk.w >> PB_COORD_BITS) & (PB_COORD_MAX - 1), pbi = (wk.w >> (2 * PB_COORD_BITS)) + gd.lo;   // global block coords
        f32x2_t AM[8], AX[8], AY[8], AZ[8];     // 16 stencil nodes x (mass, momentum): channel-major fp32 pairs over the z-nodes (c, c+1)
#pragma unroll
        for (int i = 0; i < 8; ++i) AM[i] = AX[i] = AY[i] = AZ[i] = 0ull;

        // chunks take every n_chunks-th particle of the (cell-ordered) block segment, so that every chunk holds a
        // share of every cell and all 64 (cell) thread groups have work in phase 1
        const int n_chunks = (cnt + P2G_CH - 1) / P2G_CH;
        for (int ck = 0; ck < n_chunks; ++ck) {
            const int nch = n_chunks == 1 ? cnt : (cnt - ck + n_chunks - 1) / n_chunks;          // slots ck, ck+n_chunks, ...
            MPM_SMEM_EPOCH();
            if (ck > 0) {
                if (t < 64) S.cell_cnt[t] = 0;
                __syncthreads();
            }
            // ---- derive per-particle data (P2G_PPT particles per thread, loads of both issued back to back) ----
            int cell_rank[P2G_PPT], gids[P2G_PPT];   // cell (low 8 bits) and rank inside the cell: the counting atomic's return value
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u) {
                const int q = t + u * P2G_T;
                cell_rank[u] = 0; gids[u] = 0;
                if (q < nch) {
                    const int gid = ck == 0 ? gid_pref[u] : sorted_ids[start + ck + q * n_chunks];
                    const P2GIn in = p2g_load_planes<MODE>(P, gid);
                    const float4 xm = in.xm;
                    float mch, a0[3], A[9];
                    p2g_coeffs<MODE>(in, sc.dinv, dt, mch, a0, A);
                    float wx[4], wy[4], wz[4];
                    // cell index and weights from ONE pos/h quotient per axis (the same operations as cell_of + axis_weights,
                    // which form the quotient twice: identical bits, 68 fewer instructions per particle)
                    const int cx = cell_and_weights_t<SQ>(xm.x, sc.pd, wx), cy = cell_and_weights_t<SQ>(xm.y, sc.pd, wy), cz = cell_and_weights_t<SQ>(xm.z, sc.pd, wz);
                    const float d0 = (float)(cx - 1) * sc.h - xm.x, d1 = (float)(cy - 1) * sc.h - xm.y, d2 = (float)(cz - 1) * sc.h - xm.z;
                    MPM_SMEM_PROBE(1, u, &S.u.c.wx[q], 16); MPM_SMEM_PROBE(2, u, &S.u.c.hA8[q], 4);
                    S.u.c.wx[q] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                    S.u.c.wy[q] = make_float4(wy[0], wy[1], wy[2], wy[3]);
                    S.u.c.wz[q] = make_float4(wz[0], wz[1], wz[2], wz[3]);
                    S.u.c.qc[q] = make_float4(mch, a0[0] + A[0] * d0 + A[1] * d1 + A[2] * d2, a0[1] + A[3] * d0 + A[4] * d1 + A[5] * d2,
                                              a0[2] + A[6] * d0 + A[7] * d1 + A[8] * d2);
                    S.u.c.hA0[q] = make_float4(A[0] * sc.h, A[1] * sc.h, A[2] * sc.h, A[3] * sc.h);
                    S.u.c.hA1[q] = make_float4(A[4] * sc.h, A[5] * sc.h, A[6] * sc.h, A[7] * sc.h);
                    S.u.c.hA8[q] = A[8] * sc.h;
                    gids[u] = gid;
                    if (FUPD) {      // the F-update of this block reads these five planes a few microseconds from now: pull them into L2
                        prefetch_l2(&P.p[6][gid]); prefetch_l2(&P.p[7][gid]); prefetch_l2(&P.p[8][gid]); prefetch_l2(&P.p[9][gid]); prefetch_l2(&P.p[10][gid]);
                    }
                    const int lc = (((cx - 1) - 4 * pbi) * 4 + ((cy - 1) - 4 * pbj)) * 4 + ((cz - 1) - 4 * pbk);
                    cell_rank[u] = lc | (atomicAdd(&S.cell_cnt[lc], 1) << 8);
                }
            }
            __syncthreads();
            MPM_PROF(1);          // derive (+ barrier)
            // ---- counting sort of the chunk by cell (64 bins): EVERY warp scans the 64 counts in its own registers (two bins
            // per lane, five shuffles), so no warp waits for a scanning warp and no barrier separates the scan from its use ----
            const int c0 = S.cell_cnt[2 * lane], c1 = S.cell_cnt[2 * lane + 1];
            int inc = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            const int ex = inc - (c0 + c1);          // exclusive prefix of bin 2*lane; bin 2*lane+1 starts at ex + c0
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u) {
                const int q = t + u * P2G_T;
                const int c = cell_rank[u] & 255;
                const int e = __shfl_sync(0xffffffffu, ex, c >> 1), f = __shfl_sync(0xffffffffu, c0, c >> 1);
                if (q < nch) {           // sorted slot = start of my cell + my rank in it (no second round of atomics)
                    const int slot = e + ((c & 1) ? f : 0) + (cell_rank[u] >> 8);
                    MPM_SMEM_PROBE(5, u, &S.u.c.order[slot], 2);
                    S.u.c.order[slot] = (unsigned short)q;
                    // keep the block's segment of sorted_ids (approximately) cell-ordered: the G2P lanes of a warp then share
                    // cells (smem broadcasts) and the re-sorted particle buffer stays cell-coherent for the next substep
                    // (measured, profiles/r2_experiments.md: on cell-coherent order the rewrite costs nothing, on a shuffled upload
                    // the gather runs 2.2x slower without it -- so it stays on in every substep. Moving the F-update into this
                    // derive phase, which needs the ids left alone, was tried and was SLOWER: 5.32 vs 5.09 ms at 64 Mi.)
                    sorted_ids[start + ck + slot * n_chunks] = gids[u];
                }
            }
            int i0, i1;
            {
                const int e = __shfl_sync(0xffffffffu, ex, my_cell >> 1), f = __shfl_sync(0xffffffffu, c0, my_cell >> 1), g = __shfl_sync(0xffffffffu, c1, my_cell >> 1);
                i0 = e + ((my_cell & 1) ? f : 0);
                i1 = i0 + ((my_cell & 1) ? g : f);
            }
            __syncthreads();
            MPM_PROF(2);          // counting sort + id rewrite
            // ---- phase 1: register accumulation over the particles of my cell ----
            // The 8 cells of a warp read 8 different records per iteration; a record's 16-byte bank group is its slot
            // mod 8. Once the ids are cell-ordered (after the first substep) a cell's records sit in consecutive slots, and
            // in an 8-particles-per-cell scene every cell's run starts at a multiple of 8: walking all runs from their
            // first record puts the 8 reads of EVERY iteration into the same bank group (8 wavefronts per LDS instead of
            // 1-2; 2.4 k of the 2.9 k bank-conflict wavefronts per block ncu counted at 64 Mi, profiles/r1_analysis.md).
            // So cell c starts its (cyclic) walk at record c mod 8: with runs of 8 the reads of an iteration then hit 8
            // different bank groups, and ragged runs are no worse off than before. Only the summation order changes.
            int i = i0;
            if (sc.p2g_rotate && i1 - i0 > 1) i += (my_cell & 7) % (i1 - i0);
            int pi = i1 > i0 ? S.u.c.order[i] : 0;          // record index, fetched one iteration ahead of its record
            // (W = 3, quadratic stencil: the x-slab a = 3 and the y-row b = 3 carry zero weights and are skipped; the z pairs
            // keep their zero fourth weight)
            const int n_vis = (WN == 4 || my_a < WN) ? i1 - i0 : 0;
#pragma unroll 1
            for (int k = 0; k < n_vis; ++k) {
                MPM_SMEM_PROBE(10, k, &S.u.c.order[i], 2);
                i = (i + 1 == i1) ? i0 : i + 1;
                MPM_SMEM_PROBE(11, k, &reinterpret_cast<const float*>(&S.u.c.wx[pi])[my_a], 4);
                MPM_SMEM_PROBE(12, k, &S.u.c.wy[pi], 16); MPM_SMEM_PROBE(13, k, &S.u.c.wz[pi], 16); MPM_SMEM_PROBE(14, k, &S.u.c.qc[pi], 16);
                MPM_SMEM_PROBE(15, k, &S.u.c.hA0[pi], 16); MPM_SMEM_PROBE(16, k, &S.u.c.hA1[pi], 16); MPM_SMEM_PROBE(17, k, &S.u.c.hA8[pi], 4);
                const float wxa = reinterpret_cast<const float*>(&S.u.c.wx[pi])[my_a];
                const float4 wy = S.u.c.wy[pi], wz = S.u.c.wz[pi], qc = S.u.c.qc[pi], h0 = S.u.c.hA0[pi], h1 = S.u.c.hA1[pi];
                const float h8 = S.u.c.hA8[pi];
                pi = S.u.c.order[i];                       // (slot i is inside the cell's run, possibly wrapped: always valid)
                // value(a,b,c)_r = c0_r + a*hA[r][0] + b*hA[r][1] + c*hA[r][2]
                const float bx = qc.y + fa * h0.x, by = qc.z + fa * h0.w, bz = qc.w + fa * h1.z;
                const float wyv[4] = { wy.x, wy.y, wy.z, wy.w };
                // pairs over two consecutive z-nodes (c, c+1): weights W = wab * (wz_c, wz_c+1) straight from the aligned
                // halves of the wz record, channel values v0 + (c, c+1) * step by one FFMA2 with broadcast operands. The affine
                // value at node c is fma(c, step, v0) (one rounding). Per node pair 1 FMUL2 + 7 FFMA2 where scalar code needs
                // 2 FMUL + 8 FFMA + 6 FADD (measured 3.87 -> 3.60 ms at 64 Mi; the scalar form is gone).
                const f32x2_t wzp[2] = { pack2(wz.x, wz.y), pack2(wz.z, wz.w) };
                const f32x2_t cp2[2] = { pack2(0.f, 1.f), pack2(2.f, 3.f) };
                const f32x2_t mm = pack2(qc.x, qc.x), sx = pack2(h0.z, h0.z), sy = pack2(h1.y, h1.y), sz = pack2(h8, h8);
#pragma unroll
                for (int bb = 0; bb < WN; ++bb) {
                    const float wab = wxa * wyv[bb];
                    const float vx = bx + (float)bb * h0.y, vy = by + (float)bb * h1.x, vz = bz + (float)bb * h1.w;
                    const f32x2_t wab2 = pack2(wab, wab), vx2 = pack2(vx, vx), vy2 = pack2(vy, vy), vz2 = pack2(vz, vz);
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp) {
                        const f32x2_t Wg = fmul2(wzp[cp], wab2);
                        ffma2_acc(AM[bb * 2 + cp], Wg, mm);
                        ffma2_acc(AX[bb * 2 + cp], Wg, ffma2(cp2[cp], sx, vx2));
                        ffma2_acc(AY[bb * 2 + cp], Wg, ffma2(cp2[cp], sy, vy2));
                        ffma2_acc(AZ[bb * 2 + cp], Wg, ffma2(cp2[cp], sz, vz2));
                    }
                }
            }
            MPM_PROF(3);          // accumulation loop
            __syncthreads();
            MPM_PROF(4);          // barrier after the accumulation
        }
        float4 acc[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[2 * i] = make_float4(lo2(AM[i]), lo2(AX[i]), lo2(AY[i]), lo2(AZ[i]));
            acc[2 * i + 1] = make_float4(hi2(AM[i]), hi2(AX[i]), hi2(AY[i]), hi2(AZ[i]));
        }
        if (t == 0) wk_reg = w_ticket < n_work ? pblock_list[PEER ? peer_work_order(w_ticket, n_work) : w_ticket] : no_work;   // issued here, stored below
        // ---- phase 2a: fold the four cells of a z-column with warp shuffles (lanes at stride 4) ----
        // node k = cz + c; this lane ends up owning k = cz (slot 0) and k = cz + 4 (slot 1, cz <= 2)
        float4 s0[4], s1[4];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) { s0[bb] = acc[bb * 4]; s1[bb] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) {
            const int src = (lane & ~12) | (((my_cz - cc) & 3) << 2);     // the lane holding cell cz' = (cz - cc) mod 4
            const bool lo = cc <= my_cz;                                   // cz' = cz - cc  -> k = cz ; else cz' = cz+4-cc -> k = cz+4
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                float4 v;
                v.x = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].x, src); v.y = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].y, src);
                v.z = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].z, src); v.w = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].w, src);
                if (lo) { s0[bb].x += v.x; s0[bb].y += v.y; s0[bb].z += v.z; s0[bb].w += v.w; }
                else { s1[bb].x += v.x; s1[bb].y += v.y; s1[bb].z += v.z; s1[bb].w += v.w; }
            }
        }
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            MPM_SMEM_PROBE(20, bb, &S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz], 16);
            S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz] = s0[bb];
            if (my_cz < 3) {
                MPM_SMEM_PROBE(21, bb, &S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz + 4], 16);
                S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz + 4] = s1[bb];
            }
        }
        if (t == 0) { S.work = wk_reg; w_ticket = atomicAdd(&dc->work_a, 1); }
        __syncthreads();
        MPM_PROF(5);              // z-fold (shuffles) + patch stores + barrier
        { int nchn; p2g_first_chunk_ids(S.work, sorted_ids, t, gid_pref, nchn); }       // ids of the next block's first chunk: in flight while this block's tile is reduced and written back
        // ---- phase 2b: fold x and y from smem (<= 16 terms per tile node), one vector red per node ----
        // (visiting the nodes sorted by fold length to even out the trip counts was measured SLOWER than natural order:
        // the contiguous k-runs of natural order matter more to the smem pipe than the divergence costs)
        for (int n = t; n < 343; n += P2G_T) {
            const int ni = n / 49, nj = (n / 7) % 7, nk = n % 7;
            // the <= 4 x 4 patches that cover this node, as a fixed, fully unrolled walk with predicated loads: all loads of a
            // node are in flight together (the bounded loops this replaces waited for one shared-memory round trip per term)
            float4 v[16];
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
#pragma unroll
                for (int dy = 0; dy < 4; ++dy) {
                    const int cx = ni - dx, cy = nj - dy;          // patch of cell (cx, cy, .), its local node (dx, dy, nk)
                    const bool ok = cx >= 0 && cx <= 3 && cy >= 0 && cy <= 3;
                    MPM_SMEM_PROBE(30, ((n / P2G_T) * 4 + dx) * 4 + dy, &S.u.t1[ok ? cx : 0][ok ? cy : 0][dx][dy * 7 + nk], 16);
                    v[dx * 4 + dy] = ok ? S.u.t1[cx][cy][dx][dy * 7 + nk] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) { sum.x += v[i].x; sum.y += v[i].y; sum.z += v[i].z; sum.w += v[i].w; }
            if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f) {
                const size_t idx = node_index(gd, 4 * pbi + ni, 4 * pbj + nj, 4 * pbk + nk);
                atomicAdd(&grid[idx], sum);
                if (PEER) {
                    const int layer = ((4 * pbi + ni) >> 2) - gd.lo;                   // block layer of the node inside my slab
                    const size_t layer_nodes = (size_t)gd.nbj * gd.nbk * 64;
                    if (layer == gd.hi - gd.lo) { if (peer.up) atomicAdd(&peer.up[idx - (size_t)layer * layer_nodes], sum); }
                    else if (layer == 0) { if (peer.dn) atomicAdd(&peer.dn[idx], sum); }
                }
            }
        }
        {   // the next block's first chunk: its ids have arrived during the fold; pull its six P2G planes into L2 ahead of the derive loads
            const int nck_n = (S.work.z + P2G_CH - 1) / P2G_CH;
            const int nch_n = S.work.x < 0 ? 0 : (nck_n <= 1 ? S.work.z : (S.work.z + nck_n - 1) / nck_n);
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u)
                if (t + u * P2G_T < nch_n) {
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (k == 0 || (k <= 3 ? MODE != P2G_FORCE : MODE != P2G_MOMENTUM)) prefetch_l2(&P.p[k][gid_pref[u]]);      // the planes p2g_load_planes<MODE> reads
                }
        }
        MPM_PROF(6);              // x/y fold + reds
        if (FUPD) {
            // F-update of this block's particles, two per thread and round with both particles' loads issued first.
            // sorted_ids[start .. start+cnt) is final here (the cell-ordering writes above are behind CTA barriers).
#pragma unroll 1
            for (int q = t; q < cnt; q += 2 * P2G_T) {
                const int j0 = start + q, j1 = j0 + P2G_T;
                const bool two = q + P2G_T < cnt;
                const int p0 = sorted_ids[j0], p1 = two ? sorted_ids[j1] : p0;
                const FUpdIn in0 = fupd_load(P, p0);
                const FUpdIn in1 = fupd_load(P, p1);
                fupd_compute_store<FUPD == 2>(in0, Nx, j0, dc, sc, dt);
                if (two) fupd_compute_store<FUPD == 2>(in1, Nx, j1, dc, sc, dt);
            }
            MPM_PROF(7);          // in-kernel F-update
        }
    }
#ifdef MPM_P2G_PROFILE
    if (t == 0) for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&dc->prof[i]), (unsigned long long)prof_t[i]);
#endif
#undef MPM_PROF
}

// ---------------------------------------------------------------------------------------------------------
// G2P
// ---------------------------------------------------------------------------------------------------------
#ifndef MPM_HOST_EMU
MPM_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
MPM_DI void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MPM_DI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MPM_DI void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
MPM_DI void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}
MPM_DI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#else
MPM_DI void mbar_init(unsigned long long* bar, int count) { emu::mbar_init(bar, count); }
MPM_DI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu::mbar_expect_tx(bar, bytes); }
MPM_DI void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) { emu::tma_load_1d(dst_smem, src_gmem, bytes, bar); }
MPM_DI void mbar_wait(unsigned long long* bar, unsigned phase) { emu::mbar_wait(bar, phase); }
MPM_DI void mbar_fence_init() {}
#endif

#ifndef G2P_MIN_CTAS
#define G2P_MIN_CTAS 2
#endif
constexpr int G2P_WARPS = 8, G2P_T = G2P_WARPS * 32;
// The 2x2x2-grid-block tile of a particle block, re-laid out LINEARLY in shared memory: slot(ti,tj,tk) = ti*73 + tj*9 + tk
// (rows padded 8 -> 9, planes 72 -> 73, so the 16-byte bank group of a node is (ti+tj+tk) mod 8). The 64 stencil reads of a
// particle then are LDS.128 [base + immediate] with no per-node address arithmetic (the blocked tile of round 1 spent ~190
// integer instructions per particle on addresses). The re-layout is done by the copy engine: 128 bulk copies of 64 B
// (one 4-node k-row each), 4 per lane, completing on the warp's own mbarrier.
constexpr int G2P_LIN_ROW = 9, G2P_LIN_PLANE = 73, G2P_LIN_SLOTS = 8 * G2P_LIN_PLANE;
struct G2PSmem {
    float4 tile[G2P_WARPS][G2P_LIN_SLOTS];
    unsigned long long bar[G2P_WARPS];
};

// Warp-per-block gather: every warp owns a whole particle block at a time (its own TMA-loaded tile, its own mbarrier), so
// there is no CTA-wide barrier and no ragged-tail idling beyond the last 32-particle slice of a block.
// The separable gather is issued as packed fp32 pairs: sm_100a has FFMA2 (PTX fma.rn.f32x2), two IEEE fma.rn per lane per
// instruction, and ptxas folds a duplicated {x, x} operand into a scalar broadcast, so (s0, s1) += (wz, wz*dz) * n needs ONE
// instruction instead of two: 576 FFMA per particle become 252 FFMA2 + 84 FFMA.
// Measured at 64 Mi particles (profiles/r2_ab_64M.md): blocked tile + scalar FMA 2.40 ms, linear tile 2.22, packed pairs
// 2.19, both 1.99 -> only this form is kept.
// FLAGS: G2P_GATHER always, optionally G2P_ADVECT, G2P_REORDER, G2P_HIST (the F-update runs in P2G or in k_fupdate).
// G2P_GRADW (implicit time integration, mpm_implicit.cuh): the same separable gather with (w, dw/dx) pairs instead of
// (w, w * (x_i - x_p)) pairs, i.e. the gradient of the field in `grid` with the true B-spline derivative (hpp:59-71); the
// result I + dt * grad v goes to `aux` (3 float4 per sorted rank) and no particle plane is written.
template <int FLAGS, int W = 4 /* stencil: 4 = cubic (compile time), 3 = quadratic (compile time: 27 tile reads instead of 64), 0 = either (run time) */>
__global__ void __launch_bounds__(G2P_T, G2P_MIN_CTAS)
k_g2p_tile(Planes cur, Planes nxt, const int* __restrict__ sorted_ids, const int4* __restrict__ pblock_list, DevCounters* dc,
           const float4* __restrict__ grid, GridDims gd, SimConst sc, float dt, int* __restrict__ key_out = nullptr, int* __restrict__ blk_count = nullptr,
           MigOut mo = MigOut{ nullptr, nullptr, 0 }, float4* __restrict__ aux = nullptr) {
    MPM_DYN_SMEM(g2p_smem_raw, 128);
    G2PSmem& S = *reinterpret_cast<G2PSmem*>(g2p_smem_raw);
    constexpr int SQ = W == 3 ? 1 : (W == 4 ? 0 : 2), WN = W == 3 ? 3 : 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_work = dc->n_active_pblocks;
    float4* __restrict__ tile = S.tile[wid];
    unsigned long long* bar = &S.bar[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    unsigned phase = 0;
    for (;;) {
        int w = 0, bc = 0, start = 0, cnt = 0;
        if (lane == 0) {
            w = atomicAdd(&dc->work_b, 1);
            if (w < n_work) {
                const int4 wk = pblock_list[w];
                bc = wk.w; start = wk.y; cnt = wk.z;
                mbar_expect_tx(bar, 8 * 1024);
            }
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_work) break;
        bc = __shfl_sync(0xffffffffu, bc, 0); start = __shfl_sync(0xffffffffu, start, 0); cnt = __shfl_sync(0xffffffffu, cnt, 0);
        const int pbk = bc & (PB_COORD_MAX - 1), pbj = (bc >> PB_COORD_BITS) & (PB_COORD_MAX - 1), pbi = (bc >> (2 * PB_COORD_BITS)) + gd.lo;
        MPM_SMEM_EPOCH();
        {
            // lane 0 has armed the barrier with the byte count (ordered before the copies by the shuffles above);
            // every lane issues 4 of the 128 row copies: row = lane & 15 of grid blocks d = (lane >> 4) + 2m
            const int row = lane & 15, li = row >> 2, lj = row & 3;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int d = (lane >> 4) + 2 * m, di = d >> 2, dj = (d >> 1) & 1, dk = d & 1;
                const size_t gb = ((size_t)(pbi - gd.lo + di) * gd.nbj + pbj + dj) * gd.nbk + pbk + dk;
                tma_load_1d(&tile[(di * 4 + li) * G2P_LIN_PLANE + (dj * 4 + lj) * G2P_LIN_ROW + dk * 4], grid + gb * 64 + row * 4, 64, bar);
            }
        }
        // software pipeline over the 32-particle slices: ids two slices ahead, positions one slice ahead, so neither
        // the sorted_ids -> position dependent-load chain nor the position load is waited for
        int p = (lane < cnt) ? sorted_ids[start + lane] : -1;
        int p_nn = (32 + lane < cnt) ? sorted_ids[start + 32 + lane] : -1;
        float4 a0 = (p >= 0) ? cur.p[0][p] : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(bar, phase);
        phase ^= 1;
        for (int base = 0; base < cnt; base += 32) {
            const int j = start + base + lane;
            const bool active = base + lane < cnt;
            const int p_cur = p;
            const float4 a_cur = a0;
            p = p_nn;
            a0 = (p >= 0) ? cur.p[0][p] : make_float4(0.f, 0.f, 0.f, 0.f);
            p_nn = (base + 64 + lane < cnt) ? sorted_ids[j + 64] : -1;
            int key_new = KEY_DEAD;                                 // (G2P_HIST) block key of the advected position
            if (active) {
                struct { float x[3], m, v[3], B[9]; } r;     // the gather touches only x (in) and x, v, B (out)
                r.x[0] = a_cur.x; r.x[1] = a_cur.y; r.x[2] = a_cur.z; r.m = a_cur.w;
                {
                    float wx[4], wy[4], wz[4];
                    // one pos/h quotient per axis feeds the cell index and the weights (same bits as cell_of + axis_weights)
                    int cx, cy, cz;
                    float dwx[4], dwy[4], dwz[4];
                    if constexpr ((FLAGS & G2P_GRADW) != 0) {
                        cx = cell_of_t<SQ>(r.x[0], sc.pd); cy = cell_of_t<SQ>(r.x[1], sc.pd); cz = cell_of_t<SQ>(r.x[2], sc.pd);
                        axis_weights_and_derivatives(r.x[0], sc.pd, cx, wx, dwx);
                        axis_weights_and_derivatives(r.x[1], sc.pd, cy, wy, dwy);
                        axis_weights_and_derivatives(r.x[2], sc.pd, cz, wz, dwz);
                    } else {
                        cx = cell_and_weights_t<SQ>(r.x[0], sc.pd, wx); cy = cell_and_weights_t<SQ>(r.x[1], sc.pd, wy); cz = cell_and_weights_t<SQ>(r.x[2], sc.pd, wz);
                    }
                    const int ox = (cx - 1) - 4 * pbi, oy = (cy - 1) - 4 * pbj, oz = (cz - 1) - 4 * pbk;
                    const float4* __restrict__ lin = tile + (ox * G2P_LIN_PLANE + oy * G2P_LIN_ROW + oz);
                    // pairs: S = (s0, s1) <- (wz, wz dz) * n ;  T = (t0, t1y) <- (wy, wy dy) * s0 ;  V = (v, Bx) <- (wx, wx dx) * t0
                    f32x2_t WX[4], WY[4], WZ[4], V[3] = { 0ull, 0ull, 0ull };
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if constexpr ((FLAGS & G2P_GRADW) != 0) {
                            const float ih = 1.0f / sc.h;
                            WX[a] = pack2(wx[a], dwx[a] * ih); WY[a] = pack2(wy[a], dwy[a] * ih); WZ[a] = pack2(wz[a], dwz[a] * ih);
                        } else {
                            WX[a] = pack2(wx[a], wx[a] * ((float)(cx - 1 + a) * sc.h - r.x[0]));
                            WY[a] = pack2(wy[a], wy[a] * ((float)(cy - 1 + a) * sc.h - r.x[1]));
                            WZ[a] = pack2(wz[a], wz[a] * ((float)(cz - 1 + a) * sc.h - r.x[2]));
                        }
                    }
                    float By[3] = { 0, 0, 0 }, Bz[3] = { 0, 0, 0 };
#pragma unroll
                    for (int a = 0; a < WN; ++a) {
                        f32x2_t T[3] = { 0ull, 0ull, 0ull };
                        float t1z[3] = { 0, 0, 0 };
#pragma unroll
                        for (int bb = 0; bb < WN; ++bb) {
                            f32x2_t Sx[3] = { 0ull, 0ull, 0ull };
#pragma unroll
                            for (int cc = 0; cc < WN; ++cc) {
                                MPM_SMEM_PROBE(40, (base / 32) * 64 + (a * 4 + bb) * 4 + cc, &lin[a * G2P_LIN_PLANE + bb * G2P_LIN_ROW + cc], 16);
                                const float4 n = lin[a * G2P_LIN_PLANE + bb * G2P_LIN_ROW + cc];
                                Sx[0] = ffma2(WZ[cc], pack2(n.y, n.y), Sx[0]);
                                Sx[1] = ffma2(WZ[cc], pack2(n.z, n.z), Sx[1]);
                                Sx[2] = ffma2(WZ[cc], pack2(n.w, n.w), Sx[2]);
                            }
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const float s0q = lo2(Sx[q]);
                                T[q] = ffma2(WY[bb], pack2(s0q, s0q), T[q]);
                                t1z[q] += lo2(WY[bb]) * hi2(Sx[q]);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float t0q = lo2(T[q]);
                            V[q] = ffma2(WX[a], pack2(t0q, t0q), V[q]);
                            By[q] += lo2(WX[a]) * hi2(T[q]); Bz[q] += lo2(WX[a]) * t1z[q];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) { r.v[q] = lo2(V[q]); r.B[q] = hi2(V[q]); r.B[3 + q] = By[q]; r.B[6 + q] = Bz[q]; }
                }
                if (FLAGS & G2P_ADVECT) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {                  // cpp:344-350, 381-388
                        float x = add_rn(r.x[a], mul_rn(r.v[a], dt));
                        if (x < sc.pos_lo) x = sc.pos_lo;
                        if (sc.pos_hi[a] < x) x = sc.pos_hi[a];
                        r.x[a] = x;
                    }
                }
                const Planes& D = (FLAGS & G2P_REORDER) ? nxt : cur;
                const int q = (FLAGS & G2P_REORDER) ? j : p_cur;
                if constexpr ((FLAGS & G2P_GRADW) != 0) {      // I + dt grad v (column-major like every 3x3 here), by sorted rank
                    aux[3 * (size_t)j + 0] = make_float4(fmaf(dt, r.B[0], 1.0f), dt * r.B[1], dt * r.B[2], dt * r.B[3]);
                    aux[3 * (size_t)j + 1] = make_float4(fmaf(dt, r.B[4], 1.0f), dt * r.B[5], dt * r.B[6], dt * r.B[7]);
                    aux[3 * (size_t)j + 2] = make_float4(fmaf(dt, r.B[8], 1.0f), 0.0f, 0.0f, 0.0f);
                } else {
                if (FLAGS & (G2P_ADVECT | G2P_REORDER)) D.p[0][q] = make_float4(r.x[0], r.x[1], r.x[2], r.m);
                D.p[1][q] = make_float4(r.B[0], r.B[1], r.B[2], r.B[3]);
                D.p[2][q] = make_float4(r.B[4], r.B[5], r.B[6], r.B[7]);
                D.p[3][q] = make_float4(r.B[8], r.v[0], r.v[1], r.v[2]);
                }
                if (FLAGS & G2P_HIST) {
                    // next substep's binning, first half, done here where the new position is in registers: block key of slot j
                    // of the re-sorted buffer (what k_bin_count would re-read P0 for); a particle that left the slab goes
                    // straight into the packed migration buffer (planes 4..10 of this slot were written by the F-update earlier
                    // in the substep) and retires its slot
                    int cells[3];
                    key_new = particle_key<SQ>(make_float4(r.x[0], r.x[1], r.x[2], r.m), gd, sc.pd, cells);
                    if (key_new > gd.n_pblocks && mig_try_pack(mo, key_new == gd.n_pblocks + 2, D, q, dc)) key_new = KEY_DEAD;
                    key_out[j] = key_new;
                }
            }
            if (FLAGS & G2P_HIST) {        // warp-aggregated histogram of the new keys
                const unsigned peers = __match_any_sync(0xffffffffu, key_new);
                if (key_new >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&blk_count[key_new], __popc(peers));
            }
        }
        __syncwarp();     // every lane is done with the tile before the next bulk copy overwrites it
    }
}

// parked (out-of-grid) particles ride along unchanged through a re-sorting G2P
__global__ void k_copy_parked(Planes cur, Planes nxt, const int* __restrict__ sorted_ids, DevCounters* dc,
                              GridDims gd, PosDiv pd, int* __restrict__ key_out, int* __restrict__ blk_count, MigOut mo) {
    // after this re-sort the new buffer holds n_sorted contiguous slots (nothing in this kernel reads n_slots)
    if (blockIdx.x == 0 && threadIdx.x == 0) dc->n_slots = dc->n_sorted;
    const int n_sorted = dc->n_sorted;
    for (int j = dc->n_binned + blockIdx.x * blockDim.x + threadIdx.x; j < n_sorted; j += gridDim.x * blockDim.x) {      // (usually empty)
        const int p = sorted_ids[j];
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) nxt.p[k][j] = cur.p[k][p];
        if (key_out) {
            // (fused binning) these particles did not move: parked ones stay parked; a leaver that found the migration buffer
            // full last substep is offered to the neighbour again
            int cells[3];
            int k = particle_key(nxt.p[0][j], gd, pd, cells);
            if (k > gd.n_pblocks && mig_try_pack(mo, k == gd.n_pblocks + 2, nxt, j, dc)) k = KEY_DEAD;
            key_out[j] = k;
            if (k >= 0) atomicAdd(&blk_count[k], 1);
        }
    }
}

#if !defined(MPM_HOST_EMU) || defined(MPM_HOST_EMU_API)        // host launch code (nvcc; or the whole-library emulation build of tests/emu)
inline cudaError_t tile_kernels_init() {
    cudaError_t e;
#define MPM_SET_SMEM(K, T) if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(T))) != cudaSuccess) return e
#define MPM_SET_P2G(MODE, FU, PE) MPM_SET_SMEM((k_p2g_tile<MODE, FU, PE, 4>), P2GSmem); MPM_SET_SMEM((k_p2g_tile<MODE, FU, PE, 0>), P2GSmem)
    MPM_SET_P2G(P2G_MOMENTUM, 0, false); MPM_SET_P2G(P2G_FORCE, 0, false); MPM_SET_P2G(P2G_FUSED, 1, false);
    MPM_SET_P2G(P2G_FUSED, 0, true); MPM_SET_P2G(P2G_FUSED, 1, true); MPM_SET_P2G(P2G_FUSED, 2, true);
#undef MPM_SET_P2G
    MPM_SET_SMEM((k_p2g_tile<P2G_FUSED, 0, false, 4>), P2GSmem); MPM_SET_SMEM((k_p2g_tile<P2G_FUSED, 2, false, 4>), P2GSmem);
    MPM_SET_SMEM((k_p2g_tile<P2G_FUSED, 0, false, 3>), P2GSmem); MPM_SET_SMEM((k_p2g_tile<P2G_FUSED, 2, false, 3>), P2GSmem);
    MPM_SET_SMEM((k_g2p_tile<G2P_GATHER, 4>), G2PSmem); MPM_SET_SMEM((k_g2p_tile<G2P_GATHER, 0>), G2PSmem);
    MPM_SET_SMEM((k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER, 4>), G2PSmem); MPM_SET_SMEM((k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER, 0>), G2PSmem);
    MPM_SET_SMEM((k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER | G2P_HIST, 4>), G2PSmem);
    MPM_SET_SMEM((k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER | G2P_HIST, 3>), G2PSmem);
    MPM_SET_SMEM((k_g2p_tile<G2P_GATHER | G2P_GRADW, 4>), G2PSmem);
#undef MPM_SET_SMEM
    return cudaSuccess;
}

template <int MODE>
cudaError_t launch_p2g_tile(Planes P, int* sorted_ids, const int4* pblock_list,
                            DevCounters* dc, float4* grid, GridDims gd, SimConst sc, float dt, int num_sms, int n_bound, cudaStream_t st,
                            const Planes* fupd_target = nullptr, bool fupd_fast = false, const PeerLayers* peer = nullptr) {
    (void)n_bound;
    const bool w3 = sc.pd.quadratic != 0;        // quadratic stencil: the fused substep's kernels have W = 3 instantiations
    cudaError_t e = cudaMemsetAsync(&dc->work_a, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const Planes Nx = fupd_target ? *fupd_target : P;
    const PeerLayers pl = peer ? *peer : PeerLayers{ nullptr, nullptr };
    const int fu = (fupd_target && MODE == P2G_FUSED) ? (fupd_fast ? 2 : 1) : 0;      // only the fused substep moves the F-update into P2G
    const bool pe = peer && MODE == P2G_FUSED;
    // stencil: the cubic (reference) path always runs W = 4 instantiations, which carry no trace of the switch; the quadratic
    // stencil runs W = 3 for the fused substep's kernels and the generic W = 0 (4-wide, run-time weights) elsewhere
#define MPM_P2G_LAUNCH_W(FU, PE, WW) k_p2g_tile<MODE, FU, PE, WW><<<num_sms * 2, P2G_T, sizeof(P2GSmem), st>>>(P, sorted_ids, pblock_list, dc, grid, gd, sc, dt, Nx, pl)
#define MPM_P2G_LAUNCH(FU, PE) do { if (w3) MPM_P2G_LAUNCH_W(FU, PE, 0); else MPM_P2G_LAUNCH_W(FU, PE, 4); } while (0)
    if constexpr (MODE == P2G_FUSED) {
        if (pe) { if (fu == 2) MPM_P2G_LAUNCH(2, true); else if (fu == 1) MPM_P2G_LAUNCH(1, true); else MPM_P2G_LAUNCH(0, true); }
        else if (w3 && fu == 2) MPM_P2G_LAUNCH_W(2, false, 3);
        else if (w3 && fu == 0) MPM_P2G_LAUNCH_W(0, false, 3);
        else { if (fu == 2) MPM_P2G_LAUNCH_W(2, false, 4); else if (fu == 1) MPM_P2G_LAUNCH(1, false); else MPM_P2G_LAUNCH_W(0, false, 4); }
    } else {
        MPM_P2G_LAUNCH(0, false);
    }
#undef MPM_P2G_LAUNCH_W
#undef MPM_P2G_LAUNCH
    return cudaGetLastError();
}

struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; int gather_ctas_per_sm; cudaEvent_t mid; bool mid_recorded; };
template <int FLAGS>
cudaError_t launch_g2p_tile(Planes C, Planes N, const int* sorted_ids, const int4* pblock_list,
                            DevCounters* dc, const float4* grid, GridDims gd, SimConst sc, float dt, int num_sms, int n_bound, cudaStream_t st,
                            SideStream* side, bool fupd_fast = false, int* key_out = nullptr, int* blk_count = nullptr, MigOut mo = MigOut{ nullptr, nullptr, 0 }) {
    cudaError_t e = cudaMemsetAsync(&dc->work_b, 0, sizeof(int), st);
    if (side) side->mid_recorded = false;
    if (e != cudaSuccess) return e;
    // The F-update (HBM-bound, 64 registers) and the gather (issue-bound) touch disjoint planes: when both run, the
    // F-update goes to a side stream so that its CTAs share the SMs with the gather's persistent CTAs.
    const bool overlap = side && side->stream && (FLAGS & G2P_F) && (FLAGS & G2P_GATHER);
    if (FLAGS & G2P_F) {
        cudaStream_t fs = st;
        if (overlap) {
            if ((e = cudaEventRecord(side->fork, st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return e;
            fs = side->stream;
        }
        if (fupd_fast) k_fupdate<(FLAGS & G2P_REORDER) != 0, true><<<n_bound > 0 ? (n_bound + 255) / 256 : 1, 256, 0, fs>>>(C, N, sorted_ids, dc, sc, dt);
        else k_fupdate<(FLAGS & G2P_REORDER) != 0><<<n_bound > 0 ? (n_bound + 255) / 256 : 1, 256, 0, fs>>>(C, N, sorted_ids, dc, sc, dt);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (!overlap && side && side->mid && (FLAGS & G2P_GATHER)) {      // per-kernel timing: F-update | gather
            if ((e = cudaEventRecord(side->mid, st)) != cudaSuccess) return e;
            side->mid_recorded = true;
        }
    }
    if (FLAGS & G2P_GATHER) {
        const int per_sm = overlap ? side->gather_ctas_per_sm : G2P_MIN_CTAS;
        constexpr int GF = FLAGS & ~G2P_F;
        constexpr bool CAN_HIST = (GF & G2P_REORDER) != 0 && (GF & G2P_ADVECT) != 0;       // the fused substep's gather
#define MPM_G2P_LAUNCH_W(F, WW) k_g2p_tile<F, WW><<<num_sms * per_sm, G2P_T, sizeof(G2PSmem), st>>>(C, N, sorted_ids, pblock_list, dc, grid, gd, sc, dt, key_out, blk_count, mo)
#define MPM_G2P_LAUNCH(F) do { if (sc.pd.quadratic) MPM_G2P_LAUNCH_W(F, 0); else MPM_G2P_LAUNCH_W(F, 4); } while (0)
        if constexpr (CAN_HIST) {
            if (key_out && sc.pd.quadratic) MPM_G2P_LAUNCH_W(GF | G2P_HIST, 3);       // quadratic stencil: W = 3 instantiation of the fused substep's gather
            else if (key_out) MPM_G2P_LAUNCH_W(GF | G2P_HIST, 4);
            else MPM_G2P_LAUNCH(GF);
        } else MPM_G2P_LAUNCH(GF);
#undef MPM_G2P_LAUNCH_W
#undef MPM_G2P_LAUNCH
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (overlap) {
        if ((e = cudaEventRecord(side->join, side->stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(st, side->join, 0)) != cudaSuccess) return e;
    }
    if (FLAGS & G2P_REORDER) {
        k_copy_parked<<<64, 256, 0, st>>>(C, N, sorted_ids, dc, gd, sc.pd, key_out, blk_count, mo);     // parked particles are few; 16 K slots per launch wave
        e = cudaGetLastError();
    }
    return e;
}

#endif  // host launch code

}  // namespace mpm

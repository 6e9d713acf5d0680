// Single-pass substep kernel ("G2P2G"): for every occupied particle block, ONE CTA
//   * fetches the block's 2x2x2 tile of grid(t) (post-update velocities) with TMA, one block ahead,
//   * per particle (2 per thread): gathers v and the APIC matrix at x_t from the tile, runs the F-update, advects, writes
//     the particle into the other buffer at its sorted rank, forms next substep's block key + histogram, and derives the
//     scatter record of the NEW position -- the stress comes straight from this thread's SVD factors,
//   * counting-sorts the chunk by (new) cell, accumulates in registers, folds and issues one vector red per tile node into
//     grid(t+1) -- the phases of k_p2g_tile.
// It replaces gather(t) + P2G(t+1) of the two-kernel substep (main.cpp:192-218: G2P and advect of one frame, P2G of the
// next; the grid update still runs between two launches), reads a particle once and writes it once per substep, and puts
// the particle loads next to the ~1400 instructions of gather + F-update that hide their latency.
// A particle whose new stencil base left its home block's 4^3 base nodes cannot join the register accumulation of a
// (cell, x-slab) thread. It moved less than one cell (|v| dt < h), so its stencil stays inside the home blocks dilated by
// one grid block, which is the active list k_mark_dilated builds: it scatters its 64 nodes directly with vector reds
// ("stray": a few per cent at 200 m/s, none in a resting slab). A particle that moved further ("far stray", only in an
// unstable run) is put on a list and scattered by k_g2p2g_far, which activates and clears the blocks it needs first.
#pragma once
#include "mpm_tile_kernels.cuh"

namespace mpm {

struct G2P2GSmem {
    P2GSmem p;                                     // records / fold buffer / cell counts / work item of the scatter half
    float4 tile[2][G2P_LIN_SLOTS];                 // grid(t) tile of the current and the next block (linear padded layout of the gather)
    float4 xs[P2G_CH];                             // pre-sort: position + mass of chunk slot q
    int gid_s[P2G_CH];                             // pre-sort: particle slot of chunk slot q
    unsigned short order_in[P2G_CH];               // chunk slots in (old) cell order
    int cell_cnt_in[64];
    unsigned long long bar[2];
};

// 128 row-wise bulk copies of one tile, issued by the 32 lanes of one warp (lane 0 arms the barrier first)
MPM_DI void g2p2g_fetch_tile(float4* tile, unsigned long long* bar, const float4* __restrict__ grid, const GridDims& gd, int bc, int lane) {
    const int pbk = bc & (PB_COORD_MAX - 1), pbj = (bc >> PB_COORD_BITS) & (PB_COORD_MAX - 1), pbi_l = bc >> (2 * PB_COORD_BITS);
    if (lane == 0) mbar_expect_tx(bar, 8 * 1024);
    __syncwarp();
    const int row = lane & 15, li = row >> 2, lj = row & 3;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int d = (lane >> 4) + 2 * m, di = d >> 2, dj = (d >> 1) & 1, dk = d & 1;
        const size_t gb = ((size_t)(pbi_l + di) * gd.nbj + pbj + dj) * gd.nbk + pbk + dk;
        tma_load_1d(&tile[(di * 4 + li) * G2P_LIN_PLANE + (dj * 4 + lj) * G2P_LIN_ROW + dk * 4], grid + gb * 64 + row * 4, 64, bar);
    }
}

// the 64 contributions of one particle as vector reds (strays)
MPM_DI void g2p2g_scatter_direct(float4* __restrict__ grid, const GridDims& gd, const SimConst& sc, int bx, int by, int bz,
                                 const float (&wx)[4], const float (&wy)[4], const float (&wz)[4], float m, float c0x, float c0y, float c0z, const float (&A)[9]) {
#pragma unroll 1
    for (int a = 0; a < 4; ++a)
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const float wab = wx[a] * wy[b];
            const float ha = (float)a * sc.h, hb = (float)b * sc.h;
            const float vx = c0x + A[0] * ha + A[1] * hb, vy = c0y + A[3] * ha + A[4] * hb, vz = c0z + A[6] * ha + A[7] * hb;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float w = wab * wz[c], hc = (float)c * sc.h;
                if (w == 0.0f) continue;
                atomicAdd(&grid[node_index(gd, bx + a, by + b, bz + c)], make_float4(w * m, w * (vx + A[2] * hc), w * (vy + A[5] * hc), w * (vz + A[8] * hc)));
            }
        }
}
// scatter coefficients of a particle from its (new) state: contribution to node x_i is w * (m, c0 + A (x_i - x_base))
MPM_DI void g2p2g_coeffs(const float (&x)[3], float m, const float (&v)[3], const float (&B)[9], const float (&tau)[6], const SimConst& sc, float dt,
                         int cx, int cy, int cz, float& c0x, float& c0y, float& c0z, float (&A)[9]) {
    const float md = m * sc.dinv;
    const float M[9] = { tau[0], tau[3], tau[4], tau[3], tau[1], tau[5], tau[4], tau[5], tau[2] };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = md * B[c * 3 + r] - dt * M[r * 3 + c];
    const float d0 = (float)(cx - 1) * sc.h - x[0], d1 = (float)(cy - 1) * sc.h - x[1], d2 = (float)(cz - 1) * sc.h - x[2];
    c0x = m * v[0] + A[0] * d0 + A[1] * d1 + A[2] * d2;
    c0y = m * v[1] + A[3] * d0 + A[4] * d1 + A[5] * d2;
    c0z = m * v[2] + A[6] * d0 + A[7] * d1 + A[8] * d2;
}

// occupied particle blocks dilated by one grid block per side -> active grid-block list of grid(t+1) (stamp -epoch: the
// 2x2x2 marks of k_scan_apply carry +epoch)
__global__ void k_mark_dilated(const int4* __restrict__ pblock_list, DevCounters* dc, GridDims gd, int* __restrict__ gflag, int* __restrict__ gblock_list) {
    const int n = dc->n_active_pblocks * 27, stamp = -dc->epoch;
    const int layers = gd.hi - gd.lo + 1;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int b = idx / 27, d = idx - b * 27;
        const int bc = pblock_list[b].w;
        const int gi = (bc >> (2 * PB_COORD_BITS)) + d / 9 - 1, gj = ((bc >> PB_COORD_BITS) & (PB_COORD_MAX - 1)) + (d / 3) % 3 - 1, gk = (bc & (PB_COORD_MAX - 1)) + d % 3 - 1;
        if (gi < 0 || gi >= layers || gj < 0 || gj >= gd.nbj || gk < 0 || gk >= gd.nbk) continue;
        const int gb = (gi * gd.nbj + gj) * gd.nbk + gk;
        if (atomicExch(&gflag[gb], stamp) != stamp) gblock_list[atomicAdd(&dc->n_active_gblocks, 1)] = gb;
    }
}

template <bool FAST /* tolerance-form F-update */>
__global__ void __launch_bounds__(P2G_T, 2)
k_g2p2g(Planes P, Planes N, const int* __restrict__ sorted_ids, const int4* __restrict__ pblock_list, DevCounters* dc,
        const float4* __restrict__ grid_prev, float4* __restrict__ grid_next, GridDims gd, SimConst sc, float dt,
        int* __restrict__ key_out, int* __restrict__ blk_count, int* __restrict__ far_list, int far_cap) {
    MPM_DYN_SMEM(g2p_smem_raw, 128);
    G2P2GSmem& G = *reinterpret_cast<G2P2GSmem*>(g2p_smem_raw);
    P2GSmem& S = G.p;
    const int t = threadIdx.x, lane = t & 31;
    const int my_cell = t >> 2, my_a = t & 3;
    const int my_cx = my_cell >> 4, my_cy = (my_cell >> 2) & 3, my_cz = my_cell & 3;
    const float fa = (float)my_a;
    const int n_work = dc->n_active_pblocks;
    int w_ticket = 0;
    int4 wk_reg = make_int4(-1, 0, 0, 0);
    const int4 no_work = make_int4(-1, 0, 0, 0);
    if (t == 0) {
        mbar_init(&G.bar[0], 1); mbar_init(&G.bar[1], 1);
        mbar_fence_init();
        const int w0 = atomicAdd(&dc->work_a, 1);
        S.work = w0 < n_work ? pblock_list[w0] : no_work;
        w_ticket = atomicAdd(&dc->work_a, 1);
    }
    __syncthreads();
    if (t < 32 && S.work.x >= 0) g2p2g_fetch_tile(G.tile[0], &G.bar[0], grid_prev, gd, S.work.w, lane);
    int gid_pref[P2G_PPT];
    float4 xm_pref[P2G_PPT];
    {
        int nch0;
        p2g_first_chunk_ids(S.work, sorted_ids, t, gid_pref, nch0);
#pragma unroll
        for (int u = 0; u < P2G_PPT; ++u) xm_pref[u] = (t + u * P2G_T < nch0) ? P.p[0][gid_pref[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned phase_bits = 0;                       // bit b = parity the next wait on bar[b] expects
    for (int kb = 0;; ++kb) {
        const int4 wk = S.work;
        if (t < 64) { S.cell_cnt[t] = 0; G.cell_cnt_in[t] = 0; }
        __syncthreads();          // everyone has read S.work; the fold of the previous block is done with t1
        if (wk.x < 0) break;
        const int start = wk.y, cnt = wk.z;
        const int pbk = wk.w & (PB_COORD_MAX - 1), pbj = (wk.w >> PB_COORD_BITS) & (PB_COORD_MAX - 1), pbi = (wk.w >> (2 * PB_COORD_BITS)) + gd.lo;
        const float4* __restrict__ tile = G.tile[kb & 1];
        bool tile_ready = false;
        f32x2_t AM[8], AX[8], AY[8], AZ[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) AM[i] = AX[i] = AY[i] = AZ[i] = 0ull;
        const int n_chunks = (cnt + P2G_CH - 1) / P2G_CH;
        for (int ck = 0; ck < n_chunks; ++ck) {
            const int nch = n_chunks == 1 ? cnt : (cnt - ck + n_chunks - 1) / n_chunks;
            if (ck > 0) {
                if (t < 64) { S.cell_cnt[t] = 0; G.cell_cnt_in[t] = 0; }
                __syncthreads();
            }
            // ---- pre-sort of the chunk by the cell of x_t: the lanes of a warp then gather from few cells (shared-memory
            // broadcasts instead of bank conflicts) and the re-sorted buffer is written in cell order ----
            int cr_in[P2G_PPT];
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u) {
                const int q = t + u * P2G_T;
                cr_in[u] = -1;
                if (q < nch) {
                    const int gid = ck == 0 ? gid_pref[u] : sorted_ids[start + ck + q * n_chunks];
                    const float4 xm = ck == 0 ? xm_pref[u] : P.p[0][gid];
                    const int lc = (((cell_of_t<0>(xm.x, sc.pd) - 1) - 4 * pbi) * 4 + ((cell_of_t<0>(xm.y, sc.pd) - 1) - 4 * pbj)) * 4 + ((cell_of_t<0>(xm.z, sc.pd) - 1) - 4 * pbk);
                    G.xs[q] = xm; G.gid_s[q] = gid;
                    cr_in[u] = (lc & 63) | (atomicAdd(&G.cell_cnt_in[lc & 63], 1) << 8);
                }
            }
            __syncthreads();
            {
                const int c0 = G.cell_cnt_in[2 * lane], c1 = G.cell_cnt_in[2 * lane + 1];
                int inc = c0 + c1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                const int ex = inc - (c0 + c1);
#pragma unroll
                for (int u = 0; u < P2G_PPT; ++u) {
                    const int c = cr_in[u] < 0 ? 0 : (cr_in[u] & 63);
                    const int e = __shfl_sync(0xffffffffu, ex, c >> 1), f = __shfl_sync(0xffffffffu, c0, c >> 1);
                    if (cr_in[u] >= 0) G.order_in[e + ((c & 1) ? f : 0) + (cr_in[u] >> 8)] = (unsigned short)(t + u * P2G_T);
                }
            }
            __syncthreads();
            if (!tile_ready) {
                mbar_wait(&G.bar[kb & 1], (phase_bits >> (kb & 1)) & 1u);
                phase_bits ^= 1u << (kb & 1);
                tile_ready = true;
            }
            // ---- particle phase: gather(t), F-update(t), advect, store, key; scatter record of x_{t+1} ----
            int cell_rank[P2G_PPT];
#pragma unroll 1
            for (int u = 0; u < P2G_PPT; ++u) {
                const int q = t + u * P2G_T;                    // slot in cell order
                cell_rank[u] = -1;
                int key_new = KEY_DEAD;
                if (q < nch) {
                    const int qi = G.order_in[q];
                    const int j = start + ck + q * n_chunks;    // rank the particle is stored at
                    const int gid = G.gid_s[qi];
                    const float4 a0 = G.xs[qi];
                    const FUpdIn fin = fupd_load(P, gid);
                    float x[3] = { a0.x, a0.y, a0.z };
                    const float m = a0.w;
                    float vn[3], Bn[9];
                    {   // separable gather on packed pairs (k_g2p_tile)
                        float wx[4], wy[4], wz[4];
                        const int cx = cell_and_weights_t<0>(x[0], sc.pd, wx), cy = cell_and_weights_t<0>(x[1], sc.pd, wy), cz = cell_and_weights_t<0>(x[2], sc.pd, wz);
                        const int ox = (cx - 1) - 4 * pbi, oy = (cy - 1) - 4 * pbj, oz = (cz - 1) - 4 * pbk;
                        const float4* __restrict__ lin = tile + (ox * G2P_LIN_PLANE + oy * G2P_LIN_ROW + oz);
                        f32x2_t WX[4], WY[4], WZ[4], V[3] = { 0ull, 0ull, 0ull };
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            WX[a] = pack2(wx[a], wx[a] * ((float)(cx - 1 + a) * sc.h - x[0]));
                            WY[a] = pack2(wy[a], wy[a] * ((float)(cy - 1 + a) * sc.h - x[1]));
                            WZ[a] = pack2(wz[a], wz[a] * ((float)(cz - 1 + a) * sc.h - x[2]));
                        }
                        float By[3] = { 0, 0, 0 }, Bz[3] = { 0, 0, 0 };
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            f32x2_t T[3] = { 0ull, 0ull, 0ull };
                            float t1z[3] = { 0, 0, 0 };
#pragma unroll
                            for (int bb = 0; bb < 4; ++bb) {
                                f32x2_t Sx[3] = { 0ull, 0ull, 0ull };
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) {
                                    const float4 n = lin[a * G2P_LIN_PLANE + bb * G2P_LIN_ROW + cc];
                                    Sx[0] = ffma2(WZ[cc], pack2(n.y, n.y), Sx[0]);
                                    Sx[1] = ffma2(WZ[cc], pack2(n.z, n.z), Sx[1]);
                                    Sx[2] = ffma2(WZ[cc], pack2(n.w, n.w), Sx[2]);
                                }
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const float s0q = lo2(Sx[c]);
                                    T[c] = ffma2(WY[bb], pack2(s0q, s0q), T[c]);
                                    t1z[c] += lo2(WY[bb]) * hi2(Sx[c]);
                                }
                            }
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float t0q = lo2(T[c]);
                                V[c] = ffma2(WX[a], pack2(t0q, t0q), V[c]);
                                By[c] += lo2(WX[a]) * hi2(T[c]); Bz[c] += lo2(WX[a]) * t1z[c];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c) { vn[c] = lo2(V[c]); Bn[c] = hi2(V[c]); Bn[3 + c] = By[c]; Bn[6 + c] = Bz[c]; }
                    }
                    // F-update (cpp:306-330) with the B of the previous gather, and the stress of the new FE
                    float tau[6];
                    {
                        float Bo[9] = { fin.a1.x, fin.a1.y, fin.a1.z, fin.a1.w, fin.a2.x, fin.a2.y, fin.a2.z, fin.a2.w, fin.a3.x };
                        float FE[9] = { fin.a6.z, fin.a6.w, fin.a7.x, fin.a7.y, fin.a7.z, fin.a7.w, fin.a8.x, fin.a8.y, fin.a8.z };
                        float FP[9] = { fin.a8.w, fin.a9.x, fin.a9.y, fin.a9.z, fin.a9.w, fin.a10.x, fin.a10.y, fin.a10.z, fin.a10.w };
                        float Ug[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, Sg[3] = { 1, 1, 1 };
                        if (FAST) {
                            if (!f_update_fast(Bo, FE, FP, sc.dinv * dt, sc.clamp_lo, sc.clamp_hi, Ug, Sg)) dc->svd_failed = 1;
                            tau_from_factors_fast(Ug, Sg, m3_det_fast(FE), m3_det_fast(FP), fin.a6.x, sc.dinv, sc.mu0, sc.lambda0, sc.xi, tau);
                        } else {
                            if (!f_update_rn(Bo, FE, FP, sc.dinv, dt, sc.clamp_lo, sc.clamp_hi, Ug, Sg)) dc->svd_failed = 1;
                            tau_from_factors(Ug, Sg, m3_det_rn(FE), m3_det_rn(FP), fin.a6.x, sc.dinv, sc.E, sc.nu, sc.xi, tau);
                        }
                        N.p[6][j] = make_float4(fin.a6.x, fin.a6.y, FE[0], FE[1]);
                        N.p[7][j] = make_float4(FE[2], FE[3], FE[4], FE[5]);
                        N.p[8][j] = make_float4(FE[6], FE[7], FE[8], FP[0]);
                        N.p[9][j] = make_float4(FP[1], FP[2], FP[3], FP[4]);
                        N.p[10][j] = make_float4(FP[5], FP[6], FP[7], FP[8]);
                    }
                    // advect (cpp:344-350, 381-388) and store the particle at its sorted rank in the other buffer
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        float xa = add_rn(x[a], mul_rn(vn[a], dt));
                        if (xa < sc.pos_lo) xa = sc.pos_lo;
                        if (sc.pos_hi[a] < xa) xa = sc.pos_hi[a];
                        x[a] = xa;
                    }
                    N.p[0][j] = make_float4(x[0], x[1], x[2], m);
                    N.p[1][j] = make_float4(Bn[0], Bn[1], Bn[2], Bn[3]);
                    N.p[2][j] = make_float4(Bn[4], Bn[5], Bn[6], Bn[7]);
                    N.p[3][j] = make_float4(Bn[8], vn[0], vn[1], vn[2]);
                    N.p[4][j] = make_float4(tau[0], tau[1], tau[2], tau[3]);
                    N.p[5][j] = make_float4(tau[4], tau[5], 0.0f, 0.0f);
                    // next substep's block key (k_bin_count's job) and the scatter record of the new position
                    float wx[4], wy[4], wz[4];
                    const int cx = cell_and_weights_t<0>(x[0], sc.pd, wx), cy = cell_and_weights_t<0>(x[1], sc.pd, wy), cz = cell_and_weights_t<0>(x[2], sc.pd, wz);
                    const bool in_grid = cx >= 2 && cy >= 2 && cz >= 2 && cx + 2 <= gd.I - 1 && cy + 2 <= gd.J - 1 && cz + 2 <= gd.K - 1;
                    const int nbi = (cx - 1) >> 2, nbj_ = (cy - 1) >> 2, nbk_ = (cz - 1) >> 2;
                    key_new = m < 0.0f ? KEY_DEAD : (!in_grid ? gd.n_pblocks : (nbi < gd.lo ? gd.n_pblocks + 1 : (nbi >= gd.hi ? gd.n_pblocks + 2 : ((nbi - gd.lo) * gd.npbj + nbj_) * gd.npbk + nbk_)));
                    key_out[j] = key_new;
                    if (key_new >= 0 && key_new < gd.n_pblocks) {
                        float c0x, c0y, c0z, A[9];
                        g2p2g_coeffs(x, m, vn, Bn, tau, sc, dt, cx, cy, cz, c0x, c0y, c0z, A);
                        const int lx = (cx - 1) - 4 * pbi, ly = (cy - 1) - 4 * pbj, lz = (cz - 1) - 4 * pbk;
                        if ((unsigned)lx < 4u && (unsigned)ly < 4u && (unsigned)lz < 4u) {
                            S.u.c.wx[q] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                            S.u.c.wy[q] = make_float4(wy[0], wy[1], wy[2], wy[3]);
                            S.u.c.wz[q] = make_float4(wz[0], wz[1], wz[2], wz[3]);
                            S.u.c.qc[q] = make_float4(m, c0x, c0y, c0z);
                            S.u.c.hA0[q] = make_float4(A[0] * sc.h, A[1] * sc.h, A[2] * sc.h, A[3] * sc.h);
                            S.u.c.hA1[q] = make_float4(A[4] * sc.h, A[5] * sc.h, A[6] * sc.h, A[7] * sc.h);
                            S.u.c.hA8[q] = A[8] * sc.h;
                            const int lc = (lx * 4 + ly) * 4 + lz;
                            cell_rank[u] = lc | (atomicAdd(&S.cell_cnt[lc], 1) << 8);
                        } else if (lx >= -1 && lx <= 4 && ly >= -1 && ly <= 4 && lz >= -1 && lz <= 4) {
                            g2p2g_scatter_direct(grid_next, gd, sc, cx - 1, cy - 1, cz - 1, wx, wy, wz, m, c0x, c0y, c0z, A);      // stray
                        } else {
                            const int f = atomicAdd(&dc->n_far, 1);                                                              // far stray
                            if (f < far_cap) far_list[f] = j; else dc->far_overflow = 1;
                        }
                    }
                }
                // warp-aggregated histogram of the new keys (every lane takes part)
                const unsigned peers = __match_any_sync(0xffffffffu, key_new);
                if (key_new >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&blk_count[key_new], __popc(peers));
            }
            __syncthreads();
            // ---- counting sort of the chunk by (new) cell: every warp scans the 64 counts in its own registers ----
            const int c0 = S.cell_cnt[2 * lane], c1 = S.cell_cnt[2 * lane + 1];
            int inc = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            const int ex = inc - (c0 + c1);
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u) {
                const int c = cell_rank[u] < 0 ? 0 : (cell_rank[u] & 255);
                const int e = __shfl_sync(0xffffffffu, ex, c >> 1), f = __shfl_sync(0xffffffffu, c0, c >> 1);
                if (cell_rank[u] >= 0) S.u.c.order[e + ((c & 1) ? f : 0) + (cell_rank[u] >> 8)] = (unsigned short)(t + u * P2G_T);
            }
            int i0, i1;
            {
                const int e = __shfl_sync(0xffffffffu, ex, my_cell >> 1), f = __shfl_sync(0xffffffffu, c0, my_cell >> 1), g = __shfl_sync(0xffffffffu, c1, my_cell >> 1);
                i0 = e + ((my_cell & 1) ? f : 0);
                i1 = i0 + ((my_cell & 1) ? g : f);
            }
            __syncthreads();
            // ---- register accumulation over the particles of my cell (k_p2g_tile, packed pairs) ----
            int i = i0;
            if (sc.p2g_rotate && i1 - i0 > 1) i += (my_cell & 7) % (i1 - i0);
            int pi = i1 > i0 ? S.u.c.order[i] : 0;
#pragma unroll 1
            for (int k = 0; k < i1 - i0; ++k) {
                i = (i + 1 == i1) ? i0 : i + 1;
                const float wxa = reinterpret_cast<const float*>(&S.u.c.wx[pi])[my_a];
                const float4 wy = S.u.c.wy[pi], wz = S.u.c.wz[pi], qc = S.u.c.qc[pi], h0 = S.u.c.hA0[pi], h1 = S.u.c.hA1[pi];
                const float h8 = S.u.c.hA8[pi];
                pi = S.u.c.order[i];
                const float bx = qc.y + fa * h0.x, by = qc.z + fa * h0.w, bz = qc.w + fa * h1.z;
                const float wyv[4] = { wy.x, wy.y, wy.z, wy.w };
                const f32x2_t wzp[2] = { pack2(wz.x, wz.y), pack2(wz.z, wz.w) };
                const f32x2_t cp2[2] = { pack2(0.f, 1.f), pack2(2.f, 3.f) };
                const f32x2_t mm = pack2(qc.x, qc.x), sx = pack2(h0.z, h0.z), sy = pack2(h1.y, h1.y), sz = pack2(h8, h8);
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
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
            __syncthreads();
        }
        float4 acc[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[2 * i] = make_float4(lo2(AM[i]), lo2(AX[i]), lo2(AY[i]), lo2(AZ[i]));
            acc[2 * i + 1] = make_float4(hi2(AM[i]), hi2(AX[i]), hi2(AY[i]), hi2(AZ[i]));
        }
        if (t == 0) wk_reg = w_ticket < n_work ? pblock_list[w_ticket] : no_work;
        // ---- z-fold with warp shuffles ----
        float4 s0[4], s1[4];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) { s0[bb] = acc[bb * 4]; s1[bb] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) {
            const int src = (lane & ~12) | (((my_cz - cc) & 3) << 2);
            const bool lo = cc <= my_cz;
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
            S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz] = s0[bb];
            if (my_cz < 3) S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz + 4] = s1[bb];
        }
        if (t == 0) { S.work = wk_reg; w_ticket = atomicAdd(&dc->work_a, 1); }
        __syncthreads();
        // the next block: its tile (into the buffer last read one block ago, behind several barriers) and the ids of its first chunk
        if (t < 32 && S.work.x >= 0) g2p2g_fetch_tile(G.tile[(kb + 1) & 1], &G.bar[(kb + 1) & 1], grid_prev, gd, S.work.w, lane);
        int nchn;
        p2g_first_chunk_ids(S.work, sorted_ids, t, gid_pref, nchn);
        // ---- x/y fold, one vector red per tile node into grid(t+1) ----
        for (int n = t; n < 343; n += P2G_T) {
            const int ni = n / 49, nj = (n / 7) % 7, nk = n % 7;
            float4 v[16];
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
#pragma unroll
                for (int dy = 0; dy < 4; ++dy) {
                    const int cx = ni - dx, cy = nj - dy;
                    const bool ok = cx >= 0 && cx <= 3 && cy >= 0 && cy <= 3;
                    v[dx * 4 + dy] = ok ? S.u.t1[cx][cy][dx][dy * 7 + nk] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) { sum.x += v[i].x; sum.y += v[i].y; sum.z += v[i].z; sum.w += v[i].w; }
            if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f)
                atomicAdd(&grid_next[node_index(gd, 4 * pbi + ni, 4 * pbj + nj, 4 * pbk + nk)], sum);
        }
        // positions of the next block's first chunk (their ids have arrived during the fold)
#pragma unroll
        for (int u = 0; u < P2G_PPT; ++u) xm_pref[u] = (t + u * P2G_T < nchn) ? P.p[0][gid_pref[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Far strays (moved more than one cell in one substep: an unstable run, but it must stay correct): ONE CTA activates and
// clears the grid blocks they touch that are not active yet, then scatters them from their stored state.
__global__ void __launch_bounds__(256)
k_g2p2g_far(Planes N, DevCounters* dc, const int* __restrict__ far_list, int far_cap, float4* __restrict__ grid_next, GridDims gd, SimConst sc, float dt,
            int* __restrict__ gflag, int* __restrict__ gblock_list) {
    const int n = min(dc->n_far, far_cap);
    if (n == 0) return;
    const int stamp = -dc->epoch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 xm = N.p[0][far_list[i]];
        const int b[3] = { cell_of_t<0>(xm.x, sc.pd) - 1, cell_of_t<0>(xm.y, sc.pd) - 1, cell_of_t<0>(xm.z, sc.pd) - 1 };
        for (int d = 0; d < 8; ++d) {
            const int gi = ((b[0] + ((d & 4) ? 3 : 0)) >> 2) - gd.lo, gj = (b[1] + ((d & 2) ? 3 : 0)) >> 2, gk = (b[2] + ((d & 1) ? 3 : 0)) >> 2;
            const int gb = (gi * gd.nbj + gj) * gd.nbk + gk;
            if (atomicExch(&gflag[gb], stamp) != stamp) {
                gblock_list[atomicAdd(&dc->n_active_gblocks, 1)] = gb;
                for (int k = 0; k < 64; ++k) grid_next[(size_t)gb * 64 + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) dc->n_far = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = far_list[i];
        const float4 xm = N.p[0][j], b0 = N.p[1][j], b1 = N.p[2][j], b2 = N.p[3][j], t0 = N.p[4][j], t1 = N.p[5][j];
        const float x[3] = { xm.x, xm.y, xm.z }, v[3] = { b2.y, b2.z, b2.w }, B[9] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x };
        const float tau[6] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y };
        float wx[4], wy[4], wz[4], c0x, c0y, c0z, A[9];
        const int cx = cell_and_weights_t<0>(x[0], sc.pd, wx), cy = cell_and_weights_t<0>(x[1], sc.pd, wy), cz = cell_and_weights_t<0>(x[2], sc.pd, wz);
        g2p2g_coeffs(x, xm.w, v, B, tau, sc, dt, cx, cy, cz, c0x, c0y, c0z, A);
        g2p2g_scatter_direct(grid_next, gd, sc, cx - 1, cy - 1, cz - 1, wx, wy, wz, xm.w, c0x, c0y, c0z, A);
    }
}

#if !defined(MPM_HOST_EMU) || defined(MPM_HOST_EMU_API)
inline cudaError_t g2p2g_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_g2p2g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(G2P2GSmem))) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_g2p2g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(G2P2GSmem));
}
#endif

}  // namespace mpm

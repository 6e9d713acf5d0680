// Implicit (optimisation-based) grid velocity update: the objective of LagrangeEulerView::Energy / ElasticPotential /
// ElasticPlasticEnergyDensity (material_point_method.cpp:160-209) and its analytic gradient on the device, and the
// vector kernels of the L-BFGS driver in mpm_api.cu (the optimiser mathy.hpp:10-38 configures: external/mcloptlib
// LBFGS.hpp + Backtracking.hpp). The reference never calls this path (README.md:17 lists it as a TODO) and differentiates
// its float objective by central differences of 2.2e-6, i.e. searches along rounding noise; the device path evaluates
//     E(v) = sum_i 1/2 m_i |v_i - v*_i|^2 + sum_p V_p psi((I + dt sum_i v_i grad w_ip^T) FE_p, FP_p)
//     psi(F, FP) = mu |F - R|_F^2 + lambda/2 (det F - 1)^2,   mu = mu0 e, lambda = lambda0 e, e = exp(xi*1 - det FP) (as written)
//     dE/dv_i = m_i (v_i - v*_i) + dt sum_p V_p (2 mu (F - R) + lambda (J - 1) J F^-T) FE_p^T grad w_ip
// with grad w from hpp:59-71 (cubic B-spline and its derivative, hpp:20-52). Unknowns live in grid layout (float4 per
// node, .x unused) over the active grid blocks of the last binning; nodes without mass get a zero gradient and never move,
// which is the reference's "used_cells" (cpp:105-110).
// Precision: per-particle energies and the polar factor in double (the B200 has the fp64 rate to spare here), sums in
// double atomics, the gradient scatter in float vector reds.
#pragma once
#include "mpm_tile_kernels.cuh"

namespace mpm {

struct ImplicitConst { float mu0, lambda0, xi; int hardening; };

// rotation factor of the polar decomposition by Newton's iteration R <- (R + R^-T)/2; false if singular. The iteration is
// self-correcting and quadratically convergent, so it runs in float down to ~1e-6 and only the last steps in double (from an
// error e one double step leaves e^2 / 2): the fp64 work of the stress kernel drops from ~7 steps to 2.
MPM_DI bool polar_rotation_d(const double (&F)[9], double (&R)[9]) {
    {
        float X[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) X[i] = (float)F[i];
        for (int it = 0; it < 40; ++it) {
            const float a = X[0], b = X[3], c = X[6], d = X[1], e = X[4], f = X[7], g = X[2], hh = X[5], k = X[8];
            const float c00 = e * k - f * hh, c01 = -(d * k - f * g), c02 = d * hh - e * g;
            const float c10 = -(b * k - c * hh), c11 = a * k - c * g, c12 = -(a * hh - b * g);
            const float c20 = b * f - c * e, c21 = -(a * f - c * d), c22 = a * e - b * d;
            const float det = a * c00 + b * c01 + c * c02;
            if (det == 0.0f || !isfinite(det)) return false;
            const float id = 1.0f / det;
            const float cof[9] = { c00, c10, c20, c01, c11, c21, c02, c12, c22 };
            float diff = 0.0f;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float v = 0.5f * (X[i] + cof[i] * id);
                diff = fmaxf(diff, fabsf(v - X[i]));
                X[i] = v;
            }
            if (diff < 2e-6f) break;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (double)X[i];
    }
    for (int it = 0; it < 60; ++it) {
        // column-major: element (r, c) = R[c*3 + r]
        const double a = R[0], b = R[3], c = R[6], d = R[1], e = R[4], f = R[7], g = R[2], hh = R[5], k = R[8];
        const double c00 = e * k - f * hh, c01 = -(d * k - f * g), c02 = d * hh - e * g;
        const double c10 = -(b * k - c * hh), c11 = a * k - c * g, c12 = -(a * hh - b * g);
        const double c20 = b * f - c * e, c21 = -(a * f - c * d), c22 = a * e - b * d;
        const double det = a * c00 + b * c01 + c * c02;
        if (det == 0.0 || !isfinite(det)) return false;
        const double id = 1.0 / det;
        const double cof[9] = { c00, c10, c20, c01, c11, c21, c02, c12, c22 };      // cof[c*3 + r] = cofactor(r, c)
        double diff = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double v = 0.5 * (R[i] + cof[i] * id);
            diff = fmax(diff, fabs(v - R[i]));
            R[i] = v;
        }
        if (diff < 1e-13) break;      // the step that produced it squared the error once more
    }
    return true;
}

// node term: 1/2 m |v - v*|^2 summed into acc[0]; gradient m (v - v*) written (not added) when G != nullptr
__global__ void __launch_bounds__(256)
k_imp_nodes(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, const float4* __restrict__ grid,
            const float4* __restrict__ X, float4* __restrict__ G, double* __restrict__ acc) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    double e = 0.0;
    for (int b = blockIdx.x * per + sub; b < nb; b += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[b] * 64 + t;
        const float4 n = grid[idx];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n.x != 0.0f) {
            const float4 x = X[idx];
            const float d0 = x.y - n.y, d1 = x.z - n.z, d2 = x.w - n.w;
            e += 0.5 * (double)n.x * ((double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2);
            g = make_float4(0.f, n.x * d0, n.x * d1, n.x * d2);
        }
        if (G) G[idx] = g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(&acc[0], e);
}

// particle term: V_p psi summed into acc[1]; with GRAD the 64 gradient contributions of the particle as vector reds
template <bool GRAD>
__global__ void __launch_bounds__(128)
k_imp_particles(Planes P, const int* __restrict__ sorted_ids, DevCounters* dc, const float4* __restrict__ X, float4* __restrict__ G,
                GridDims gd, SimConst sc, float dt, ImplicitConst ic, double* __restrict__ acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (j < dc->n_binned) {
        const int p = sorted_ids[j];
        const float4 xm = P.p[0][p], a6 = P.p[6][p], a7 = P.p[7][p], a8 = P.p[8][p], a9 = P.p[9][p], a10 = P.p[10][p];
        const float FE[9] = { a6.z, a6.w, a7.x, a7.y, a7.z, a7.w, a8.x, a8.y, a8.z };
        const float FP[9] = { a8.w, a9.x, a9.y, a9.z, a9.w, a10.x, a10.y, a10.z, a10.w };
        const float V0 = a6.x;
        const int cx = cell_of_t<0>(xm.x, sc.pd), cy = cell_of_t<0>(xm.y, sc.pd), cz = cell_of_t<0>(xm.z, sc.pd);
        float wx[4], wy[4], wz[4], dx[4], dy[4], dz[4];
        axis_weights_and_derivatives(xm.x, sc.pd, cx, wx, dx);
        axis_weights_and_derivatives(xm.y, sc.pd, cy, wy, dy);
        axis_weights_and_derivatives(xm.z, sc.pd, cz, wz, dz);
        const float ih = 1.0f / sc.h;
        // A = I + dt sum_i v_i (grad w_i)^T, column-major A[c*3 + r] += dt v_r g_c
        double A[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 v = X[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)];
                    const float g0 = ih * dx[a] * wy[b] * wz[c], g1 = ih * wx[a] * dy[b] * wz[c], g2 = ih * wx[a] * wy[b] * dz[c];
                    const float vt[3] = { v.y * dt, v.z * dt, v.w * dt };
#pragma unroll
                    for (int r = 0; r < 3; ++r) { A[0 + r] += (double)(vt[r] * g0); A[3 + r] += (double)(vt[r] * g1); A[6 + r] += (double)(vt[r] * g2); }
                }
        double F[9], R[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) F[c * 3 + r] = A[0 + r] * FE[c * 3 + 0] + A[3 + r] * FE[c * 3 + 1] + A[6 + r] * FE[c * 3 + 2];
        const double detFP = (double)m3_det_fast(FP);
        const double hard = exp(ic.hardening == 0 ? (double)ic.xi - detFP : (double)ic.xi * (1.0 - detFP));
        const double mu = ic.mu0 * hard, lambda = ic.lambda0 * hard;
        if (!polar_rotation_d(F, R)) dc->svd_failed = 1;
        double fn2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double d = F[i] - R[i]; fn2 += d * d; }
        const double c00 = F[4] * F[8] - F[7] * F[5], c01 = F[7] * F[2] - F[1] * F[8], c02 = F[1] * F[5] - F[4] * F[2];      // cofactors (0,0), (0,1), (0,2)
        const double J = F[0] * c00 + F[3] * c01 + F[6] * c02;
        e = (double)V0 * (mu * fn2 + 0.5 * lambda * (J - 1.0) * (J - 1.0));
        if (GRAD) {
            // J F^-T = cofactor matrix: cof(r, c) stored column-major
            const double cof[9] = { c00, -(F[3] * F[8] - F[6] * F[5]), F[3] * F[7] - F[6] * F[4],
                                    c01, F[0] * F[8] - F[6] * F[2], -(F[0] * F[7] - F[6] * F[1]),
                                    c02, -(F[0] * F[5] - F[3] * F[2]), F[0] * F[4] - F[3] * F[1] };      // cof[c*3 + r] = cofactor(r, c)
            double Pk[9], Gm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Pk[i] = 2.0 * mu * (F[i] - R[i]) + lambda * (J - 1.0) * cof[i];
            // Gm = V0 Pk FE^T: Gm(r, c) = sum_k Pk(r, k) FE(c, k)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) Gm[c * 3 + r] = (double)V0 * (Pk[0 + r] * FE[0 + c] + Pk[3 + r] * FE[3 + c] + Pk[6 + r] * FE[6 + c]);
            const float Gf[9] = { (float)(Gm[0] * dt), (float)(Gm[1] * dt), (float)(Gm[2] * dt), (float)(Gm[3] * dt), (float)(Gm[4] * dt),
                                  (float)(Gm[5] * dt), (float)(Gm[6] * dt), (float)(Gm[7] * dt), (float)(Gm[8] * dt) };
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float g0 = ih * dx[a] * wy[b] * wz[c], g1 = ih * wx[a] * dy[b] * wz[c], g2 = ih * wx[a] * wy[b] * dz[c];
                        if (g0 == 0.0f && g1 == 0.0f && g2 == 0.0f) continue;
                        const float4 val = make_float4(0.f, Gf[0] * g0 + Gf[3] * g1 + Gf[6] * g2, Gf[1] * g0 + Gf[4] * g1 + Gf[7] * g2, Gf[2] * g0 + Gf[5] * g1 + Gf[8] * g2);
                        atomicAdd(&G[node_index(gd, cx - 1 + a, cy - 1 + b, cz - 1 + c)], val);
                    }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(&acc[1], e);
}

// ---- block-tile form of the particle term (the default; k_imp_particles above is the baseline, p2g_variant / g2p_variant 1) ----
// 1. k_g2p_tile<G2P_GATHER | G2P_GRADW> gathers A = I + dt grad v of the trial field from TMA-staged tiles into aux[3 j .. 3 j + 2]
// 2. k_imp_stress: thread per sorted rank j: F = A FE, polar factor and energy in double; with GRAD the 3x3
//    Gm = dt V0 (2 mu (F - R) + lambda (J - 1) J F^-T) FE^T overwrites aux[3 j ..]
// 3. k_imp_scatter_tile: grad_i += Gm grad w_ip, accumulated per (cell, x-slab) thread in registers and folded like P2G
template <bool GRAD>
__global__ void __launch_bounds__(128)
k_imp_stress(Planes P, const int* __restrict__ sorted_ids, DevCounters* dc, float4* __restrict__ aux, float dt, ImplicitConst ic, double* __restrict__ acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (j < dc->n_binned) {
        const int p = sorted_ids[j];
        const float4 a6 = P.p[6][p], a7 = P.p[7][p], a8 = P.p[8][p], a9 = P.p[9][p], a10 = P.p[10][p];
        const float4 q0 = aux[3 * (size_t)j], q1 = aux[3 * (size_t)j + 1], q2 = aux[3 * (size_t)j + 2];
        const float FE[9] = { a6.z, a6.w, a7.x, a7.y, a7.z, a7.w, a8.x, a8.y, a8.z };
        const float FP[9] = { a8.w, a9.x, a9.y, a9.z, a9.w, a10.x, a10.y, a10.z, a10.w };
        const double A[9] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x };
        const double V0 = a6.x;
        double F[9], R[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) F[c * 3 + r] = A[0 + r] * FE[c * 3 + 0] + A[3 + r] * FE[c * 3 + 1] + A[6 + r] * FE[c * 3 + 2];
        const double detFP = (double)m3_det_fast(FP);
        const double hard = exp(ic.hardening == 0 ? (double)ic.xi - detFP : (double)ic.xi * (1.0 - detFP));
        const double mu = ic.mu0 * hard, lambda = ic.lambda0 * hard;
        if (!polar_rotation_d(F, R)) dc->svd_failed = 1;
        double fn2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double d = F[i] - R[i]; fn2 += d * d; }
        const double c00 = F[4] * F[8] - F[7] * F[5], c01 = F[7] * F[2] - F[1] * F[8], c02 = F[1] * F[5] - F[4] * F[2];
        const double J = F[0] * c00 + F[3] * c01 + F[6] * c02;
        e = V0 * (mu * fn2 + 0.5 * lambda * (J - 1.0) * (J - 1.0));
        if (GRAD) {
            const double cof[9] = { c00, -(F[3] * F[8] - F[6] * F[5]), F[3] * F[7] - F[6] * F[4],
                                    c01, F[0] * F[8] - F[6] * F[2], -(F[0] * F[7] - F[6] * F[1]),
                                    c02, -(F[0] * F[5] - F[3] * F[2]), F[0] * F[4] - F[3] * F[1] };
            double Pk[9];
            float Gm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Pk[i] = 2.0 * mu * (F[i] - R[i]) + lambda * (J - 1.0) * cof[i];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) Gm[c * 3 + r] = (float)((double)dt * V0 * (Pk[0 + r] * FE[0 + c] + Pk[3 + r] * FE[3 + c] + Pk[6 + r] * FE[6 + c]));
            aux[3 * (size_t)j] = make_float4(Gm[0], Gm[1], Gm[2], Gm[3]);
            aux[3 * (size_t)j + 1] = make_float4(Gm[4], Gm[5], Gm[6], Gm[7]);
            aux[3 * (size_t)j + 2] = make_float4(Gm[8], 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(&acc[1], e);
}

struct ImpScatterSmem {
    union {
        struct {
            float4 wx[P2G_CH], wy[P2G_CH], wz[P2G_CH];      // weights of the four stencil nodes per axis
            float4 dx[P2G_CH], dy[P2G_CH], dz[P2G_CH];      // their derivatives / h
            float4 g0[P2G_CH], g1[P2G_CH];                  // Gm entries 0..3, 4..7 (column-major)
            float g8[P2G_CH];
            unsigned short order[P2G_CH];
        } c;
        float4 t1[4][4][4][42];                             // z-folded partial sums, as in P2GSmem
    } u;
    int cell_cnt[64];
    int4 work;
};
// Thread t = cell*4 + a as in k_p2g_tile; three channels (no mass), value at node (a, b, c) of the particle's stencil:
//   grad_r += G_r0 dwx_a wy_b wz_c + G_r1 wx_a dwy_b wz_c + G_r2 wx_a wy_b dwz_c
//           = wz_c (pa_r wy_b + qa_r dwy_b) + dwz_c (ra_r wy_b),      pa = G_.0 dwx_a, qa = G_.1 wx_a, ra = G_.2 wx_a
// as packed pairs over two consecutive z-nodes. sorted_ids is NOT re-ordered here: aux is indexed by sorted rank.
__global__ void __launch_bounds__(P2G_T, 2)
k_imp_scatter_tile(Planes P, const int* __restrict__ sorted_ids, const int4* __restrict__ pblock_list, DevCounters* dc,
                   const float4* __restrict__ aux, float4* __restrict__ G, GridDims gd, SimConst sc) {
    MPM_DYN_SMEM(imp_smem_raw, 16);
    ImpScatterSmem& S = *reinterpret_cast<ImpScatterSmem*>(imp_smem_raw);
    const int t = threadIdx.x, lane = t & 31;
    const int my_cell = t >> 2, my_a = t & 3;
    const int my_cx = my_cell >> 4, my_cy = (my_cell >> 2) & 3, my_cz = my_cell & 3;
    const int n_work = dc->n_active_pblocks;
    const float ih = 1.0f / sc.h;
    for (;;) {
        __syncthreads();          // the fold of the previous block is done with t1 and everyone has read S.work
        if (t == 0) {
            const int w = atomicAdd(&dc->work_a, 1);
            S.work = w < n_work ? pblock_list[w] : make_int4(-1, 0, 0, 0);
        }
        if (t < 64) S.cell_cnt[t] = 0;
        __syncthreads();
        const int4 wk = S.work;
        if (wk.x < 0) break;
        const int start = wk.y, cnt = wk.z;
        const int pbk = wk.w & (PB_COORD_MAX - 1), pbj = (wk.w >> PB_COORD_BITS) & (PB_COORD_MAX - 1), pbi = (wk.w >> (2 * PB_COORD_BITS)) + gd.lo;
        f32x2_t AX[8], AY[8], AZ[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) AX[i] = AY[i] = AZ[i] = 0ull;
        const int n_chunks = (cnt + P2G_CH - 1) / P2G_CH;
        for (int ck = 0; ck < n_chunks; ++ck) {
            const int nch = n_chunks == 1 ? cnt : (cnt - ck + n_chunks - 1) / n_chunks;
            if (ck > 0) {
                if (t < 64) S.cell_cnt[t] = 0;
                __syncthreads();
            }
            int cell_rank[P2G_PPT];
#pragma unroll
            for (int u = 0; u < P2G_PPT; ++u) {
                const int q = t + u * P2G_T;
                cell_rank[u] = -1;
                if (q < nch) {
                    const int j = start + ck + q * n_chunks;
                    const float4 xm = P.p[0][sorted_ids[j]];
                    const float4 g0 = aux[3 * (size_t)j], g1 = aux[3 * (size_t)j + 1], g2 = aux[3 * (size_t)j + 2];
                    float wx[4], wy[4], wz[4], dx[4], dy[4], dz[4];
                    const int cx = cell_of_t<0>(xm.x, sc.pd), cy = cell_of_t<0>(xm.y, sc.pd), cz = cell_of_t<0>(xm.z, sc.pd);
                    axis_weights_and_derivatives(xm.x, sc.pd, cx, wx, dx);
                    axis_weights_and_derivatives(xm.y, sc.pd, cy, wy, dy);
                    axis_weights_and_derivatives(xm.z, sc.pd, cz, wz, dz);
                    S.u.c.wx[q] = make_float4(wx[0], wx[1], wx[2], wx[3]);
                    S.u.c.wy[q] = make_float4(wy[0], wy[1], wy[2], wy[3]);
                    S.u.c.wz[q] = make_float4(wz[0], wz[1], wz[2], wz[3]);
                    S.u.c.dx[q] = make_float4(dx[0] * ih, dx[1] * ih, dx[2] * ih, dx[3] * ih);
                    S.u.c.dy[q] = make_float4(dy[0] * ih, dy[1] * ih, dy[2] * ih, dy[3] * ih);
                    S.u.c.dz[q] = make_float4(dz[0] * ih, dz[1] * ih, dz[2] * ih, dz[3] * ih);
                    S.u.c.g0[q] = g0; S.u.c.g1[q] = g1; S.u.c.g8[q] = g2.x;
                    const int lc = (((cx - 1) - 4 * pbi) * 4 + ((cy - 1) - 4 * pbj)) * 4 + ((cz - 1) - 4 * pbk);
                    cell_rank[u] = lc | (atomicAdd(&S.cell_cnt[lc], 1) << 8);
                }
            }
            __syncthreads();
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
#pragma unroll 1
            for (int i = i0; i < i1; ++i) {
                const int pi = S.u.c.order[i];
                const float wxa = reinterpret_cast<const float*>(&S.u.c.wx[pi])[my_a], dxa = reinterpret_cast<const float*>(&S.u.c.dx[pi])[my_a];
                const float4 wy = S.u.c.wy[pi], wz = S.u.c.wz[pi], dy = S.u.c.dy[pi], dz = S.u.c.dz[pi], g0 = S.u.c.g0[pi], g1 = S.u.c.g1[pi];
                const float g8 = S.u.c.g8[pi];
                // Gm column-major: G_r0 = (g0.x, g0.y, g0.z), G_r1 = (g0.w, g1.x, g1.y), G_r2 = (g1.z, g1.w, g8)
                const float pa[3] = { g0.x * dxa, g0.y * dxa, g0.z * dxa }, qa[3] = { g0.w * wxa, g1.x * wxa, g1.y * wxa }, ra[3] = { g1.z * wxa, g1.w * wxa, g8 * wxa };
                const float wyv[4] = { wy.x, wy.y, wy.z, wy.w }, dyv[4] = { dy.x, dy.y, dy.z, dy.w };
                const f32x2_t wzp[2] = { pack2(wz.x, wz.y), pack2(wz.z, wz.w) }, dzp[2] = { pack2(dz.x, dz.y), pack2(dz.z, dz.w) };
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const float u0 = fmaf(pa[0], wyv[bb], qa[0] * dyv[bb]), u1 = fmaf(pa[1], wyv[bb], qa[1] * dyv[bb]), u2 = fmaf(pa[2], wyv[bb], qa[2] * dyv[bb]);
                    const float v0 = ra[0] * wyv[bb], v1 = ra[1] * wyv[bb], v2 = ra[2] * wyv[bb];
                    const f32x2_t U0 = pack2(u0, u0), U1 = pack2(u1, u1), U2 = pack2(u2, u2), V0 = pack2(v0, v0), V1 = pack2(v1, v1), V2 = pack2(v2, v2);
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp) {
                        ffma2_acc(AX[bb * 2 + cp], wzp[cp], U0); ffma2_acc(AX[bb * 2 + cp], dzp[cp], V0);
                        ffma2_acc(AY[bb * 2 + cp], wzp[cp], U1); ffma2_acc(AY[bb * 2 + cp], dzp[cp], V1);
                        ffma2_acc(AZ[bb * 2 + cp], wzp[cp], U2); ffma2_acc(AZ[bb * 2 + cp], dzp[cp], V2);
                    }
                }
            }
            __syncthreads();
        }
        float4 acc[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[2 * i] = make_float4(0.f, lo2(AX[i]), lo2(AY[i]), lo2(AZ[i]));
            acc[2 * i + 1] = make_float4(0.f, hi2(AX[i]), hi2(AY[i]), hi2(AZ[i]));
        }
        // z-fold with warp shuffles (k_p2g_tile phase 2a), then x/y fold from shared memory, one vector red per tile node
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
                v.x = 0.f; v.y = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].y, src);
                v.z = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].z, src); v.w = __shfl_sync(0xffffffffu, acc[bb * 4 + cc].w, src);
                if (lo) { s0[bb].y += v.y; s0[bb].z += v.z; s0[bb].w += v.w; }
                else { s1[bb].y += v.y; s1[bb].z += v.z; s1[bb].w += v.w; }
            }
        }
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz] = s0[bb];
            if (my_cz < 3) S.u.t1[my_cx][my_cy][my_a][bb * 7 + my_cz + 4] = s1[bb];
        }
        __syncthreads();
        for (int n = t; n < 343; n += P2G_T) {
            const int ni = n / 49, nj = (n / 7) % 7, nk = n % 7;
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
#pragma unroll
                for (int dy = 0; dy < 4; ++dy) {
                    const int cx = ni - dx, cy = nj - dy;
                    if (cx >= 0 && cx <= 3 && cy >= 0 && cy <= 3) {
                        const float4 v = S.u.t1[cx][cy][dx][dy * 7 + nk];
                        sum.y += v.y; sum.z += v.z; sum.w += v.w;
                    }
                }
            if (sum.y != 0.f || sum.z != 0.f || sum.w != 0.f) atomicAdd(&G[node_index(gd, 4 * pbi + ni, 4 * pbj + nj, 4 * pbk + nk)], sum);
        }
    }
}

// ---- vectors over the active nodes (float4 per node; .x unused and kept 0) ----
// z = a x + b y (any of them may alias)
__global__ void __launch_bounds__(256)
k_vec_lin(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, float4* z, float a, const float4* x, float b, const float4* y,
          const float* __restrict__ b_dev = nullptr) {
    if (b_dev) b = *b_dev;                 // coefficient left on the device by k_lbfgs_coef
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    for (int blk = blockIdx.x * per + sub; blk < nb; blk += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[blk] * 64 + t;
        const float4 xv = x[idx];
        float4 r = make_float4(0.f, a * xv.y, a * xv.z, a * xv.w);
        if (y) { const float4 yv = y[idx]; r.y += b * yv.y; r.z += b * yv.z; r.w += b * yv.w; }
        z[idx] = r;
    }
}
// out[0] += x . y, out[1] = max(out[1], |x|_inf) (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256)
k_vec_dot(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, const float4* __restrict__ x, const float4* __restrict__ y,
          double* __restrict__ out, int* __restrict__ absmax_bits) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    double s = 0.0;
    float m = 0.0f;
    for (int blk = blockIdx.x * per + sub; blk < nb; blk += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[blk] * 64 + t;
        const float4 xv = x[idx], yv = y[idx];
        s += (double)xv.y * yv.y + (double)xv.z * yv.z + (double)xv.w * yv.w;
        m = fmaxf(m, fmaxf(fabsf(xv.y), fmaxf(fabsf(xv.z), fabsf(xv.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0) {
        if (s != 0.0) atomicAdd(out, s);
        if (absmax_bits && m > 0.0f) atomicMax(absmax_bits, __float_as_int(m));
    }
}
// scalars of the L-BFGS two-loop recursion (LBFGS.hpp:88-99) without a host round trip: sc[0] holds the dot product k_vec_dot
// has just accumulated; first loop: alpha_i = rho_i (s_i . q), coefficient -alpha_i; second loop: beta = rho_i (y_i . q),
// coefficient alpha_i - beta. fc[0..7] = alpha, fc[8] = the coefficient the next k_vec_lin reads. Float like the reference.
__global__ void k_lbfgs_coef(double* __restrict__ sc, float* __restrict__ fc, int i, float rho, int second) {
    const float d = (float)sc[0];
    if (!second) { fc[i] = rho * d; fc[8] = -fc[i]; }
    else fc[8] = fc[i] - rho * d;
    sc[0] = 0.0;
}
// trial velocities <-> grid: X = (0, v) of the grid; grid velocity = X where the node has mass
__global__ void __launch_bounds__(256)
k_imp_from_grid(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, const float4* __restrict__ grid, float4* __restrict__ X) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    for (int blk = blockIdx.x * per + sub; blk < nb; blk += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[blk] * 64 + t;
        const float4 n = grid[idx];
        X[idx] = make_float4(0.f, n.y, n.z, n.w);
    }
}
__global__ void __launch_bounds__(256)
k_imp_to_grid(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, float4* __restrict__ grid, const float4* __restrict__ X) {
    const int nb = dc->n_active_gblocks;
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, per = blockDim.x >> 6;
    for (int blk = blockIdx.x * per + sub; blk < nb; blk += gridDim.x * per) {
        const size_t idx = (size_t)gblock_list[blk] * 64 + t;
        const float4 n = grid[idx];
        if (n.x != 0.0f) { const float4 x = X[idx]; grid[idx] = make_float4(n.x, x.y, x.z, x.w); }
    }
}
// dense host-order (i*J*K + j*K + k) x 3 arrays <-> grid layout (diagnostic entry points: mpm_energy, mpm_energy_gradient)
__global__ void k_imp_import(float4* __restrict__ X, const float4* __restrict__ grid, GridDims gd, const float* __restrict__ in3, int add_to_grid_velocity) {
    const size_t n = (size_t)gd.I * gd.J * gd.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % gd.K), j = (int)((t / gd.K) % gd.J), i = (int)(t / ((size_t)gd.K * gd.J));
    const size_t idx = node_index(gd, i, j, k);
    const float4 g = add_to_grid_velocity ? grid[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    X[idx] = make_float4(0.f, g.y + in3[t * 3], g.z + in3[t * 3 + 1], g.w + in3[t * 3 + 2]);
}
__global__ void k_imp_export(const float4* __restrict__ G, GridDims gd, float* __restrict__ out3) {
    const size_t n = (size_t)gd.I * gd.J * gd.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % gd.K), j = (int)((t / gd.K) % gd.J), i = (int)(t / ((size_t)gd.K * gd.J));
    const float4 g = G[node_index(gd, i, j, k)];
    out3[t * 3] = g.y; out3[t * 3 + 1] = g.z; out3[t * 3 + 2] = g.w;
}

#if !defined(MPM_HOST_EMU) || defined(MPM_HOST_EMU_API)
inline cudaError_t implicit_kernels_init() {
    return cudaFuncSetAttribute(k_imp_scatter_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ImpScatterSmem));
}
#endif

}  // namespace mpm

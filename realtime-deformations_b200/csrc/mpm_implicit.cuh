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
#include "mpm_kernels.cuh"

namespace mpm {

struct ImplicitConst { float mu0, lambda0, xi; int hardening; };

// weights and weight derivatives of the four stencil nodes cell-1 .. cell+2 (offsets fx+1, fx, fx-1, fx-2: the branches of
// hpp:20-52 are known statically)
MPM_DI void axis_weights_and_derivatives(float x, const PosDiv& d, int cell, float w[4], float dw[4]) {
    const float fx = sub_rn(pos_div(x, d), (float)cell);
    const float gx = 1.0f - fx;
    w[0] = 0.16666667163372040f * gx * gx * gx;
    w[1] = fmaf(fmaf(0.5f, fx, -1.0f), fx * fx, 0.66666668653488159f);
    w[2] = fmaf(fmaf(0.5f, gx, -1.0f), gx * gx, 0.66666668653488159f);
    w[3] = 0.16666667163372040f * fx * fx * fx;
    dw[0] = -0.5f * gx * gx;
    dw[1] = fx * fmaf(1.5f, fx, -2.0f);
    dw[2] = gx * fmaf(-1.5f, gx, 2.0f);
    dw[3] = 0.5f * fx * fx;
}

// rotation factor of the polar decomposition by Newton's iteration R <- (R + R^-T)/2 in double; false if singular
MPM_DI bool polar_rotation_d(const double (&F)[9], double (&R)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = F[i];
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
        if (diff < 1e-15) break;
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

// ---- vectors over the active nodes (float4 per node; .x unused and kept 0) ----
// z = a x + b y (any of them may alias)
__global__ void __launch_bounds__(256)
k_vec_lin(const int* __restrict__ gblock_list, const DevCounters* __restrict__ dc, float4* z, float a, const float4* x, float b, const float4* y) {
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

}  // namespace mpm

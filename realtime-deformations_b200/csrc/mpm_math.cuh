// Device-side 3x3 / weight / SVD arithmetic of the MPM substep.
//
// Two families live here:
//  * "_rn" routines reproduce the reference's CPU arithmetic operation by operation (glm 0.9.7.1 association,
//    Eigen 3.4.90 JacobiSVD control flow) with __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn, which nvcc never
//    contracts into FMAs. They are used where the reference result is a pure function of per-particle or
//    per-node inputs (F-update, advection, grid update, collisions, weights), so those stages are bit-exact
//    against the CPU path given identical inputs.
//  * plain routines (FMA allowed) are used inside the order-nondeterministic sums (P2G scatter, G2P gather).
//
// 3x3 matrices are float[9] in glm column-major order: m[c*3+r] == glm m[c][r].
#pragma once
#include <cuda_runtime.h>
#include <cfloat>

namespace mpm {

#define MPM_DI __device__ __forceinline__

MPM_DI float mul_rn(float a, float b) { return __fmul_rn(a, b); }
MPM_DI float add_rn(float a, float b) { return __fadd_rn(a, b); }
MPM_DI float sub_rn(float a, float b) { return __fsub_rn(a, b); }
MPM_DI float div_rn(float a, float b) { return __fdiv_rn(a, b); }
MPM_DI float rcp_rn(float a) { return __frcp_rn(a); }      // correctly rounded 1/a == IEEE 1.0f / a

// packed fp32 pairs (sm_100a FFMA2 / FADD2: two IEEE-rounded operations per lane per instruction; ptxas folds a duplicated
// {x, x} operand into a scalar broadcast). Used by the accumulation loop of P2G, the gather and the implicit scatter.
#ifndef MPM_HOST_EMU
typedef unsigned long long f32x2_t;
MPM_DI f32x2_t pack2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
MPM_DI float lo2(f32x2_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
MPM_DI float hi2(f32x2_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
MPM_DI f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
MPM_DI void ffma2_acc(f32x2_t& c, f32x2_t a, f32x2_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }   // c += a * b, in place
MPM_DI f32x2_t fadd2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
MPM_DI f32x2_t fmul2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#else   // host emulation of the kernels (tests/emu): same per-component IEEE operations
typedef unsigned long long f32x2_t;
MPM_DI f32x2_t pack2(float lo, float hi) { float v[2] = { lo, hi }; f32x2_t r; memcpy(&r, v, 8); return r; }
MPM_DI float lo2(f32x2_t v) { float f[2]; memcpy(f, &v, 8); return f[0]; }
MPM_DI float hi2(f32x2_t v) { float f[2]; memcpy(f, &v, 8); return f[1]; }
MPM_DI f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) { return pack2(fmaf(lo2(a), lo2(b), lo2(c)), fmaf(hi2(a), hi2(b), hi2(c))); }
MPM_DI void ffma2_acc(f32x2_t& c, f32x2_t a, f32x2_t b) { c = ffma2(a, b, c); }
MPM_DI f32x2_t fadd2(f32x2_t a, f32x2_t b) { return pack2(__fadd_rn(lo2(a), lo2(b)), __fadd_rn(hi2(a), hi2(b))); }
MPM_DI f32x2_t fmul2(f32x2_t a, f32x2_t b) { return pack2(__fmul_rn(lo2(a), lo2(b)), __fmul_rn(hi2(a), hi2(b))); }
#endif

// (x', y') = (c*x + s*y, -s*x + c*y), every product and every sum rounded once: the Jacobi rotation of Eigen's
// applyOnTheLeft / applyOnTheRight. (A packed mul.rn.f32x2 / add.rn.f32x2 form of this was tried in round 1 and is gone:
// on hardware it did NOT reproduce the scalar bits -- 36512 of 38646 F-update KAT values differed, profiles/r2_ab_64M.md.)
MPM_DI void rot_pair_rn(float c, float s, float x, float y, float& xo, float& yo) {
    const float a = add_rn(mul_rn(c, x), mul_rn(s, y));
    const float b = add_rn(mul_rn(-s, x), mul_rn(c, y));
    xo = a; yo = b;
}

// ---- reference weightNx, material_point_method.hpp:20-31 ----
// The reference evaluates the polynomial in fp64 and rounds to fp32 (weight_nx_exact reproduces that bit for bit).
// The hot kernels use the fp32/FMA form below: it differs from the exact value by <= 1 ulp (6e-8 relative), far
// inside the fp32 noise of the scatter/gather sums the weights feed, and avoids the F2F/FP64 pipes.
MPM_DI float weight_nx(float x) {
    const float m = fabsf(x);
    if (m < 1.0f) return fmaf(fmaf(0.5f, m, -1.0f), m * m, 0.66666668653488159f);
    if (m < 2.0f) { const float a = 2.0f - m; return 0.16666667163372040f * a * a * a; }
    return 0.0f;
}
MPM_DI float weight_nx_exact(float x) {
    const float modx = fabsf(x);
    const float modx2 = mul_rn(modx, modx);
    const float modx3 = mul_rn(mul_rn(modx, modx), modx);
    if (modx < 1.0f) return (float)__dadd_rn(__dsub_rn(__dmul_rn(0.5, (double)modx3), (double)modx2), 2.0 / 3.0);
    if (modx < 2.0f) {
        const double a = (double)sub_rn(2.0f, modx);
        return (float)__dmul_rn(__dmul_rn(__dmul_rn(1.0 / 6.0, a), a), a);
    }
    return 0.0f;
}

// pos / h must be the IEEE-754 quotient (the cell index ivec3(pos / h), material_point_method.cpp:83, is bit-exact
// index work). h is a per-scene constant, so the quotient is formed as q0 = x*rh, q = fma(fma(-q0, h, x), rh, q0) with
// rh = RN(1/h): three dependent FMA-pipe ops instead of the ~10-instruction __fdiv_rn sequence. That shortcut is
// only used on [lo, hi], where mpm_create has checked it EXHAUSTIVELY (every fp32 in the range) against __fdiv_rn
// for this h; elsewhere, or if the check ever failed, the exact intrinsic runs.
struct PosDiv { float h, rh, lo, hi; int fast; int quadratic; };    // quadratic: MpmParams.stencil = 1 (see cell_of below)
MPM_DI float pos_div(float x, const PosDiv& d) {
    if (d.fast && x >= d.lo && x <= d.hi) {
        const float q0 = mul_rn(x, d.rh);
        return __fmaf_rn(__fmaf_rn(-q0, d.h, x), d.rh, q0);
    }
    return div_rn(x, d.h);
}
// cell index and the four non-zero per-axis weights (nodes cell-1 .. cell+2).
// material_point_method.cpp:83 (ivec3(pos / h): IEEE divide + truncation), hpp:53-58 (pos/h - idx).
// Stencil order (MpmParams.stencil): with the QUADRATIC B-spline -- not reference behaviour -- a particle
// touches the three nodes base .. base+2 per axis, base = int(pos/h - 1/2). Every kernel addresses a particle's stencil as
// "cell - 1 + a", a = 0..3, so the quadratic case simply reports the pseudo-cell base + 1 and a fourth weight of zero: block
// keys, tiles, local cell numbers and node offsets need no second code path (the tile kernels additionally have W = 3
// instantiations that skip the zero column).
// Q selects the stencil at COMPILE time in the hot kernels (0 cubic, 1 quadratic) or at run time (2: d.quadratic), so that the
// benchmark path (cubic) carries no trace of the switch.
template <int Q>
MPM_DI int cell_of_t(float x, const PosDiv& d) {
    const float q = pos_div(x, d);
    const bool quad = Q == 1 || (Q == 2 && d.quadratic);
    return quad ? __float2int_rz(q - 0.5f) + 1 : __float2int_rz(q);
}
MPM_DI int cell_of(float x, const PosDiv& d) { return cell_of_t<2>(x, d); }
// quadratic weights at offset t = q - (base + 1) in [-1/2, 1/2): N(t + 1), N(t), N(t - 1) with N(x) = 3/4 - x^2 (|x| < 1/2),
// (3/2 - |x|)^2 / 2 (|x| < 3/2)
MPM_DI void quadratic_weights(float t, float w[4]) {
    const float a = 0.5f - t, b = 0.5f + t;
    w[0] = 0.5f * a * a;
    w[1] = fmaf(-t, t, 0.75f);
    w[2] = 0.5f * b * b;
    w[3] = 0.0f;
}
// The four stencil offsets are q-(cell-1) = fx+1, fx, fx-1, fx-2 with fx = q - cell in [0,1) (all exact in fp32), so
// the reference's |x|<1 / |x|<2 branches are known statically: far, near, near, far. Branch-free, <= 1 ulp from
// weight_nx_exact, partition of unity to fp32 rounding.
MPM_DI void axis_weights(float x, const PosDiv& d, int cell, float w[4]) {
    const float fx = sub_rn(pos_div(x, d), (float)cell);
    if (d.quadratic) { quadratic_weights(fx, w); return; }
    const float gx = 1.0f - fx;
    w[0] = 0.16666667163372040f * gx * gx * gx;
    w[1] = fmaf(fmaf(0.5f, fx, -1.0f), fx * fx, 0.66666668653488159f);
    w[2] = fmaf(fmaf(0.5f, gx, -1.0f), gx * gx, 0.66666668653488159f);
    w[3] = 0.16666667163372040f * fx * fx * fx;
}
// cubic weights and weight DERIVATIVES (hpp:32-52; implicit time integration, mpm_implicit.cuh) of the four stencil nodes
// cell-1 .. cell+2: offsets fx+1, fx, fx-1, fx-2, so the branches of hpp:20-52 are known statically
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

// same as cell_of + axis_weights with the quotient formed once
template <int Q>
MPM_DI int cell_and_weights_t(float x, const PosDiv& d, float w[4]) {
    const float q = pos_div(x, d);
    if (Q == 1 || (Q == 2 && d.quadratic)) {
        const int cell = __float2int_rz(q - 0.5f) + 1;
        quadratic_weights(sub_rn(q, (float)cell), w);
        return cell;
    }
    const int cell = __float2int_rz(q);
    const float fx = sub_rn(q, (float)cell);
    const float gx = 1.0f - fx;
    w[0] = 0.16666667163372040f * gx * gx * gx;
    w[1] = fmaf(fmaf(0.5f, fx, -1.0f), fx * fx, 0.66666668653488159f);
    w[2] = fmaf(fmaf(0.5f, gx, -1.0f), gx * gx, 0.66666668653488159f);
    w[3] = 0.16666667163372040f * fx * fx * fx;
    return cell;
}
MPM_DI int cell_and_weights(float x, const PosDiv& d, float w[4]) { return cell_and_weights_t<2>(x, d, w); }
MPM_DI void axis_weights_exact(float x, const PosDiv& d, int cell, float w[4]) {
    const float q = pos_div(x, d);
#pragma unroll
    for (int d = 0; d < 4; ++d) w[d] = weight_nx_exact(sub_rn(q, (float)(cell - 1 + d)));
}

// ---- glm value-type arithmetic, reference association (see oracle/mpm_oracle.c for the file:line map) ----
MPM_DI float dot3_rn(float a0, float b0, float a1, float b1, float a2, float b2) {
    return add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2));
}
MPM_DI void m3_mul_rn(float* R, const float* A, const float* B) {   // R may not alias A or B
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            R[c * 3 + r] = dot3_rn(A[0 + r], B[c * 3 + 0], A[3 + r], B[c * 3 + 1], A[6 + r], B[c * 3 + 2]);
}
#define MG(m, c, r) ((m)[(c) * 3 + (r)])
MPM_DI float m3_det_rn(const float* m) {
    const float t0 = mul_rn(MG(m,0,0), sub_rn(mul_rn(MG(m,1,1), MG(m,2,2)), mul_rn(MG(m,2,1), MG(m,1,2))));
    const float t1 = mul_rn(MG(m,1,0), sub_rn(mul_rn(MG(m,0,1), MG(m,2,2)), mul_rn(MG(m,2,1), MG(m,0,2))));
    const float t2 = mul_rn(MG(m,2,0), sub_rn(mul_rn(MG(m,0,1), MG(m,1,2)), mul_rn(MG(m,1,1), MG(m,0,2))));
    return add_rn(sub_rn(t0, t1), t2);
}
MPM_DI void m3_inverse_rn(float* R, const float* m) {               // R may not alias m
    const float ood = rcp_rn(m3_det_rn(m));
    MG(R,0,0) =  mul_rn(sub_rn(mul_rn(MG(m,1,1), MG(m,2,2)), mul_rn(MG(m,2,1), MG(m,1,2))), ood);
    MG(R,1,0) = -mul_rn(sub_rn(mul_rn(MG(m,1,0), MG(m,2,2)), mul_rn(MG(m,2,0), MG(m,1,2))), ood);
    MG(R,2,0) =  mul_rn(sub_rn(mul_rn(MG(m,1,0), MG(m,2,1)), mul_rn(MG(m,2,0), MG(m,1,1))), ood);
    MG(R,0,1) = -mul_rn(sub_rn(mul_rn(MG(m,0,1), MG(m,2,2)), mul_rn(MG(m,2,1), MG(m,0,2))), ood);
    MG(R,1,1) =  mul_rn(sub_rn(mul_rn(MG(m,0,0), MG(m,2,2)), mul_rn(MG(m,2,0), MG(m,0,2))), ood);
    MG(R,2,1) = -mul_rn(sub_rn(mul_rn(MG(m,0,0), MG(m,2,1)), mul_rn(MG(m,2,0), MG(m,0,1))), ood);
    MG(R,0,2) =  mul_rn(sub_rn(mul_rn(MG(m,0,1), MG(m,1,2)), mul_rn(MG(m,1,1), MG(m,0,2))), ood);
    MG(R,1,2) = -mul_rn(sub_rn(mul_rn(MG(m,0,0), MG(m,1,2)), mul_rn(MG(m,1,0), MG(m,0,2))), ood);
    MG(R,2,2) =  mul_rn(sub_rn(mul_rn(MG(m,0,0), MG(m,1,1)), mul_rn(MG(m,1,0), MG(m,0,1))), ood);
}
MPM_DI void m3_transpose(float* R, const float* A) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) R[c * 3 + r] = A[r * 3 + c];
}

// ---- Eigen 3.4.90 JacobiSVD<MatrixXf, ComputeFullU|ComputeFullV> on a 3x3, in registers ----
// Control flow, sweep order (1,0),(2,0),(2,1), 2x2 kernel, sign fix-up and selection sort follow
// external/Eigen/src/SVD/JacobiSVD.h:689-817, misc/RealSvd2x2.h:21-51, Jacobi/Jacobi.h:96-126,326-337
// (restated in oracle/mpm_oracle.c: oracle_svd3). Arrays are ROW-major here (a[r*3+c] == Eigen m(r,c)).
template <int P, int Q>
MPM_DI void jacobi_pair(float (&W)[9], float (&U)[9], float (&V)[9], float& maxDiag, bool& finished) {
    const float pm = mul_rn(2.0f * FLT_EPSILON, maxDiag);
    const float threshold = (FLT_MIN < pm) ? pm : FLT_MIN;
    if (!(fabsf(W[P * 3 + Q]) > threshold || fabsf(W[Q * 3 + P]) > threshold)) return;
    finished = false;
    float m00 = W[P * 3 + P], m01 = W[P * 3 + Q], m10 = W[Q * 3 + P], m11 = W[Q * 3 + Q];
    float c1, s1;
    const float t = add_rn(m00, m11), d = sub_rn(m10, m01);
    if (fabsf(d) < FLT_MIN) { s1 = 0.0f; c1 = 1.0f; }
    else {
        const float u = div_rn(t, d);
        const float tmp = __fsqrt_rn(add_rn(1.0f, mul_rn(u, u)));
        s1 = rcp_rn(tmp); c1 = div_rn(u, tmp);
    }
    if (!(c1 == 1.0f && s1 == 0.0f)) {
        float a0, b0, a1, b1;
        rot_pair_rn(c1, s1, m00, m10, a0, b0);
        rot_pair_rn(c1, s1, m01, m11, a1, b1);
        m00 = a0; m10 = b0; m01 = a1; m11 = b1;
    }
    float cr, sr;
    const float deno = mul_rn(2.0f, fabsf(m01));
    if (deno < FLT_MIN) { cr = 1.0f; sr = 0.0f; }
    else {
        const float tau = div_rn(sub_rn(m00, m11), deno);
        const float w = __fsqrt_rn(add_rn(mul_rn(tau, tau), 1.0f));
        const float tt = rcp_rn(tau > 0.0f ? add_rn(tau, w) : sub_rn(tau, w));
        const float nn = rcp_rn(__fsqrt_rn(add_rn(mul_rn(tt, tt), 1.0f)));
        // Eigen: s = -sign_t * (y/|y|) * |t| * n. The first three factors are exact (+-1, +-1, |t|): -sign_t*|t| == -t,
        // so s = RN((-t * sgn(y)) * n) -- one rounding, identical bits.
        sr = mul_rn(copysignf(1.0f, m01) * -tt, nn);
        cr = nn;
    }
    const float cl = sub_rn(mul_rn(c1, cr), mul_rn(s1, -sr));
    const float sl = add_rn(mul_rn(c1, -sr), mul_rn(s1, cr));
    if (!(cl == 1.0f && sl == 0.0f)) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {       // W.applyOnTheLeft(p,q,j_left)
            rot_pair_rn(cl, sl, W[P * 3 + c], W[Q * 3 + c], W[P * 3 + c], W[Q * 3 + c]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {       // U.applyOnTheRight(p,q,j_left.transpose())
            rot_pair_rn(cl, sl, U[r * 3 + P], U[r * 3 + Q], U[r * 3 + P], U[r * 3 + Q]);
        }
    }
    const float s = -sr;
    if (!(cr == 1.0f && s == 0.0f)) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {       // W.applyOnTheRight(p,q,j_right); V.applyOnTheRight(p,q,j_right)
            rot_pair_rn(cr, s, W[r * 3 + P], W[r * 3 + Q], W[r * 3 + P], W[r * 3 + Q]);
            rot_pair_rn(cr, s, V[r * 3 + P], V[r * 3 + Q], V[r * 3 + P], V[r * 3 + Q]);
        }
    }
    const float dm = fmaxf(fabsf(W[P * 3 + P]), fabsf(W[Q * 3 + Q]));
    if (maxDiag < dm) maxDiag = dm;
}

// returns false on non-finite input (Eigen: InvalidInput). MAX_SWEEPS only guards the GPU against a hang;
// Eigen itself has no cap and converges in 3-5 sweeps on finite input.
MPM_DI bool svd3_eigen(const float (&A)[9], float (&U)[9], float (&S)[3], float (&V)[9]) {
    float W[9];
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {           // maxCoeff<PropagateNaN>: a NaN, once seen, sticks
        const float a = fabsf(A[i]);
        if (!(scale != scale) && (a != a || a > scale)) scale = a;
    }
    if (!isfinite(scale)) return false;
    if (scale == 0.0f) scale = 1.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { W[i] = div_rn(A[i], scale); U[i] = (i % 4 == 0) ? 1.0f : 0.0f; V[i] = U[i]; }
    float maxDiag = fmaxf(fmaxf(fabsf(W[0]), fabsf(W[4])), fabsf(W[8]));
    bool finished = false;
    constexpr int MAX_SWEEPS = 64;
    for (int sweep = 0; sweep < MAX_SWEEPS && !finished; ++sweep) {
        finished = true;
        jacobi_pair<1, 0>(W, U, V, maxDiag, finished);
        jacobi_pair<2, 0>(W, U, V, maxDiag, finished);
        jacobi_pair<2, 1>(W, U, V, maxDiag, finished);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = W[i * 3 + i];
        S[i] = mul_rn(fabsf(a), scale);
        if (a < 0.0f) { U[0 + i] = -U[0 + i]; U[3 + i] = -U[3 + i]; U[6 + i] = -U[6 + i]; }
    }
    // selection sort, descending, first maximum wins (JacobiSVD.h:795-814)
#define MPM_SWAPCOL(i, j)                                                          \
    do {                                                                           \
        float t_ = S[i]; S[i] = S[j]; S[j] = t_;                                   \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {                            \
            t_ = U[r * 3 + i]; U[r * 3 + i] = U[r * 3 + j]; U[r * 3 + j] = t_;     \
            t_ = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + j]; V[r * 3 + j] = t_;     \
        }                                                                          \
    } while (0)
    {
        int pos = 0;
        if (S[1] > S[pos]) pos = 1;
        if (S[2] > S[pos]) pos = 2;
        if (S[pos] == 0.0f) return true;
        if (pos == 1) MPM_SWAPCOL(0, 1); else if (pos == 2) MPM_SWAPCOL(0, 2);
        if (S[2] > S[1]) { MPM_SWAPCOL(1, 2); }
    }
#undef MPM_SWAPCOL
    return true;
}

// ---- F-update of one particle, material_point_method.cpp:306-330 (bit-faithful) ----
// in: B (previous substep's APIC matrix), FE, FP; out: FE, FP overwritten, factors of the new FE for the stress.
// Ug/Sg: FE_new = Ug * diag(Sg) * Vg^T as glm matrices (Ug is the glm view of Eigen's U, see utils.h:25-33).
MPM_DI bool f_update_rn(const float (&B)[9], float (&FE)[9], float (&FP)[9], float dinv, float dt, float clamp_lo,
                        float clamp_hi, float (&Ug)[9], float (&Sg)[3]) {
    float T0[9], T1[9], T[9], FPinv[9], Fh[9], Vg[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {           // m3t(1.0) + B * DpInverse * dt
        const float v = mul_rn(mul_rn(B[i], dinv), dt);
        T0[i] = add_rn((i % 4 == 0) ? 1.0f : 0.0f, v);
    }
    m3_mul_rn(T1, T0, FE);
    m3_mul_rn(T, T1, FP);             // FPn1
    m3_inverse_rn(FPinv, FP);
    m3_mul_rn(Fh, T, FPinv);          // FEpKryshka
    if (!svd3_eigen(Fh, Ug, Sg, Vg)) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) { float s = Sg[k]; if (s < clamp_lo) s = clamp_lo; if (clamp_hi < s) s = clamp_hi; Sg[k] = s; }
    float US[9], Vt[9], FEinv[9];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) US[c * 3 + r] = mul_rn(Ug[c * 3 + r], Sg[c]);   // U * S (S diagonal)
    m3_transpose(Vt, Vg);
    m3_mul_rn(FE, US, Vt);            // U * S * transpose(V)
    m3_inverse_rn(FEinv, FE);
    m3_mul_rn(FP, FEinv, T);
    return true;
}

// ---- F-update, tolerance form (the fused substep's default; the staged API and fupdate_exact keep f_update_rn) ----
// Same algorithm and control flow as f_update_rn / svd3_eigen above (Eigen's two-sided Jacobi sweeps in the order
// (1,0),(2,0),(2,1), its 2x2 kernel, threshold, sign fix-up and descending selection sort; the reference's transposed
// re-assembly), but with the freedoms a floating-point tolerance gives: products and sums contract into FMAs, divisions
// and square roots are single MUFU approximations (<= 2 ulp), the max|a| pre-scaling is dropped (it only guards Eigen
// against overflow; F is O(1) here) and F^ = (I + dt C) FE is taken directly instead of ((I + dt C) FE FP) FP^-1
// (cpp:309-311 multiplies by FP and by its inverse again). About 350 instead of 1050 instructions per particle at rest.
// Differences to the bit-faithful form are of the size of the reference's own sensitivity to FMA contraction (SURVEY
// App. C): the trajectory tests hold this path to 4x that noise floor.
#ifndef MPM_HOST_EMU
MPM_DI float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MPM_DI float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MPM_DI float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MPM_DI float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
MPM_DI float fast_rcp(float x) { return 1.0f / x; }
MPM_DI float fast_rsqrt(float x) { return 1.0f / std::sqrt(x); }
MPM_DI float fast_sqrt(float x) { return std::sqrt(x); }
MPM_DI float fast_ex2(float x) { return std::exp2(x); }
#endif
MPM_DI void rot_pair_fast(float c, float s, float x, float y, float& xo, float& yo) {
    const float a = fmaf(c, x, s * y), b = fmaf(c, y, -s * x);
    xo = a; yo = b;
}
MPM_DI void m3_mul_fast(float* R, const float* A, const float* B) {   // R may not alias A or B
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            R[c * 3 + r] = fmaf(A[6 + r], B[c * 3 + 2], fmaf(A[3 + r], B[c * 3 + 1], A[0 + r] * B[c * 3 + 0]));
}
MPM_DI float m3_det_fast(const float* m) {
    return MG(m,0,0) * (MG(m,1,1) * MG(m,2,2) - MG(m,2,1) * MG(m,1,2))
         - MG(m,1,0) * (MG(m,0,1) * MG(m,2,2) - MG(m,2,1) * MG(m,0,2))
         + MG(m,2,0) * (MG(m,0,1) * MG(m,1,2) - MG(m,1,1) * MG(m,0,2));
}
MPM_DI void m3_inverse_fast(float* R, const float* m) {             // R may not alias m; adjugate / det
    const float c00 = MG(m,1,1) * MG(m,2,2) - MG(m,2,1) * MG(m,1,2);
    const float c01 = MG(m,2,1) * MG(m,0,2) - MG(m,0,1) * MG(m,2,2);
    const float c02 = MG(m,0,1) * MG(m,1,2) - MG(m,1,1) * MG(m,0,2);
    const float ood = fast_rcp(fmaf(MG(m,2,0), c02, fmaf(MG(m,1,0), c01, MG(m,0,0) * c00)));
    MG(R,0,0) = c00 * ood; MG(R,0,1) = c01 * ood; MG(R,0,2) = c02 * ood;
    MG(R,1,0) = (MG(m,2,0) * MG(m,1,2) - MG(m,1,0) * MG(m,2,2)) * ood;
    MG(R,1,1) = (MG(m,0,0) * MG(m,2,2) - MG(m,2,0) * MG(m,0,2)) * ood;
    MG(R,1,2) = (MG(m,1,0) * MG(m,0,2) - MG(m,0,0) * MG(m,1,2)) * ood;
    MG(R,2,0) = (MG(m,1,0) * MG(m,2,1) - MG(m,2,0) * MG(m,1,1)) * ood;
    MG(R,2,1) = (MG(m,2,0) * MG(m,0,1) - MG(m,0,0) * MG(m,2,1)) * ood;
    MG(R,2,2) = (MG(m,0,0) * MG(m,1,1) - MG(m,1,0) * MG(m,0,1)) * ood;
}
template <int P, int Q>
MPM_DI void jacobi_pair_fast(float (&W)[9], float (&U)[9], float (&V)[9], float& maxDiag, bool& finished) {
    const float threshold = fmaxf(FLT_MIN, (2.0f * FLT_EPSILON) * maxDiag);
    if (!(fabsf(W[P * 3 + Q]) > threshold || fabsf(W[Q * 3 + P]) > threshold)) return;
    finished = false;
    float m00 = W[P * 3 + P], m01 = W[P * 3 + Q], m10 = W[Q * 3 + P], m11 = W[Q * 3 + Q];
    float c1 = 1.0f, s1 = 0.0f;
    const float t = m00 + m11, d = m10 - m01;
    const float n2 = fmaf(t, t, d * d);
    if (!(fabsf(d) < FLT_MIN) && n2 >= FLT_MIN) {       // u = t/d; s = 1/sqrt(1+u^2) = |d|/sqrt(t^2+d^2); c = u*s
        const float rs = fast_rsqrt(n2);
        s1 = fabsf(d) * rs; c1 = copysignf(t * rs, t * d);
        rot_pair_fast(c1, s1, m00, m10, m00, m10);
        rot_pair_fast(c1, s1, m01, m11, m01, m11);
    }
    float cr = 1.0f, sr = 0.0f;
    const float deno = 2.0f * fabsf(m01);
    if (!(deno < FLT_MIN)) {
        const float tau = (m00 - m11) * fast_rcp(deno);
        const float w = fast_sqrt(fmaf(tau, tau, 1.0f));
        const float tt = fast_rcp(tau > 0.0f ? tau + w : tau - w);
        const float nn = fast_rsqrt(fmaf(tt, tt, 1.0f));
        sr = copysignf(1.0f, m01) * -tt * nn;
        cr = nn;
    }
    const float cl = fmaf(c1, cr, s1 * sr);              // j_left = rot1 * j_right^T
    const float sl = fmaf(s1, cr, -c1 * sr);
#pragma unroll
    for (int c = 0; c < 3; ++c) rot_pair_fast(cl, sl, W[P * 3 + c], W[Q * 3 + c], W[P * 3 + c], W[Q * 3 + c]);
#pragma unroll
    for (int r = 0; r < 3; ++r) rot_pair_fast(cl, sl, U[r * 3 + P], U[r * 3 + Q], U[r * 3 + P], U[r * 3 + Q]);
    const float s = -sr;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rot_pair_fast(cr, s, W[r * 3 + P], W[r * 3 + Q], W[r * 3 + P], W[r * 3 + Q]);
        rot_pair_fast(cr, s, V[r * 3 + P], V[r * 3 + Q], V[r * 3 + P], V[r * 3 + Q]);
    }
    maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[P * 3 + P]), fabsf(W[Q * 3 + Q])));
}
MPM_DI bool svd3_fast(const float (&A)[9], float (&U)[9], float (&S)[3], float (&V)[9]) {
    float W[9];
    float chk = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { W[i] = A[i]; chk += fabsf(A[i]); U[i] = (i % 4 == 0) ? 1.0f : 0.0f; V[i] = U[i]; }
    if (!isfinite(chk)) return false;                    // Eigen: InvalidInput
    float maxDiag = fmaxf(fmaxf(fabsf(W[0]), fabsf(W[4])), fabsf(W[8]));
    bool finished = false;
    constexpr int MAX_SWEEPS = 64;
    for (int sweep = 0; sweep < MAX_SWEEPS && !finished; ++sweep) {
        finished = true;
        jacobi_pair_fast<1, 0>(W, U, V, maxDiag, finished);
        jacobi_pair_fast<2, 0>(W, U, V, maxDiag, finished);
        jacobi_pair_fast<2, 1>(W, U, V, maxDiag, finished);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = W[i * 3 + i];
        S[i] = fabsf(a);
        if (a < 0.0f) { U[0 + i] = -U[0 + i]; U[3 + i] = -U[3 + i]; U[6 + i] = -U[6 + i]; }
    }
#define MPM_SWAPCOL(i, j)                                                          \
    do {                                                                           \
        float t_ = S[i]; S[i] = S[j]; S[j] = t_;                                   \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {                            \
            t_ = U[r * 3 + i]; U[r * 3 + i] = U[r * 3 + j]; U[r * 3 + j] = t_;     \
            t_ = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + j]; V[r * 3 + j] = t_;     \
        }                                                                          \
    } while (0)
    {
        int pos = 0;
        if (S[1] > S[pos]) pos = 1;
        if (S[2] > S[pos]) pos = 2;
        if (S[pos] == 0.0f) return true;
        if (pos == 1) MPM_SWAPCOL(0, 1); else if (pos == 2) MPM_SWAPCOL(0, 2);
        if (S[2] > S[1]) { MPM_SWAPCOL(1, 2); }
    }
#undef MPM_SWAPCOL
    return true;
}
// s = dinv * dt. Outputs as f_update_rn.
MPM_DI bool f_update_fast(const float (&B)[9], float (&FE)[9], float (&FP)[9], float s, float clamp_lo, float clamp_hi,
                          float (&Ug)[9], float (&Sg)[3]) {
    float T0[9], Fh[9], T[9], Vg[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) T0[i] = fmaf(B[i], s, (i % 4 == 0) ? 1.0f : 0.0f);
    m3_mul_fast(Fh, T0, FE);
    m3_mul_fast(T, Fh, FP);
    if (!svd3_fast(Fh, Ug, Sg, Vg)) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) Sg[k] = fminf(fmaxf(Sg[k], clamp_lo), clamp_hi);
    float US[9], Vt[9], FEinv[9];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) US[c * 3 + r] = Ug[c * 3 + r] * Sg[c];
    m3_transpose(Vt, Vg);
    m3_mul_fast(FE, US, Vt);
    m3_inverse_fast(FEinv, FE);
    m3_mul_fast(FP, FEinv, T);
    return true;
}

// ---- stress term of computeExplicitGridForces (cpp:235-252) as one symmetric matrix ----
// M = V0 * Dinv * dPsi * FE^T = A diag(d_k) A^T with FE = A diag(sig) B^T,
// d_k = V0*dinv*(2 mu sig_k (sig_k - 1) + lambda (J - 1) J);  tau = (xx, yy, zz, xy, xz, yz)
MPM_DI void lame(float detFP, float E, float nu, float xi, float& mu, float& lambda) {
    const float e = expf(xi * (1.0f - detFP));
    mu = E / (2.0f * (1.0f + nu)) * e;
    lambda = (E * nu) / ((1.0f + nu) * (1.0f - 2.0f * nu)) * e;
}
MPM_DI void tau_from_factors(const float (&Ug)[9], const float (&Sg)[3], float J, float detFP, float V0, float dinv,
                             float E, float nu, float xi, float (&tau)[6]) {
    float mu, lambda;
    lame(detFP, E, nu, xi, mu, lambda);
    const float vol = V0 * dinv;
    const float iso = lambda * (J - 1.0f) * J;
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = vol * (2.0f * mu * Sg[k] * (Sg[k] - 1.0f) + iso);
    // A(r,k) = Ug[k*3+r]
#define MPM_SYM(r, s) (Ug[0 + r] * d[0] * Ug[0 + s] + Ug[3 + r] * d[1] * Ug[3 + s] + Ug[6 + r] * d[2] * Ug[6 + s])
    tau[0] = MPM_SYM(0, 0); tau[1] = MPM_SYM(1, 1); tau[2] = MPM_SYM(2, 2);
    tau[3] = MPM_SYM(0, 1); tau[4] = MPM_SYM(0, 2); tau[5] = MPM_SYM(1, 2);
#undef MPM_SYM
}
// tolerance form for the fused path: Lame parameters of the undeformed material precomputed on the host (mu0, lambda0:
// the same IEEE divisions), exp through one ex2.approx
MPM_DI void tau_from_factors_fast(const float (&Ug)[9], const float (&Sg)[3], float J, float detFP, float V0, float dinv,
                                  float mu0, float lambda0, float xi, float (&tau)[6]) {
    const float e = fast_ex2(xi * (1.0f - detFP) * 1.4426950408889634f);
    const float vol = V0 * dinv, two_mu = 2.0f * mu0 * e;
    const float iso = lambda0 * e * (J - 1.0f) * J;
    float Ud[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dk = vol * fmaf(two_mu * Sg[k], Sg[k] - 1.0f, iso);
        Ud[k * 3 + 0] = Ug[k * 3 + 0] * dk; Ud[k * 3 + 1] = Ug[k * 3 + 1] * dk; Ud[k * 3 + 2] = Ug[k * 3 + 2] * dk;
    }
#define MPM_SYMF(r, s) fmaf(Ud[6 + r], Ug[6 + s], fmaf(Ud[3 + r], Ug[3 + s], Ud[0 + r] * Ug[0 + s]))
    tau[0] = MPM_SYMF(0, 0); tau[1] = MPM_SYMF(1, 1); tau[2] = MPM_SYMF(2, 2);
    tau[3] = MPM_SYMF(0, 1); tau[4] = MPM_SYMF(0, 2); tau[5] = MPM_SYMF(1, 2);
#undef MPM_SYMF
}
// general FE (after an upload): rotation by Newton iteration R <- (R + R^-T)/2, quadratically convergent
MPM_DI void tau_general(const float (&FE)[9], float detFP, float V0, float dinv, float E, float nu, float xi,
                        float (&tau)[6]) {
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = FE[i];
    for (int it = 0; it < 40; ++it) {
        const float c00 = MG(R,1,1) * MG(R,2,2) - MG(R,2,1) * MG(R,1,2);
        const float c10 = MG(R,2,0) * MG(R,1,2) - MG(R,1,0) * MG(R,2,2);
        const float c20 = MG(R,1,0) * MG(R,2,1) - MG(R,2,0) * MG(R,1,1);
        const float det = MG(R,0,0) * c00 + MG(R,0,1) * c10 + MG(R,0,2) * c20;
        if (det == 0.0f || !isfinite(det)) break;
        const float id = 1.0f / det;
        // cofactor matrix C(c,r) (glm indexing) so that R^-T = C / det
        float C[9];
        MG(C,0,0) = c00; MG(C,0,1) = c10; MG(C,0,2) = c20;
        MG(C,1,0) = MG(R,2,1) * MG(R,0,2) - MG(R,0,1) * MG(R,2,2);
        MG(C,1,1) = MG(R,0,0) * MG(R,2,2) - MG(R,2,0) * MG(R,0,2);
        MG(C,1,2) = MG(R,2,0) * MG(R,0,1) - MG(R,0,0) * MG(R,2,1);
        MG(C,2,0) = MG(R,0,1) * MG(R,1,2) - MG(R,1,1) * MG(R,0,2);
        MG(C,2,1) = MG(R,1,0) * MG(R,0,2) - MG(R,0,0) * MG(R,1,2);
        MG(C,2,2) = MG(R,0,0) * MG(R,1,1) - MG(R,1,0) * MG(R,0,1);
        float diff = 0.0f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float v = 0.5f * (R[i] + C[i] * id);
            diff = fmaxf(diff, fabsf(v - R[i]));
            R[i] = v;
        }
        if (diff < 2e-7f) break;
    }
    float mu, lambda;
    lame(detFP, E, nu, xi, mu, lambda);
    const float J = MG(FE,0,0) * (MG(FE,1,1) * MG(FE,2,2) - MG(FE,2,1) * MG(FE,1,2))
                  - MG(FE,1,0) * (MG(FE,0,1) * MG(FE,2,2) - MG(FE,2,1) * MG(FE,0,2))
                  + MG(FE,2,0) * (MG(FE,0,1) * MG(FE,1,2) - MG(FE,1,1) * MG(FE,0,2));
    const float vol = V0 * dinv, iso = lambda * (J - 1.0f) * J;
    // M(r,s) = vol * (2 mu * sum_c (FE - R)(r,c) FE(s,c) + iso * delta_rs), symmetrised
    float M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += (MG(FE,c,r) - MG(R,c,r)) * MG(FE,c,s);
            M[r * 3 + s] = vol * (2.0f * mu * acc + (r == s ? iso : 0.0f));
        }
    tau[0] = M[0]; tau[1] = M[4]; tau[2] = M[8];
    tau[3] = 0.5f * (M[1] + M[3]); tau[4] = 0.5f * (M[2] + M[6]); tau[5] = 0.5f * (M[5] + M[7]);
}

// ---- box collider, hpp:79-86 + mathy.hpp:40-56 + cpp:264-296 (bit-faithful) ----
struct BoxCollider { float w2l[16]; float half[3]; float vel[3]; };

MPM_DI float box_sdf_rn(const BoxCollider& c, float px, float py, float pz) {
    float p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        p[r] = add_rn(add_rn(mul_rn(c.w2l[0 + r], px), mul_rn(c.w2l[4 + r], py)),
                      add_rn(mul_rn(c.w2l[8 + r], pz), mul_rn(c.w2l[12 + r], 1.0f)));
    if (c.half[1] < 0.0f)           // sphere (mpm_sphere_collider): |p_local| - radius, every operation rounded once like the oracle's
        return sub_rn(__fsqrt_rn(add_rn(add_rn(mul_rn(p[0], p[0]), mul_rn(p[1], p[1])), mul_rn(p[2], p[2]))), c.half[0]);
    const float qx = sub_rn(fabsf(p[0]), c.half[0]), qy = sub_rn(fabsf(p[1]), c.half[1]), qz = sub_rn(fabsf(p[2]), c.half[2]);
    float mx = qx; if (mx < qy) mx = qy; if (mx < qz) mx = qz; if (mx < 0.0f) mx = 0.0f;
    float in = qy < qz ? qz : qy; in = qx < in ? in : qx; in = 0.0f < in ? 0.0f : in;
    return add_rn(fabsf(mx), in);
}

MPM_DI void body_collision_rn(const BoxCollider* cs, int nc, float friction, float px, float py, float pz, float (&v)[3]) {
    bool all_out = true;
    for (int k = 0; k < nc; ++k) if (!(box_sdf_rn(cs[k], px, py, pz) > 0.0f)) { all_out = false; break; }
    if (all_out) return;
    const float delta = 0.001f;
    const float two_delta = mul_rn(2.0f, delta);
    for (int k = 0; k < nc; ++k) {
        if (box_sdf_rn(cs[k], px, py, pz) > 0.0f) continue;
        float n[3];
        n[0] = div_rn(sub_rn(box_sdf_rn(cs[k], add_rn(px, delta), py, pz), box_sdf_rn(cs[k], sub_rn(px, delta), py, pz)), two_delta);
        n[1] = div_rn(sub_rn(box_sdf_rn(cs[k], px, add_rn(py, delta), pz), box_sdf_rn(cs[k], px, sub_rn(py, delta), pz)), two_delta);
        n[2] = div_rn(sub_rn(box_sdf_rn(cs[k], px, py, add_rn(pz, delta)), box_sdf_rn(cs[k], px, py, sub_rn(pz, delta))), two_delta);
        float rel[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) rel[a] = sub_rn(v[a], cs[k].vel[a]);
        const float vn = dot3_rn(rel[0], n[0], rel[1], n[1], rel[2], n[2]);
        if (vn >= 0.0f) continue;
        float vrel[3] = { 0.0f, 0.0f, 0.0f };
        // cpp:290-291: vt.length() is glm's component count (3), not the norm
        if (3.0f > mul_rn(-friction, vn)) {
            const float s = div_rn(mul_rn(friction, vn), 3.0f);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float vt = sub_rn(rel[a], mul_rn(n[a], vn));
                vrel[a] = add_rn(vt, mul_rn(vt, s));
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) v[a] = add_rn(vrel[a], cs[k].vel[a]);
    }
}

}  // namespace mpm

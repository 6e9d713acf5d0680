// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the reference's known-answer vectors (tests/golden/kat_functions.npz, made by
// the unmodified reference) pushed through the DEVICE code compiled for the host. The bit-faithful device routines use
// only IEEE-rounded operations (_rn intrinsics, sqrt, fma), so the host build must reproduce the reference bit for bit
// exactly as the GPU does: weight polynomial, bodyCollision with static and moving colliders (function and
// k_grid_update kernel), updateDeformationGradient (k_fupdate kernel).
// usage: emu_kat <dir with *.f32 files written by tests/test_kernel_emulation_cpu.py>
#include "mpm_tile_kernels.cuh"

#include <string>

using namespace mpm;

static std::vector<float> load(const std::string& dir, const char* name) {
    const std::string path = dir + "/" + name + ".f32";
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<float> v(bytes / 4);
    if (std::fread(v.data(), 4, v.size(), f) != v.size()) { std::fprintf(stderr, "short read %s\n", path.c_str()); std::exit(2); }
    std::fclose(f);
    return v;
}
static int failures = 0;
static void check(bool ok, const char* what) { std::printf("%s  %s\n", ok ? "ok  " : "FAIL", what); if (!ok) ++failures; }
static bool same(float a, float b) { return std::memcmp(&a, &b, 4) == 0 || (a == 0.0f && b == 0.0f); }

static ColliderSet colliders_from(const std::vector<float>& raw, int& nc) {     // rows: w2l[16], half[3], vel[3]
    ColliderSet cs;
    std::memset(&cs, 0, sizeof cs);
    nc = (int)(raw.size() / 22);
    for (int k = 0; k < nc; ++k) {
        std::memcpy(cs.c[k].w2l, &raw[k * 22], 64);
        std::memcpy(cs.c[k].half, &raw[k * 22 + 16], 12);
        std::memcpy(cs.c[k].vel, &raw[k * 22 + 19], 12);
    }
    return cs;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const float h = 0.05f, friction = 0.5f, dt = 1e-5f;

    {   // weightNx (hpp:20-31)
        const std::vector<float> x = load(dir, "weights_x"), w = load(dir, "weights_w");
        size_t bad = 0;
        for (size_t i = 0; i < x.size(); ++i) bad += !same(weight_nx_exact(x[i]), w[i]);
        check(bad == 0 && !x.empty(), "weight_nx_exact == reference weightNx, bit for bit");
        double worst = 0;
        for (size_t i = 0; i < x.size(); ++i) worst = std::max(worst, (double)std::fabs(weight_nx(x[i]) - w[i]));
        check(worst <= 1.2e-7, "hot-path fp32 weight polynomial within 1 ulp of the reference");
    }
    const std::vector<float> pos = load(dir, "collide_pos"), vel = load(dir, "collide_vel");
    const int rows = (int)(pos.size() / 3);
    for (int moving = 0; moving < 2; ++moving) {   // bodyCollision (cpp:264-296) as a function, every KAT row
        int nc = 0;
        const ColliderSet cs = colliders_from(load(dir, moving ? "colliders_moving" : "colliders"), nc);
        const std::vector<float> want = load(dir, moving ? "collide_moving_out" : "collide_out");
        size_t bad = 0;
        for (int r = 0; r < rows; ++r) {
            float v[3] = { vel[3 * r], vel[3 * r + 1], vel[3 * r + 2] };
            body_collision_rn(cs.c, nc, friction, pos[3 * r], pos[3 * r + 1], pos[3 * r + 2], v);
            for (int c = 0; c < 3; ++c) bad += !same(v[c], want[3 * r + c]);
        }
        check(bad == 0 && rows > 0, moving ? "body_collision_rn == reference bodyCollision, moving colliders, bit for bit"
                                           : "body_collision_rn == reference bodyCollision, static colliders, bit for bit");
    }
    {   // the same through the grid kernel on the 20^3 grid (rows 0..7999 are the node positions idx*h in i,j,k order)
        GridDims gd{};
        gd.I = gd.J = gd.K = 20;
        gd.npbi_global = gd.npbj = gd.npbk = 5; gd.nbj = gd.nbk = 6; gd.lo = 0; gd.hi = 5;
        gd.n_pblocks = 125; gd.n_gblocks = 6 * 6 * 6;
        SimConst sc{};
        sc.h = h; sc.friction = friction;
        int nc = 0;
        const ColliderSet cs = colliders_from(load(dir, "colliders_moving"), nc);
        const std::vector<float> want = load(dir, "collide_moving_out");
        std::vector<float4> grid((size_t)gd.n_gblocks * 64, make_float4(0, 0, 0, 0));
        for (int i = 0; i < 20; ++i) for (int j = 0; j < 20; ++j) for (int k = 0; k < 20; ++k) {
            const int r = (i * 20 + j) * 20 + k;
            grid[node_index(gd, i, j, k)] = make_float4(1.0f, vel[3 * r], vel[3 * r + 1], vel[3 * r + 2]);
        }
        std::vector<int> blocks(gd.n_gblocks);
        for (int b = 0; b < gd.n_gblocks; ++b) blocks[b] = gd.n_gblocks - 1 - b;
        DevCounters dc{};
        dc.n_active_gblocks = gd.n_gblocks;
        emu::launch(3, 256, 0, [&] { k_grid_update<GU_COLLIDE>(blocks.data(), &dc, grid.data(), nullptr, gd, sc, dt, cs, nc); });
        size_t bad = 0;
        for (int i = 0; i < 20; ++i) for (int j = 0; j < 20; ++j) for (int k = 0; k < 20; ++k) {
            const int r = (i * 20 + j) * 20 + k;
            const float4 n = grid[node_index(gd, i, j, k)];
            bad += !same(n.y, want[3 * r]) + !same(n.z, want[3 * r + 1]) + !same(n.w, want[3 * r + 2]);
        }
        check(bad == 0, "k_grid_update<COLLIDE> on the 20^3 grid == reference gridBasedCollisions (moving colliders), bit for bit");
    }
    for (int pk = 0; pk < 2; ++pk) {   // updateDeformationGradient (cpp:306-330) through k_fupdate; rows: FE[9] FP[9] B[9] -> FE[9] FP[9]
        const std::vector<float> in = load(dir, "fupdate_in"), want = load(dir, "fupdate_out");
        const int n = (int)(in.size() / 27), cap = n + 8;
        std::vector<float4> buf((size_t)NPLANES * cap, make_float4(0, 0, 0, 0));
        Planes P;
        for (int k = 0; k < NPLANES; ++k) P.p[k] = buf.data() + (size_t)k * cap;
        std::vector<int> ids(cap);
        for (int p = 0; p < n; ++p) {
            const float* FE = &in[27 * p]; const float* FP = FE + 9; const float* B = FE + 18;
            ids[p] = p;
            P.p[0][p] = make_float4(0.5f, 0.5f, 0.5f, 6e-5f);
            P.p[1][p] = make_float4(B[0], B[1], B[2], B[3]); P.p[2][p] = make_float4(B[4], B[5], B[6], B[7]); P.p[3][p] = make_float4(B[8], 0, 0, 0);
            P.p[6][p] = make_float4(3e-5f, __int_as_float(p), FE[0], FE[1]);
            P.p[7][p] = make_float4(FE[2], FE[3], FE[4], FE[5]); P.p[8][p] = make_float4(FE[6], FE[7], FE[8], FP[0]);
            P.p[9][p] = make_float4(FP[1], FP[2], FP[3], FP[4]); P.p[10][p] = make_float4(FP[5], FP[6], FP[7], FP[8]);
        }
        SimConst sc{};
        sc.h = h; sc.dinv = 1.0f / ((1.0f / 3.0f) * h * h);
        {   // DpInverse exactly as the reference forms it (hpp:177): glm::inverse(mat3(1) * (1.0f/3.0f) * h * h), entry [0][0]
            volatile float d = (1.0f / 3.0f); d = d * h; d = d * h;
            volatile float dd = d * d;
            volatile float det = d * dd;
            volatile float ood = 1.0f / det;
            volatile float r = dd * ood;
            sc.dinv = r;
        }
        sc.E = 1.4e5f; sc.nu = 0.2f; sc.xi = 10.f; sc.mu0 = sc.E / (2.0f * (1.0f + sc.nu)); sc.lambda0 = (sc.E * sc.nu) / ((1.0f + sc.nu) * (1.0f - 2.0f * sc.nu)); sc.clamp_lo = (float)(1.0 - 2.5e-2); sc.clamp_hi = (float)(1.0 + 5e-3);
        DevCounters dc{};
        dc.n_binned = n; dc.n_sorted = n; dc.n_slots = n;
        if (pk) emu::launch((n + 255) / 256, 256, 0, [&] { k_fupdate<false, true>(P, P, ids.data(), &dc, sc, dt); });
        else emu::launch((n + 255) / 256, 256, 0, [&] { k_fupdate<false>(P, P, ids.data(), &dc, sc, dt); });
        size_t bad = 0, far = 0, finite = 0, prod_bad = 0;
        for (int p = 0; p < n; ++p) {
            const float4 a6 = P.p[6][p], a7 = P.p[7][p], a8 = P.p[8][p], a9 = P.p[9][p], a10 = P.p[10][p];
            const float got[18] = { a6.z, a6.w, a7.x, a7.y, a7.z, a7.w, a8.x, a8.y, a8.z, a8.w, a9.x, a9.y, a9.z, a9.w, a10.x, a10.y, a10.z, a10.w };
            bool fin = true; float dmax = 0.0f;
            for (int c = 0; c < 18; ++c) {
                const float w = want[18 * p + c];
                const bool ok = same(got[c], w) || (std::isnan(got[c]) && std::isnan(w));
                bad += !ok;
                fin = fin && std::isfinite(w) && std::isfinite(got[c]);
                dmax = std::fmax(dmax, std::fabs(got[c] - w));
            }
            if (fin) {
                ++finite; far += dmax > 1e-4f;
                // FE_new * FP_new must reproduce the reference's FE_new * FP_new (= (I + dt C) FE FP: the total deformation
                // gradient does not depend on how the SVD splits it), glm column-major 3x3 products
                for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
                    double a = 0, b = 0;
                    for (int k = 0; k < 3; ++k) { a += (double)got[k * 3 + r] * got[9 + c * 3 + k]; b += (double)want[18 * p + k * 3 + r] * want[18 * p + 9 + c * 3 + k]; }
                    prod_bad += std::fabs(a - b) > 2e-5 * (1.0 + std::fabs(b));
                }
            }
        }
        if (pk) {
            std::printf("      tolerance-form F-update: %zu of %zu finite KAT rows differ by more than 1e-4 in FE/FP (singular-value order flips on near-degenerate inputs), %zu product entries off\n", far, finite, prod_bad);
            // calibration: the reference's own arithmetic rebuilt with FMA contraction (libmpm_oracle_fma) moves 366 of these 4096 rows by
            // more than 1e-4 (near-isotropic FE: rounding decides the order of the singular values, cpp:306-330 transposes the factors)
            check(n > 0 && prod_bad == 0 && far * 100 <= finite * 12, "k_fupdate<tolerance form>: FE*FP == reference's to 2e-5 on every finite KAT row; FE/FP beyond 1e-4 on <= 12 % of rows (reference vs reference+FMA: 8.9 %)");
        } else
        check(bad == 0 && n > 0, "k_fupdate == reference updateDeformationGradient (Eigen-convention Jacobi SVD), bit for bit");
    }
    std::printf("%s (%d failures)\n", failures ? "EMULATED KATS FAILED" : "all emulated known-answer tests passed", failures);
    return failures ? 1 : 0;
}

// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): runs the tile kernels of csrc/ on the host, compiled by g++ from the very
// same sources, and checks the kernel VARIANTS against each other:
//   P2G    : block-tile kernel (default) vs the per-particle-atomics baseline vs the packed-pair variant vs the variants
//            that also run the F-update; the ids of every block must stay a permutation of the block's ids
//   F-upd  : planes 4..10 written by P2G's F-update phase == k_fupdate<true> with the same ids, bit for bit
//   gather : blocked tile (default) vs linear tile (bit for bit) vs packed pairs vs the direct-gather baseline
// The default kernels are the ones validated on hardware against the oracle; this harness shows that a variant's index
// arithmetic, barrier structure and bulk-copy byte accounting are consistent with them before any GPU time is spent.
// Build/run: tests/test_kernel_emulation_cpu.py   (g++ -O1 -ffp-contract=off -include cuda_emu.h)
#include "mpm_tile_kernels.cuh"

#include <random>

using namespace mpm;

struct Host {
    GridDims gd{};
    SimConst sc{};
    int n = 0, cap = 0;
    std::vector<float4> buf[2];
    std::vector<int> ids0;              // binned ids (before any P2G re-ordering)
    std::vector<int4> work;
    std::vector<float4> grid;
    DevCounters dc{};
    Planes planes(int b) { Planes P; for (int k = 0; k < NPLANES; ++k) P.p[k] = buf[b].data() + (size_t)k * cap; return P; }
};

static void host_bin(Host& H, int bufi);
static float frand(std::mt19937& g, float lo, float hi) { return lo + (hi - lo) * (float)(g() >> 8) / 16777216.0f; }

static void put_particle(Host& H, int p, std::mt19937& g, float x, float y, float z) {
    Planes P = H.planes(0);
    float B[9], FE[9], FP[9], tau[6];
    for (int i = 0; i < 9; ++i) {
        B[i] = frand(g, -1e-3f, 1e-3f);
        FE[i] = ((i % 4 == 0) ? 1.0f : 0.0f) + frand(g, -0.05f, 0.05f);
        FP[i] = ((i % 4 == 0) ? 1.0f : 0.0f) + frand(g, -0.02f, 0.02f);
    }
    for (int i = 0; i < 6; ++i) tau[i] = frand(g, -2e-3f, 2e-3f);
    const float v[3] = { frand(g, -5.f, 5.f), frand(g, -5.f, 5.f), frand(g, -5.f, 5.f) };
    P.p[0][p] = make_float4(x, y, z, 6e-5f * frand(g, 0.5f, 1.5f));
    P.p[1][p] = make_float4(B[0], B[1], B[2], B[3]);
    P.p[2][p] = make_float4(B[4], B[5], B[6], B[7]);
    P.p[3][p] = make_float4(B[8], v[0], v[1], v[2]);
    P.p[4][p] = make_float4(tau[0], tau[1], tau[2], tau[3]);
    P.p[5][p] = make_float4(tau[4], tau[5], 0.f, 0.f);
    P.p[6][p] = make_float4(1.2e-5f * frand(g, 0.8f, 1.2f), __int_as_float(p), FE[0], FE[1]);
    P.p[7][p] = make_float4(FE[2], FE[3], FE[4], FE[5]);
    P.p[8][p] = make_float4(FE[6], FE[7], FE[8], FP[0]);
    P.p[9][p] = make_float4(FP[1], FP[2], FP[3], FP[4]);
    P.p[10][p] = make_float4(FP[5], FP[6], FP[7], FP[8]);
}

// scene: sparse random fill + blocks holding exactly 1 / 512 / 513 / 1500 particles (single, full, two and three chunks)
static void build(Host& H, int dim, unsigned seed, bool fast_div) {
    std::mt19937 g(seed);
    GridDims& gd = H.gd;
    gd.I = gd.J = gd.K = dim;
    gd.npbi_global = gd.npbj = gd.npbk = (dim + 3) / 4;
    gd.nbj = gd.npbj + 1; gd.nbk = gd.npbk + 1;
    gd.lo = 0; gd.hi = gd.npbi_global;
    gd.n_pblocks = (gd.hi - gd.lo) * gd.npbj * gd.npbk;
    gd.n_gblocks = (gd.hi - gd.lo + 1) * gd.nbj * gd.nbk;
    SimConst& c = H.sc;
    const float h = 0.05f;
    c.h = h; c.dinv = 1.0f / ((1.0f / 3.0f) * h * h); c.E = 1.4e5f; c.nu = 0.2f; c.xi = 10.f; c.mu0 = c.E / (2.0f * (1.0f + c.nu)); c.lambda0 = (c.E * c.nu) / ((1.0f + c.nu) * (1.0f - 2.0f * c.nu));
    c.clamp_lo = (float)(1.0 - 2.5e-2); c.clamp_hi = (float)(1.0 + 5e-3); c.friction = 0.5f;
    c.g[0] = 0; c.g[1] = -9.8f; c.g[2] = 0;
    c.pos_lo = 3 * h; c.pos_hi[0] = c.pos_hi[1] = c.pos_hi[2] = (dim - 3) * h;
    c.p2g_rotate = 1;
    c.pd.h = h; c.pd.rh = 1.0f / h; c.pd.lo = 2.0f * h; c.pd.hi = (float)(dim + 2) * h; c.pd.fast = fast_div ? 1 : 0;

    const int n_sparse = 900;
    struct Crowd { int bi, bj, bk, n; };
    const Crowd crowds[] = { { 1, 1, 1, 1 }, { 2, 1, 3, 512 }, { 3, 3, 2, 513 }, { 2, 3, 1, 1500 }, { 4, 2, 2, 31 } };
    int n = n_sparse;
    for (const Crowd& cr : crowds) n += cr.n;
    H.n = n; H.cap = n + 64;
    for (int b = 0; b < 2; ++b) H.buf[b].assign((size_t)NPLANES * H.cap, make_float4(NAN, NAN, NAN, NAN));
    int p = 0;
    const float lo = 3.0f * h, hi = (dim - 3) * h;
    for (int i = 0; i < n_sparse; ++i) put_particle(H, p++, g, frand(g, lo, hi), frand(g, lo, hi), frand(g, lo, hi));
    for (const Crowd& cr : crowds)          // particle block b holds cells 4b+1 .. 4b+4
        for (int i = 0; i < cr.n; ++i)
            put_particle(H, p++, g, (4 * cr.bi + 1 + frand(g, 0.01f, 3.99f)) * h, (4 * cr.bj + 1 + frand(g, 0.01f, 3.99f)) * h,
                         (4 * cr.bk + 1 + frand(g, 0.01f, 3.99f)) * h);
    host_bin(H, 0);
}

// host binning with the kernels' own key function (stands in for k_bin_count / scan / k_bin_scatter)
static void host_bin(Host& H, int bufi) {
    const GridDims& gd = H.gd;
    const SimConst& c = H.sc;
    const int n = H.n;
    std::vector<int> key(n), count(gd.n_pblocks + 3, 0), start(gd.n_pblocks + 4, 0);
    Planes P = H.planes(bufi);
    for (int q = 0; q < n; ++q) {
        int cells[3];
        key[q] = particle_key(P.p[0][q], gd, c.pd, cells);
        if (key[q] < 0 || key[q] >= gd.n_pblocks) { std::fprintf(stderr, "harness scene: particle %d is not in a real block\n", q); std::exit(2); }
        count[key[q]]++;
    }
    for (int b = 0; b < gd.n_pblocks + 3; ++b) start[b + 1] = start[b] + count[b];
    H.ids0.assign(H.cap, -1);
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int q = 0; q < n; ++q) H.ids0[cur[key[q]]++] = q;
    H.work.clear();
    for (int b = gd.n_pblocks - 1; b >= 0; --b)            // any order is legal; use a different one than the device scan
        if (count[b]) H.work.push_back(make_int4(b, start[b], count[b], pack_block_coords(b / (gd.npbk * gd.npbj), (b / gd.npbk) % gd.npbj, b % gd.npbk)));
    H.grid.assign((size_t)gd.n_gblocks * 64, make_float4(0, 0, 0, 0));
    std::memset(&H.dc, 0, sizeof H.dc);
    H.dc.n_slots = n; H.dc.n_binned = n; H.dc.n_sorted = n; H.dc.n_active_pblocks = (int)H.work.size();
}

static int failures = 0;
static void check(bool ok, const char* what) {
    std::printf("%s  %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) ++failures;
}
// max |a-b| over the four channels relative to the largest magnitude of that channel
static double grid_rel_diff(const std::vector<float4>& a, const std::vector<float4>& b) {
    double mx[4] = { 0, 0, 0, 0 }, d[4] = { 0, 0, 0, 0 };
    for (size_t i = 0; i < a.size(); ++i) {
        const float* x = &a[i].x; const float* y = &b[i].x;
        for (int c = 0; c < 4; ++c) { mx[c] = std::max(mx[c], (double)std::fabs(x[c])); d[c] = std::max(d[c], (double)std::fabs(x[c] - y[c])); }
    }
    double r = 0;
    for (int c = 0; c < 4; ++c) r = std::max(r, d[c] / std::max(mx[c], 1e-30));
    return r;
}

template <int FUPD>
static void run_p2g(Host& H, std::vector<int>& ids, std::vector<float4>& grid, float dt) {
    ids = H.ids0;
    grid.assign(H.grid.size(), make_float4(0, 0, 0, 0));
    H.dc.work_a = 0;
    Planes P = H.planes(0), N = H.planes(1);
    emu::launch(3, P2G_T, sizeof(P2GSmem), [&] {
        k_p2g_tile<P2G_FUSED, FUPD>(P, ids.data(), H.work.data(), &H.dc, grid.data(), H.gd, H.sc, dt, FUPD ? N : P);
    });
}
static bool ids_are_block_permutations(const Host& H, const std::vector<int>& ids) {
    for (const int4& w : H.work) {
        std::vector<int> a(H.ids0.begin() + w.y, H.ids0.begin() + w.y + w.z), b(ids.begin() + w.y, ids.begin() + w.y + w.z);
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) return false;
    }
    return true;
}
static void clear_planes(Host& H, int b, int k0, int k1) {
    for (int k = k0; k <= k1; ++k) std::fill(H.buf[b].begin() + (size_t)k * H.cap, H.buf[b].begin() + (size_t)(k + 1) * H.cap, make_float4(NAN, NAN, NAN, NAN));
}
static bool planes_bit_equal(const std::vector<float4>& a, const std::vector<float4>& b, int cap, int n, int k0, int k1, bool allow_zero_sign) {
    for (int k = k0; k <= k1; ++k)
        for (int j = 0; j < n; ++j) {
            const float* x = &a[(size_t)k * cap + j].x; const float* y = &b[(size_t)k * cap + j].x;
            for (int c = 0; c < 4; ++c) {
                if (std::memcmp(&x[c], &y[c], 4) == 0) continue;
                if (allow_zero_sign && x[c] == 0.0f && y[c] == 0.0f) continue;
                std::fprintf(stderr, "  plane %d slot %d comp %d: %.9g vs %.9g\n", k, j, c, x[c], y[c]);
                return false;
            }
        }
    return true;
}
static double planes_rel_diff(const std::vector<float4>& a, const std::vector<float4>& b, int cap, int n, int k0, int k1) {
    double r = 0;
    for (int k = k0; k <= k1; ++k) {
        double mx = 0, d = 0;
        for (int j = 0; j < n; ++j) {
            const float* x = &a[(size_t)k * cap + j].x; const float* y = &b[(size_t)k * cap + j].x;
            for (int c = 0; c < 4; ++c) { mx = std::max(mx, (double)std::fabs(x[c])); d = std::max(d, (double)std::fabs(x[c] - y[c])); }
        }
        r = std::max(r, d / std::max(mx, 1e-30));
    }
    return r;
}

static std::vector<float4> run_gather(Host& H, const std::vector<int>& ids, const std::vector<float4>& vel, float dt) {
    clear_planes(H, 1, 0, 3);
    H.dc.work_b = 0;
    Planes C = H.planes(0), N = H.planes(1);
    const long long before = emu::tma_bytes_total();
    emu::launch(2, G2P_T, sizeof(G2PSmem), [&] {
        k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER>(C, N, ids.data(), H.work.data(), &H.dc, vel.data(), H.gd, H.sc, dt);
    });
    const long long moved = emu::tma_bytes_total() - before;
    check(moved == (long long)H.work.size() * 8192, "gather tile: 8 KB of bulk copies per particle block");
    return H.buf[1];
}

// ---------------------------------------------------------------------------------------------------------------
// whole fused substeps, kernel sequence of mpm_substep_begin/end (binning done on the host):
//   clear -> P2G<FUSED> -> grid update -> [F-update] -> gather/advect/re-sort -> swap
// with the default kernels and with every experimental option switched on (packed P2G that also runs the F-update,
// gather on the linear tile with packed pairs). The particle state after several substeps must agree: this covers the
// data flow BETWEEN substeps (tau and FE/FP written by P2G's F-update phase into the other buffer, consumed after the swap).
// ---------------------------------------------------------------------------------------------------------------
static void build_flow_scene(Host& H, unsigned seed) {
    build(H, 24, seed, true);                       // dims and constants; particles replaced below
    std::mt19937 g(seed * 7919u + 1u);
    const float h = H.sc.h;
    const int n = 8 * 512;
    H.n = n; H.cap = n + 64;
    for (int b = 0; b < 2; ++b) H.buf[b].assign((size_t)NPLANES * H.cap, make_float4(NAN, NAN, NAN, NAN));
    Planes P = H.planes(0);
    int p = 0;
    for (int ci = 9; ci < 17; ++ci) for (int cj = 5; cj < 13; ++cj) for (int ck = 9; ck < 17; ++ck)     // 8^3 cells x 8 particles
        for (int s8 = 0; s8 < 8; ++s8, ++p) {
            const float x = (ci + 0.25f + 0.5f * (s8 & 1) + frand(g, -0.2f, 0.2f)) * h, y = (cj + 0.25f + 0.5f * ((s8 >> 1) & 1) + frand(g, -0.2f, 0.2f)) * h,
                        z = (ck + 0.25f + 0.5f * (s8 >> 2) + frand(g, -0.2f, 0.2f)) * h;
            P.p[0][p] = make_float4(x, y, z, 6e-5f);
            P.p[1][p] = make_float4(0, 0, 0, 0); P.p[2][p] = make_float4(0, 0, 0, 0);
            P.p[3][p] = make_float4(0, frand(g, -0.5f, 0.5f), -40.0f + frand(g, -0.5f, 0.5f), frand(g, -0.5f, 0.5f));
            P.p[4][p] = make_float4(0, 0, 0, 0); P.p[5][p] = make_float4(0, 0, 0, 0);
            // compressed, stress from the first substep on; singular values well separated (the reference re-assembles FE
            // from transposed SVD factors, so the axis a clamped value lands on depends on the ORDER of the singular
            // values: with nearly equal ones fp32 noise flips it -- that chaos is the reference's, not a kernel property)
            const float c0 = 0.993f, c1 = 0.9755f, c2 = 0.986f;
            P.p[6][p] = make_float4(1.5e-5f, __int_as_float(p), c0, 0);
            P.p[7][p] = make_float4(0, 0, c1, 0); P.p[8][p] = make_float4(0, 0, c2, 1);
            P.p[9][p] = make_float4(0, 0, 0, 1); P.p[10][p] = make_float4(0, 0, 0, 1);
        }
    H.dc.n_slots = n;
    emu::launch((n + 127) / 128, 128, 0, [&] { k_stress(P, &H.dc, H.sc); });
    H.grid.assign((size_t)H.gd.n_gblocks * 64, make_float4(0, 0, 0, 0));
}

template <int FU /* 0: k_fupdate, 1: bit-faithful F-update inside P2G, 2: tolerance form inside P2G */, bool BASE /* baseline kernels */>
static std::vector<float> run_flow(Host& H, int substeps, float dt, const ColliderSet& cs, int nc) {
    int cur = 0;
    std::vector<int> blocks(H.gd.n_gblocks);
    for (int b = 0; b < H.gd.n_gblocks; ++b) blocks[b] = b;
    for (int step = 0; step < substeps; ++step) {
        host_bin(H, cur);
        H.dc.n_active_gblocks = H.gd.n_gblocks;
        std::vector<int> ids = H.ids0;
        std::fill(H.grid.begin(), H.grid.end(), make_float4(0, 0, 0, 0));
        Planes C = H.planes(cur), N = H.planes(cur ^ 1);
        H.dc.work_a = 0; H.dc.work_b = 0;
        if (BASE) emu::launch((H.n + 127) / 128, 128, 0, [&] { k_p2g_atomic<P2G_FUSED>(C, ids.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt); });
        else emu::launch(3, P2G_T, sizeof(P2GSmem), [&] {
            k_p2g_tile<P2G_FUSED, FU>(C, ids.data(), H.work.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt, FU ? N : C);
        });
        emu::launch(3, 256, 0, [&] { k_grid_update<GU_NORMALIZE | GU_GRAVITY | GU_COLLIDE>(blocks.data(), &H.dc, H.grid.data(), nullptr, H.gd, H.sc, dt, cs, nc); });
        if (BASE) emu::launch((H.n + 127) / 128, 128, 0, [&] { k_g2p_direct<G2P_F | G2P_GATHER | G2P_ADVECT | G2P_REORDER>(C, N, ids.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt); });
        else {
            if (!FU) emu::launch((H.n + 255) / 256, 256, 0, [&] { k_fupdate<true>(C, N, ids.data(), &H.dc, H.sc, dt); });
            emu::launch(2, G2P_T, sizeof(G2PSmem), [&] {
                k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER>(C, N, ids.data(), H.work.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt);
            });
        }
        cur ^= 1;
    }
    // state by particle id: x y z | v | B | FE | FP | tau
    std::vector<float> out((size_t)H.n * 36, 0.0f);
    Planes P = H.planes(cur);
    for (int j = 0; j < H.n; ++j) {
        const int pid = __float_as_int(P.p[6][j].y);
        if (pid < 0 || pid >= H.n) { std::fprintf(stderr, "flow: slot %d carries particle id %d\n", j, pid); std::exit(3); }
        float* o = &out[(size_t)pid * 36];
        const float4 a0 = P.p[0][j], a1 = P.p[1][j], a2 = P.p[2][j], a3 = P.p[3][j], a4 = P.p[4][j], a5 = P.p[5][j], a6 = P.p[6][j], a7 = P.p[7][j],
                     a8 = P.p[8][j], a9 = P.p[9][j], a10 = P.p[10][j];
        const float v[36] = { a0.x, a0.y, a0.z, a3.y, a3.z, a3.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x,
                              a6.z, a6.w, a7.x, a7.y, a7.z, a7.w, a8.x, a8.y, a8.z, a8.w, a9.x, a9.y, a9.z, a9.w, a10.x, a10.y, a10.z, a10.w,
                              a4.x, a4.y, a5.y };
        std::memcpy(o, v, sizeof v);
    }
    return out;
}

static void check_flow(unsigned seed) {
    const float dt = 1e-5f;
    const int substeps = 4;
    ColliderSet cs;
    std::memset(&cs, 0, sizeof cs);
    // ground box: top face at y = 5.1 cells, i.e. under the blob's lowest particles (0.1 cell away, reached in the run)
    const float h = 0.05f;
    BoxCollider& b = cs.c[0];
    const float w2l[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -0.6f, -(5.1f * h - 1.0f), -0.6f, 1 };   // translate(-centre), centre y = top - half
    std::memcpy(b.w2l, w2l, sizeof w2l);
    b.half[0] = 5.0f; b.half[1] = 1.0f; b.half[2] = 5.0f;
    Host A, B;
    build_flow_scene(A, seed);
    build_flow_scene(B, seed);
    const std::vector<float> ra = run_flow<0, true>(A, substeps, dt, cs, 1);       // baseline kernels (per-particle atomics, direct gathers)
    const char* mode = std::getenv("EMU_FLOW_MODE");
    const int m = mode ? std::atoi(mode) : 1;
    std::vector<float> rb;
    switch (m) {
    case 0: rb = run_flow<0, false>(B, substeps, dt, cs, 1); break;      // tile kernels, F-update as its own kernel
    case 2: rb = run_flow<2, false>(B, substeps, dt, cs, 1); break;      // tile kernels, tolerance-form F-update inside P2G
    default: rb = run_flow<1, false>(B, substeps, dt, cs, 1); break;     // tile kernels, bit-faithful F-update inside P2G
    }
    const char* names[6] = { "position", "velocity", "B", "FE", "FP", "tau" };
    const int lo[6] = { 0, 3, 6, 15, 24, 33 }, hi[6] = { 3, 6, 15, 24, 33, 36 };
    bool ok = true, moved = false, stressed = false, plastic = false;
    for (int f = 0; f < 6; ++f) {
        double mx = 0, d = 0;
        for (int p = 0; p < A.n; ++p)
            for (int c = lo[f]; c < hi[f]; ++c) {
                const double x = ra[(size_t)p * 36 + c], y = rb[(size_t)p * 36 + c];
                if (!(std::isfinite(x) && std::isfinite(y))) ok = false;
                mx = std::max(mx, std::fabs(x)); d = std::max(d, std::fabs(x - y));
            }
        std::printf("      %d substeps, baseline kernels vs tile kernels (mode %d): %-8s max |d| %.3g (scale %.3g)\n", substeps, m, names[f], d, mx);
        if (!(d <= 2e-4 * mx)) ok = false;
        if (f == 5 && mx > 0) stressed = true;
    }
    for (int p = 0; p < A.n; ++p) {
        if (std::fabs(ra[(size_t)p * 36 + 4] + 40.0f) > 5.0f) moved = true;                 // some particle has hit the ground box
        if (std::fabs(ra[(size_t)p * 36 + 28] - 1.0f) > 1e-4f) plastic = true;             // FP_yy left 1: clamping happened
    }
    if (std::getenv("EMU_FLOW_DEBUG")) {
        int worst = 0; double wd = 0;
        for (int p = 0; p < A.n; ++p) for (int c = 15; c < 24; ++c) { const double d = std::fabs(ra[(size_t)p * 36 + c] - rb[(size_t)p * 36 + c]); if (d > wd) { wd = d; worst = p; } }
        std::printf("worst FE particle %d:\n", worst);
        for (int c = 0; c < 36; ++c) std::printf("  [%2d] %.9g  %.9g\n", c, ra[(size_t)worst * 36 + c], rb[(size_t)worst * 36 + c]);
        int cnt = 0;
        for (int p = 0; p < A.n; ++p) { double d = 0; for (int c = 15; c < 24; ++c) d = std::max(d, (double)std::fabs(ra[(size_t)p * 36 + c] - rb[(size_t)p * 36 + c])); cnt += d > 1e-5; }
        std::printf("particles with |dFE| > 1e-5: %d of %d\n", cnt, A.n);
    }
    check(ok, "whole substeps: baseline kernels == tile kernels with the F-update inside P2G");
    if (!(moved && stressed && plastic)) std::printf("      moved %d stressed %d plastic %d\n", (int)moved, (int)stressed, (int)plastic);
    check(moved && stressed && plastic, "flow scene exercises collision, stress and plastic clamping");
}

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory wavefronts of P2G under the emulator's bank model (cuda_emu.h: SmemProbe), on the layout the benchmark
// scene has: 8 particles in every cell, ids cell-ordered by the previous substep's P2G and re-sorted by the gather.
// Prints wavefronts per 512-particle block for every probed instruction, with the aligned record walk (the kernel
// measured in round 1) and with the rotated walk, and checks the model against the ncu count of the aligned walk.
// ---------------------------------------------------------------------------------------------------------------
struct SmemRun { std::map<int, emu::SmemProbe::Site> sites; emu::SmemProbe::Site gather; size_t blocks = 0; };
static SmemRun smem_profile_run(unsigned seed, int rotate) {
    const float dt = 1e-5f;
    Host H;
    build_flow_scene(H, seed);
    H.sc.p2g_rotate = rotate;
    ColliderSet cs;
    std::memset(&cs, 0, sizeof cs);
    int cur = 0;
    std::vector<int> blocks(H.gd.n_gblocks);
    for (int b = 0; b < H.gd.n_gblocks; ++b) blocks[b] = b;
    emu::SmemProbe& pr = emu::smem_probe_state();
    SmemRun out;
    for (int step = 0; step < 3; ++step) {
        host_bin(H, cur);
        H.dc.n_active_gblocks = H.gd.n_gblocks;
        std::vector<int> ids = H.ids0;
        std::fill(H.grid.begin(), H.grid.end(), make_float4(0, 0, 0, 0));
        Planes C = H.planes(cur), N = H.planes(cur ^ 1);
        H.dc.work_a = 0; H.dc.work_b = 0;
        pr.reset();
        pr.on = step == 2;               // steady state: the ids were cell-ordered by two earlier P2G passes
        // the lanes of a warp reach the counting sort's shared atomics together and the 8 particles of a cell are 8
        // consecutive lanes: served in lane order (as the hardware's measured conflict count implies, see below) every
        // cell keeps its particles in the same relative order -> the runs stay aligned
        emu::lane_order = true;
        emu::launch(3, P2G_T, sizeof(P2GSmem), [&] {
            k_p2g_tile<P2G_FUSED>(C, ids.data(), H.work.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt, C);
        });
        pr.on = false;
        if (step == 2) { out.sites = pr.sites; out.blocks = H.work.size(); }
        emu::launch(3, 256, 0, [&] { k_grid_update<GU_NORMALIZE | GU_GRAVITY>(blocks.data(), &H.dc, H.grid.data(), nullptr, H.gd, H.sc, dt, cs, 0); });
        emu::launch((H.n + 255) / 256, 256, 0, [&] { k_fupdate<true>(C, N, ids.data(), &H.dc, H.sc, dt); });
        pr.reset();
        pr.on = step == 2;
        emu::launch(2, G2P_T, sizeof(G2PSmem), [&] {
            k_g2p_tile<G2P_GATHER | G2P_ADVECT | G2P_REORDER>(C, N, ids.data(), H.work.data(), &H.dc, H.grid.data(), H.gd, H.sc, dt);
        });
        pr.on = false;
        if (step == 2) out.gather = pr.sites[40];
        cur ^= 1;
    }
    pr.reset();
    emu::lane_order = false;
    return out;
}
static int smem_profile(unsigned seed) {
    static const struct { int site; const char* what; int mult; } names[] = {
        { 1, "derive: STS.128 record part (x6)", 6 }, { 2, "derive: STS.32 hA8 / gid (x2)", 2 },
        { 5, "sort: STS.U16 order", 1 },
        { 10, "phase 1: LDS.U16 order[i]", 1 }, { 11, "phase 1: LDS.32 wx[a]", 1 }, { 12, "phase 1: LDS.128 wy", 1 }, { 13, "phase 1: LDS.128 wz", 1 },
        { 14, "phase 1: LDS.128 qc", 1 }, { 15, "phase 1: LDS.128 hA0", 1 }, { 16, "phase 1: LDS.128 hA1", 1 }, { 17, "phase 1: LDS.32 hA8", 1 },
        { 20, "phase 2a: STS.128 t1 (k = cz)", 1 }, { 21, "phase 2a: STS.128 t1 (k = cz+4)", 1 }, { 30, "phase 2b: LDS.128 t1", 1 } };
    const SmemRun a = smem_profile_run(seed, 0), r = smem_profile_run(seed, 1);
    std::printf("P2G shared-memory wavefronts per 512-particle block (8 per cell, %zu blocks; model: cuda_emu.h SmemProbe)\n", a.blocks);
    std::printf("  %-36s %10s %12s %12s\n", "instruction", "requests", "aligned walk", "rotated walk");
    double ta = 0, tr = 0, p1a = 0, p1r = 0, ideal_p1 = 0;
    for (const auto& nm : names) {
        const auto ia = a.sites.find(nm.site), ir = r.sites.find(nm.site);
        if (ia == a.sites.end() || ir == r.sites.end()) continue;
        const double req = (double)ia->second.requests * nm.mult / a.blocks, wa = (double)ia->second.wavefronts * nm.mult / a.blocks,
                     wr = (double)ir->second.wavefronts * nm.mult / r.blocks;
        std::printf("  %-36s %10.1f %12.1f %12.1f\n", nm.what, req, wa, wr);
        ta += wa; tr += wr;
        if (nm.site >= 10 && nm.site <= 17) { p1a += wa; p1r += wr; ideal_p1 += req * (nm.site >= 12 && nm.site <= 16 ? 2 : 1); }
    }
    std::printf("  %-36s %10s %12.1f %12.1f\n", "sum of the probed instructions", "", ta, tr);
    std::printf("  phase 1 alone: %.0f -> %.0f wavefronts per block (conflict-free: %.0f)\n", p1a, p1r, ideal_p1);
    // ncu at 64 Mi particles (profiles/ncu_r1_64M_top_kernels.csv): 703.3 M shared wavefronts, 382.7 M of them bank conflicts,
    // 131072 blocks -> 5366 and 2920 per block. The probed instructions of the aligned walk must account for most of both.
    const double ncu_wf = 703320466.0 / 131072.0, ncu_conf = 382739298.0 / 131072.0;
    std::printf("  ncu, aligned walk, per block: %.0f wavefronts, %.0f of them bank conflicts; model: %.0f wavefronts, %.0f above conflict-free in phase 1\n",
                ncu_wf, ncu_conf, ta, p1a - ideal_p1);
    // the gather's tile reads (linear tile, rows padded 8 -> 9 and planes 72 -> 73 float4): ncu counted 2.01 wavefronts per LDS.128
    // on the blocked tile of round 1, which this model reproduced (2.00); the padded linear layout must be as conflict-free
    const double g_model = (double)r.gather.wavefronts / (double)r.gather.requests;
    std::printf("  gather LDS.128 of the linear tile: model %.2f wavefronts per request\n", g_model);
    check(g_model < 2.1, "linear gather tile: the padded layout is conflict-free (2 wavefronts per 512-byte request)");
    check(p1a - ideal_p1 > 0.7 * ncu_conf && p1a - ideal_p1 < 1.1 * ncu_conf, "bank model: the aligned walk's phase-1 conflicts account for 70-110 % of ncu's bank-conflict wavefronts");
    check(p1r <= 1.05 * ideal_p1, "rotated walk: phase 1 is conflict-free on the 8-per-cell layout");
    check(tr < 0.62 * ta, "rotated walk: >= 38 % fewer shared-memory wavefronts per block");
    return failures;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? (unsigned)std::atoi(argv[1]) : 1u;
    if (argc > 3 && std::strcmp(argv[3], "smem") == 0) {
        const int f = smem_profile(seed);
        std::printf("%s (%d failures)\n", f ? "EMULATION CHECKS FAILED" : "all emulation checks passed", f);
        return f ? 1 : 0;
    }
    const bool fast_div = argc > 2 ? std::atoi(argv[2]) != 0 : true;
    const float dt = 1e-5f;
    Host H;
    build(H, 24, seed, fast_div);
    std::printf("emulated scene: %d particles, %zu occupied particle blocks, seed %u, fast pos/h %d\n", H.n, H.work.size(), seed, (int)fast_div);

    // ---------------- weights: one quotient per axis == cell_of + axis_weights ----------------
    {
        std::mt19937 g(seed * 31u + 5u);
        bool same = true;
        for (int i = 0; i < 2000000 && same; ++i) {
            // inside the validated range of the fast quotient, just outside it, and far outside (exact-division fallback)
            const float x = i % 16 == 0 ? frand(g, -1.0f, 3.0f * H.sc.pd.hi) : frand(g, H.sc.pd.lo * 0.9f, H.sc.pd.hi * 1.05f);
            float w1[4], w2[4];
            const int c1 = cell_and_weights(x, H.sc.pd, w1), c2 = cell_of(x, H.sc.pd);
            axis_weights(x, H.sc.pd, c2, w2);
            same = c1 == c2 && std::memcmp(w1, w2, sizeof w1) == 0;
        }
        check(same, "cell_and_weights == cell_of + axis_weights, bit for bit (2 M positions, fast and exact quotient)");
    }

    // ---------------- peer-halo work order: a permutation that visits both ends of the block list first ----------------
    {
        bool perm = true;
        for (int n = 1; n <= 67 && perm; ++n) {
            std::vector<int> seen(n, 0);
            for (int w = 0; w < n; ++w) { const int i = peer_work_order(w, n); if (i < 0 || i >= n || seen[i]++) perm = false; }
            if (n >= 2 && (peer_work_order(0, n) != 0 || peer_work_order(1, n) != n - 1)) perm = false;
        }
        check(perm, "peer_work_order is a permutation of the work list starting with its two ends");
    }

    // ---------------- P2G ----------------
    std::vector<int> ids_def, ids_fu, ids_pkfu, ids_base = H.ids0;
    std::vector<float4> g_def, g_fu, g_pkfu, g_base(H.grid.size(), make_float4(0, 0, 0, 0));
    {
        Planes P = H.planes(0);
        emu::launch((H.n + 127) / 128, 128, 0, [&] { k_p2g_atomic<P2G_FUSED>(P, ids_base.data(), &H.dc, g_base.data(), H.gd, H.sc, dt); });
    }
    run_p2g<0>(H, ids_def, g_def, dt);
    check(ids_are_block_permutations(H, ids_def), "P2G tile kernel: ids stay a permutation inside every block");
    const double e0 = grid_rel_diff(g_def, g_base);
    std::printf("      grid, tile kernel (packed pairs) vs per-particle atomics: rel diff %.3g\n", e0);
    check(e0 < 2e-5, "P2G tile kernel == baseline kernel");

    // ---------------- F-update inside P2G ----------------
    clear_planes(H, 1, 0, 10);
    run_p2g<1>(H, ids_fu, g_fu, dt);
    const std::vector<float4> nxt_fused = H.buf[1];
    check(ids_are_block_permutations(H, ids_fu), "P2G + F-update: ids stay a permutation inside every block");
    check(grid_rel_diff(g_fu, g_base) < 2e-5, "P2G + F-update: grid == baseline kernel");
    clear_planes(H, 1, 0, 10);
    {
        Planes C = H.planes(0), N = H.planes(1);
        emu::launch((H.n + 255) / 256, 256, 0, [&] { k_fupdate<true>(C, N, ids_fu.data(), &H.dc, H.sc, dt); });
    }
    check(planes_bit_equal(nxt_fused, H.buf[1], H.cap, H.n, 4, 10, false), "F-update inside P2G == k_fupdate<true>, planes 4..10 at every sorted rank, bit for bit");
    check(H.dc.svd_failed == 0, "no SVD failure flagged");
    clear_planes(H, 1, 0, 10);
    run_p2g<2>(H, ids_pkfu, g_pkfu, dt);
    const std::vector<float4> nxt_fused2 = H.buf[1];
    check(grid_rel_diff(g_pkfu, g_base) < 2e-5, "P2G + tolerance-form F-update: grid == baseline kernel");
    clear_planes(H, 1, 0, 10);
    {
        Planes C = H.planes(0), N = H.planes(1);
        emu::launch((H.n + 255) / 256, 256, 0, [&] { k_fupdate<true, true>(C, N, ids_pkfu.data(), &H.dc, H.sc, dt); });
    }
    check(planes_bit_equal(nxt_fused2, H.buf[1], H.cap, H.n, 4, 10, false), "tolerance-form F-update inside P2G == k_fupdate<true, FAST>, bit for bit");

    // ---------------- gather ----------------
    std::vector<float4> vel = g_base;           // node velocities = momentum / mass of the scattered grid
    for (float4& v : vel) if (v.x != 0.0f) { v.y /= v.x; v.z /= v.x; v.w /= v.x; }
    const std::vector<float4> r_tile = run_gather(H, ids_def, vel, dt);
    clear_planes(H, 1, 0, 10);
    {
        Planes C = H.planes(0), N = H.planes(1);
        emu::launch((H.n + 127) / 128, 128, 0, [&] { k_g2p_direct<G2P_GATHER | G2P_ADVECT | G2P_REORDER>(C, N, ids_def.data(), &H.dc, vel.data(), H.gd, H.sc, dt); });
    }
    const std::vector<float4> r_dir = H.buf[1];
    const double d0 = planes_rel_diff(r_tile, r_dir, H.cap, H.n, 0, 3);
    std::printf("      gather, tile kernel (linear tile, packed pairs) vs direct gathers: rel diff %.3g\n", d0);
    check(d0 < 2e-5, "gather tile kernel == direct-gather baseline");

    check_flow(seed);

    std::printf("%s (%d failures)\n", failures ? "EMULATION CHECKS FAILED" : "all emulation checks passed", failures);
    return failures ? 1 : 0;
}

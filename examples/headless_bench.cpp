// Headless benchmark of the MPM substep straight through the C ABI (no Python, no viewer): the host-language
// counterpart of the reference's render loop (main.cpp:163-236) without the GL parts.
//
//   g++ -O2 -std=c++17 examples/headless_bench.cpp -Iinclude -Lrealtime-deformations_b200 -lmpm_b200 \
//       -Wl,-rpath,'$ORIGIN/../realtime-deformations_b200' -o examples/headless_bench
//   examples/headless_bench [grid=128] [particles_per_cell_axis=2] [substeps=50]
//
// Scene: a slab of snow (ppc^3 jittered particles per cell) resting on a ground box, reference material constants.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mpm_b200.h"

static unsigned long long splitmix(unsigned long long& s) {
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static float uniform(unsigned long long& s) { return (float)(splitmix(s) >> 40) * (1.0f / 16777216.0f); }

#define CHECK(call)                                                                    \
    do {                                                                               \
        if ((call) != MPM_OK) { std::fprintf(stderr, "%s failed: %s\n", #call, mpm_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int grid = argc > 1 ? std::atoi(argv[1]) : 128;
    const int ppc = argc > 2 ? std::atoi(argv[2]) : 2;
    const int substeps = argc > 3 ? std::atoi(argv[3]) : 50;
    const float h = 0.05f, dt = 1e-5f;
    if (mpm_device_count() == 0) { std::fprintf(stderr, "no CUDA device: libmpm_b200 has no CPU fallback\n"); return 2; }

    // slab: cells [margin, grid-margin) in i and k, `thick` cells in j above the ground's top face
    const int margin = grid / 16 > 4 ? grid / 16 : 4, thick = grid / 8, j0 = 6;
    std::vector<float> pos, vel;
    unsigned long long seed = 20260117ull;
    for (int i = margin; i < grid - margin; ++i)
        for (int j = j0; j < j0 + thick; ++j)
            for (int k = margin; k < grid - margin; ++k)
                for (int s = 0; s < ppc * ppc * ppc; ++s) {
                    const int sx = s % ppc, sy = (s / ppc) % ppc, sz = s / (ppc * ppc);
                    const float jit = 0.5f / ppc;
                    pos.push_back((i + (sx + 0.5f) / ppc + (uniform(seed) - 0.5f) * jit) * h);
                    pos.push_back((j + (sy + 0.5f) / ppc + (uniform(seed) - 0.5f) * jit) * h);
                    pos.push_back((k + (sz + 0.5f) / ppc + (uniform(seed) - 0.5f) * jit) * h);
                    vel.push_back(0.0f); vel.push_back(0.0f); vel.push_back(0.0f);
                }
    const long long n = (long long)pos.size() / 3;
    std::vector<float> mass((size_t)n, 0.00006f);          // material_point_method.cpp:53

    MpmParams prm;
    mpm_default_params(&prm);
    prm.gravity[0] = 9.8f * std::sin(0.5235988f); prm.gravity[1] = -9.8f * std::cos(0.5235988f);   // 30 degree incline
    mpm_t* sim = nullptr;
    CHECK(mpm_create(&prm, grid, grid, grid, n, &sim));
    CHECK(mpm_upload_particles_soa(sim, n, pos.data(), vel.data(), mass.data(), nullptr, nullptr, nullptr, nullptr));
    CHECK(mpm_rasterize_particles_to_grid(sim));                    // main.cpp:53-54
    CHECK(mpm_compute_particle_volumes_and_densities(sim));

    // ground box through the scene front-end: pose = what a reference MeshCollider's sdf reads (scale, quaternion,
    // translation; hpp:80-83); the library forms the world-to-local matrix in glm's operation order
    const float top = (j0 - 0.5f) * h;
    MpmBoxTransform pose = { { grid * h, 2.0f, grid * h }, { 1.0f, 0.0f, 0.0f, 0.0f }, { grid * h * 0.5f, top - 2.0f, grid * h * 0.5f }, { 0.0f, 0.0f, 0.0f } };
    MpmBoxCollider ground;
    CHECK(mpm_box_collider_from_transform(&pose, &ground));

    CHECK(mpm_substep(sim, dt, &ground, 1, 5));                     // warm-up
    CHECK(mpm_synchronize(sim));
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(mpm_substep(sim, dt, &ground, 1, substeps));
    CHECK(mpm_synchronize(sim));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    MpmStats st;
    CHECK(mpm_get_stats(sim, &st));
    std::printf("{\"particles\": %lld, \"grid\": %d, \"substeps\": %d, \"ms_per_substep\": %.4f, \"particle_updates_per_s\": %.4e, "
                "\"active_nodes\": %lld, \"kernel_ms\": {\"bin\": %.4f, \"clear\": %.4f, \"p2g\": %.4f, \"grid\": %.4f, \"g2p\": %.4f}}\n",
                n, grid, substeps, sec * 1e3 / substeps, (double)n * substeps / sec, (long long)st.n_active_nodes,
                st.last_ms[0], st.last_ms[1], st.last_ms[2], st.last_ms[3], st.last_ms[4]);
    mpm_destroy(sim);
    return 0;
}

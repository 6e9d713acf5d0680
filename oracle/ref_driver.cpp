// TEST INFRASTRUCTURE — not part of the product path.
//
// Headless driver for the UNMODIFIED reference MPM class
// (MaterialPointMethod::LagrangeEulerView, /root/reference/realtime-deformations/
// material_point_method.{hpp,cpp}). The reference sources are compiled where they lie by
// oracle/Makefile; this file only supplies what the viewer's main.cpp would have supplied:
//   * no-op stand-ins for the three GLEW entry points Mesh's constructor touches (mesh.hpp:28-36),
//   * Mesh::draw and MeshPresets::Box (mesh.cpp, which needs libGL, is not built),
//   * devH / cuP2G (cudaCalc.cuh:4-7; dead path, material_point_method.cpp:65-76),
//   * a loop that repeats main.cpp:48-54 (start-up), main.cpp:119-156 (colliders) and
//     main.cpp:192-218 (the seven stage calls per substep) and dumps state as raw float32.
//
// It can be linked against either the reference's material_point_method.cpp (-> oracle/_ref/ref_mpm,
// the golden generator and "reference" CPU baseline) or adapter/lagrange_euler_view_b200.cpp
// (-> oracle/_ref/adapter_mpm, the drop-in demonstration running on the GPU library).
//
// Dump layouts (little-endian float32, no header; shapes are implied by n / grid size):
//   particles: n x 35 = mass, velocity[3], volume, pos[3], FElastic[9], FPlastic[9], B[9]
//              (3x3 blocks are glm column-major exactly as stored in memory: m[col][row])
//   grid:      I*J*K x 7 = mass, force[3], velocity[3], node index i*J*K + j*K + k
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <set>
#include <chrono>
#include <sstream>

// Stage-level golden grids need the private `grid` member; this only changes access control in
// THIS translation unit (class layout is unaffected, the reference TU is compiled untouched).
#define private public
#define protected public
#include "material_point_method.hpp"
#undef private
#undef protected
#ifndef DRIVER_NO_KAT
#include "mathy.hpp"       // Optimizer / Objective and the vendored mcloptlib (implicit-integration known-answer modes)
#endif

// ---- stand-ins for what the GL viewer would have linked -------------------------------------
static void GLAPIENTRY noopGenBuffers(GLsizei n, GLuint* b) { for (GLsizei i = 0; i < n; ++i) b[i] = 0; }
static void GLAPIENTRY noopBindBuffer(GLenum, GLuint) {}
static void GLAPIENTRY noopBufferData(GLenum, GLsizeiptr, const void*, GLenum) {}
PFNGLGENBUFFERSPROC __glewGenBuffers = noopGenBuffers;
PFNGLBINDBUFFERPROC __glewBindBuffer = noopBindBuffer;
PFNGLBUFFERDATAPROC __glewBufferData = noopBufferData;
void Mesh::draw() {}
std::vector<GLfloat> MeshPresets::Box::vertices = { -1.f, -1.f, -1.f, 1.f, 1.f, 1.f };
std::vector<GLfloat> MeshPresets::Box::colors = { 1.f, 1.f, 1.f, 1.f, 1.f, 1.f };
#ifndef DRIVER_NO_CUDA_STUBS
ftype devH;
void cuP2G(MaterialPointMethod::Particle*, MaterialPointMethod::Cell*, int, int, int, int, ftype*) {}
#endif

namespace MPM = MaterialPointMethod;

static void writeFile(const std::string& path, const std::vector<float>& data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fwrite(data.data(), sizeof(float), data.size(), f);
    fclose(f);
}

static std::vector<float> readFile(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> d(sz / sizeof(float));
    if (fread(d.data(), sizeof(float), d.size(), f) != d.size()) { fprintf(stderr, "short read %s\n", path.c_str()); exit(2); }
    fclose(f);
    return d;
}

static void dumpParticles(MPM::LagrangeEulerView& mpm, const std::string& path) {
    const int n = mpm.getNumParticles();
    const MPM::Particle* p = mpm.getParticles();
    std::vector<float> out((size_t)n * 35);
    for (int i = 0; i < n; ++i) {
        float* o = &out[(size_t)i * 35];
        o[0] = p[i].mass;
        o[1] = p[i].velocity.x; o[2] = p[i].velocity.y; o[3] = p[i].velocity.z;
        o[4] = p[i].volume;
        o[5] = p[i].pos.x; o[6] = p[i].pos.y; o[7] = p[i].pos.z;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
            o[8 + c * 3 + r] = p[i].FElastic[c][r];
            o[17 + c * 3 + r] = p[i].FPlastic[c][r];
            o[26 + c * 3 + r] = p[i].B[c][r];
        }
    }
    writeFile(path, out);
}

static void dumpGrid(MPM::LagrangeEulerView& mpm, const std::string& path) {
    const auto& g = mpm.grid.grid;
    std::vector<float> out(g.size() * 7);
    for (size_t i = 0; i < g.size(); ++i) {
        float* o = &out[i * 7];
        o[0] = g[i].mass;
        o[1] = g[i].force.x; o[2] = g[i].force.y; o[3] = g[i].force.z;
        o[4] = g[i].velocity.x; o[5] = g[i].velocity.y; o[6] = g[i].velocity.z;
    }
    writeFile(path, out);
}

static std::set<int> parseList(const char* s) {
    std::set<int> r;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) r.insert(atoi(tok.c_str()));
    return r;
}

struct BoxSpec { float t[3]; float rotZdeg; float s[3]; };

int main(int argc, char** argv) {
    int I = 20, J = 20, K = 20, n = 270 + 1877, steps = 0;
    float h = 0.05f, dt = 1e-5f;
    glm::vec3 origin(0.5f, 0.6f, 0.5f), v0(0.0f, -200.0f, 0.0f);
    std::string dumpDir, loadPath, loadFullPath, katMode, katIn, katOut, collidersPath;
    std::set<int> dumpSteps, stageSteps;
    bool bench = false, quiet = false, noColliders = false, haveColVel = false;
    glm::vec3 colVel(0.0f);        // MeshCollider::velocity of every box (main.cpp:121-151 uses 0; key_callback flips its sign)
    for (int a = 1; a < argc; ++a) {
        auto is = [&](const char* k) { return strcmp(argv[a], k) == 0; };
        if (is("--grid")) { I = atoi(argv[++a]); J = atoi(argv[++a]); K = atoi(argv[++a]); }
        else if (is("--n")) n = atoi(argv[++a]);
        else if (is("--h")) h = (float)atof(argv[++a]);
        else if (is("--dt")) dt = (float)atof(argv[++a]);
        else if (is("--steps")) steps = atoi(argv[++a]);
        else if (is("--origin")) { origin.x = atof(argv[++a]); origin.y = atof(argv[++a]); origin.z = atof(argv[++a]); }
        else if (is("--v0")) { v0.x = atof(argv[++a]); v0.y = atof(argv[++a]); v0.z = atof(argv[++a]); }
        else if (is("--dump-dir")) dumpDir = argv[++a];
        else if (is("--dump-steps")) dumpSteps = parseList(argv[++a]);
        else if (is("--stage-steps")) stageSteps = parseList(argv[++a]);
        else if (is("--load")) loadPath = argv[++a];     // n x 7 float32: pos, vel, mass (overrides the rand() scene)
        else if (is("--bench")) bench = true;
        else if (is("--quiet")) quiet = true;
        else if (is("--no-colliders")) noColliders = true;
        else if (is("--collider-vel")) { colVel.x = atof(argv[++a]); colVel.y = atof(argv[++a]); colVel.z = atof(argv[++a]); haveColVel = true; }
        else if (is("--colliders")) collidersPath = argv[++a];   // rows of 7 float32: translation[3], rotZ degrees, scale[3]
        else if (is("--load-full")) loadFullPath = argv[++a];   // n x 35 float32, same layout as the particle dumps
        else if (is("--kat-weights")) { katMode = "weights"; katIn = argv[++a]; katOut = argv[++a]; }
        else if (is("--kat-polar")) { katMode = "polar"; katIn = argv[++a]; katOut = argv[++a]; }
        else if (is("--kat-collide")) { katMode = "collide"; katIn = argv[++a]; katOut = argv[++a]; }
        else if (is("--kat-fupdate")) { katMode = "fupdate"; katOut = argv[++a]; }   // needs --load-full
        else if (is("--kat-energy")) { katMode = "energy"; katIn = argv[++a]; katOut = argv[++a]; }    // in: I*J*K x 3 velocity perturbation
        else if (is("--kat-implicit")) { katMode = "implicit"; katOut = argv[++a]; }
        else if (is("--kat-lbfgs")) { katMode = "lbfgs"; katIn = argv[++a]; katOut = argv[++a]; }
        else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
    }

    // ---- main.cpp:48-54 ----
    MPM::WeightCalculator::h = h;   // file-scope default 0.05f (material_point_method.cpp:17)
    MPM::LagrangeEulerView sim{ I, J, K, n };
    sim.setLevel(MPM::DEFAULT_LOG_LEVEL_MPM);
    if (loadPath.empty()) {
        sim.initializeParticles(origin, v0);
    } else {
        FILE* f = fopen(loadPath.c_str(), "rb");
        if (!f) { fprintf(stderr, "cannot read %s\n", loadPath.c_str()); return 2; }
        std::vector<float> in((size_t)n * 7);
        if (fread(in.data(), sizeof(float), in.size(), f) != in.size()) { fprintf(stderr, "short read\n"); return 2; }
        fclose(f);
        MPM::Particle* p = sim.getParticles();
        for (int i = 0; i < n; ++i) {
            p[i].pos = { in[i * 7 + 0], in[i * 7 + 1], in[i * 7 + 2] };
            p[i].velocity = { in[i * 7 + 3], in[i * 7 + 4], in[i * 7 + 5] };
            p[i].mass = in[i * 7 + 6];
            p[i].r = p[i].g = p[i].b = p[i].a = 255; p[i].size = 0.02f;
        }
    }
    if (!loadFullPath.empty()) {
        std::vector<float> in = readFile(loadFullPath);
        if (in.size() != (size_t)n * 35) { fprintf(stderr, "--load-full: expected %d x 35 floats\n", n); return 2; }
        MPM::Particle* p = sim.getParticles();
        for (int i = 0; i < n; ++i) {
            const float* o = &in[(size_t)i * 35];
            p[i].mass = o[0]; p[i].velocity = { o[1], o[2], o[3] }; p[i].volume = o[4]; p[i].pos = { o[5], o[6], o[7] };
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
                p[i].FElastic[c][r] = o[8 + c * 3 + r]; p[i].FPlastic[c][r] = o[17 + c * 3 + r]; p[i].B[c][r] = o[26 + c * 3 + r];
            }
        }
    }
    // ---- known-answer modes: call the reference's own functions on caller-supplied inputs ----
    if (katMode == "weights") {       // material_point_method.hpp:20-31
        std::vector<float> in = readFile(katIn), out(in.size());
        for (size_t i = 0; i < in.size(); ++i) out[i] = MPM::WeightCalculator::weightNx(in[i]);
        writeFile(katOut, out); return 0;
    }
    if (katMode == "polar") {         // utils.h:55-75 (Higham-Noferini branch), glm column-major 3x3 in, R then S out
        std::vector<float> in = readFile(katIn), out(in.size() * 2);
        for (size_t i = 0; i + 9 <= in.size(); i += 9) {
            m3t F; for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) F[c][r] = in[i + c * 3 + r];
            const auto RS = polarDecomposition(F);
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { out[2 * i + c * 3 + r] = RS.first[c][r]; out[2 * i + 9 + c * 3 + r] = RS.second[c][r]; }
        }
        writeFile(katOut, out); return 0;
    }
#ifndef DRIVER_NO_KAT
    if (katMode == "lbfgs") {
        // the vendored optimiser exactly as mathy.hpp:10-19 configures it (LBFGS<float, Dynamic>, Backtracking) with the
        // convergence rule of mathy.hpp:31-35, on a smooth test function with an analytic gradient:
        //   f(x) = sum_i 0.5 a_i (x_i - t_i)^2 + 0.25 sum_i (x_i - x_{i+1})^4        in: n, a[n], t[n], x0[n]
        const std::vector<float> in = readFile(katIn);
        const int dim = (int)in[0];
        struct Fn : public mcl::optlib::Problem<ftype, Eigen::Dynamic> {
            const float *a, *t; int n;
            ftype value(const Eigen::VectorXf& x) override {
                float f = 0.0f;
                for (int i = 0; i < n; ++i) { const float d = x[i] - t[i]; f += 0.5f * a[i] * d * d; }
                for (int i = 0; i + 1 < n; ++i) { const float d = x[i] - x[i + 1]; f += 0.25f * d * d * d * d; }
                return f;
            }
            ftype gradient(const Eigen::VectorXf& x, Eigen::VectorXf& g) override {
                for (int i = 0; i < n; ++i) g[i] = a[i] * (x[i] - t[i]);
                for (int i = 0; i + 1 < n; ++i) { const float d = x[i] - x[i + 1]; g[i] += d * d * d; g[i + 1] -= d * d * d; }
                return value(x);
            }
            bool converged(const Eigen::VectorXf& x0, const Eigen::VectorXf& x1, const Eigen::VectorXf& grad) override {
                if (grad.norm() < 1e-2) { return true; }
                if ((x0 - x1).norm() < 1e-2) { return true; }
                return false;
            }
        } fn;
        fn.a = &in[1]; fn.t = &in[1 + dim]; fn.n = dim;
        Eigen::VectorXf x(dim);
        for (int i = 0; i < dim; ++i) x[i] = in[1 + 2 * dim + i];
        mcl::optlib::LBFGS<ftype, Eigen::Dynamic> opt;
        opt.m_settings.ls_method = mcl::optlib::LSMethod::Backtracking;
        const int iters = opt.minimize(fn, x);
        std::vector<float> out(dim + 2);
        out[0] = (float)iters; out[1] = fn.value(x);
        for (int i = 0; i < dim; ++i) out[2 + i] = x[i];
        writeFile(katOut, out); return 0;
    }
    if (katMode == "energy" || katMode == "implicit") {
        // material_point_method.cpp:160-233 on the loaded state: rasterize (used_cells, grid velocities), then either
        // Energy(grid velocities + perturbation, dt) or the whole timeIntegration(dt)
        sim.rasterizeParticlesToGrid();
        if (loadFullPath.empty()) sim.computeParticleVolumesAndDensities();
        const int nu = (int)sim.used_cells.size();
        if (katMode == "energy") {
            const std::vector<float> pert = readFile(katIn);
            Eigen::VectorXf v(nu * 3);
            int q = 0;
            for (const auto& c : sim.used_cells) {
                const size_t node = ((size_t)c.x * J + c.y) * K + c.z;
                const auto gv = sim.grid(c.x, c.y, c.z).velocity;
                for (int a2 = 0; a2 < 3; ++a2) v[q++] = gv[a2] + pert[node * 3 + a2];
            }
            std::vector<float> out = { sim.Energy(v, dt), sim.ElasticPotential(v, dt), (float)nu };
            writeFile(katOut, out); return 0;
        }
        sim.timeIntegration(dt);
        dumpGrid(sim, katOut); return 0;
    }
#endif
    if (katMode == "fupdate") {       // material_point_method.cpp:306-330 on the loaded state
        sim.updateDeformationGradient(dt);
        dumpParticles(sim, katOut); return 0;
    }
    if (katMode != "collide" && loadFullPath.empty()) {
        sim.rasterizeParticlesToGrid();
        sim.computeParticleVolumesAndDensities();
    }
    if (!dumpDir.empty()) dumpParticles(sim, dumpDir + "/particles_step0000.f32");

    // ---- main.cpp:113-156: colliders (box2 is built there but never pushed) ----
    glm::mat4 VP;
    std::vector<MPM::MeshCollider> solidObjects;
    MPM::MeshCollider box1{ 0, VP, MeshPresets::Box::vertices, MeshPresets::Box::colors, {0, 0, 0} };
    {
        const auto rotation1 = glm::rotate(glm::mat4(), glm::radians(0.0f), { 0, 0, 1 });
        const auto translation1 = glm::translate(glm::mat4(), { 0.5, -0.2, 0.5 });
        const auto scaling1 = glm::scale(glm::mat4(), { 0.4f, 0.4f, 0.4f });
        box1.mesh.applyMatrix4(translation1 * rotation1 * scaling1);
    }
    MPM::MeshCollider box3{ 0, VP, MeshPresets::Box::vertices, MeshPresets::Box::colors, {0, 0, 0} };
    const auto rotation3 = glm::rotate(glm::mat4(), glm::radians(45.0f), { 0, 0, 1 });
    const auto scaling3 = glm::scale(glm::mat4(), glm::vec3(0.2f, 0.2f, 0.3f));
    {
        const auto translation3 = glm::translate(glm::mat4(), { 0.0, 0.3, 0.5 });
        box3.mesh.applyMatrix4(translation3 * rotation3 * scaling3);
    }
    MPM::MeshCollider box4{ 0, VP, MeshPresets::Box::vertices, MeshPresets::Box::colors, {0, 0, 0} };
    {
        const auto translation4 = glm::translate(glm::mat4(), { 1.0, 0.3, 0.5 });
        box4.mesh.applyMatrix4(translation4 * rotation3 * scaling3);
    }
    std::vector<MPM::MeshCollider> customBoxes;     // built exactly like main.cpp:119-151 builds box1..box4
    if (!collidersPath.empty()) {
        const std::vector<float> rows = readFile(collidersPath);
        customBoxes.reserve(rows.size() / 7);        // the sdf lambdas capture `this`: no reallocation allowed
        for (size_t r = 0; r + 7 <= rows.size(); r += 7) {
            customBoxes.emplace_back(0, VP, MeshPresets::Box::vertices, MeshPresets::Box::colors, glm::vec3{ 0, 0, 0 });
            const auto rot = glm::rotate(glm::mat4(), glm::radians(rows[r + 3]), { 0, 0, 1 });
            const auto tr = glm::translate(glm::mat4(), { rows[r + 0], rows[r + 1], rows[r + 2] });
            const auto sc = glm::scale(glm::mat4(), glm::vec3(rows[r + 4], rows[r + 5], rows[r + 6]));
            customBoxes.back().mesh.applyMatrix4(tr * rot * sc);
        }
        for (auto& b : customBoxes) solidObjects.push_back(b);
    } else if (!noColliders) {
        solidObjects.push_back(box1);
        solidObjects.push_back(box3);
        solidObjects.push_back(box4);
    }
    if (haveColVel) for (auto& o : solidObjects) o.velocity = colVel;
    if (!dumpDir.empty()) {
        // what the sdf lambda actually uses (material_point_method.hpp:80-83): scale, quat, translation
        std::vector<float> c;
        for (auto& o : solidObjects) {
            const auto& m = o.mesh;
            c.insert(c.end(), { m.scale.x, m.scale.y, m.scale.z, m.rotation.w, m.rotation.x, m.rotation.y, m.rotation.z,
                                m.translation.x, m.translation.y, m.translation.z, o.velocity.x, o.velocity.y, o.velocity.z });
            const glm::mat4 inv = glm::inverse(glm::translate(glm::mat4(), m.translation) * glm::toMat4(m.rotation));
            for (int cc = 0; cc < 4; ++cc) for (int rr = 0; rr < 4; ++rr) c.push_back(inv[cc][rr]);
        }
        writeFile(dumpDir + "/colliders.f32", c);
    }

#ifndef DRIVER_NO_KAT                 // (the adapter build has no private bodyCollision to call)
    if (katMode == "collide") {       // material_point_method.cpp:264-296; in: m x 6 (pos, vel) -> out: m x 3
        std::vector<float> in = readFile(katIn), out(in.size() / 2);
        for (size_t i = 0; i + 6 <= in.size(); i += 6) {
            const auto v = sim.bodyCollision({ in[i], in[i + 1], in[i + 2] }, { in[i + 3], in[i + 4], in[i + 5] }, dt, solidObjects);
            out[i / 2] = v.x; out[i / 2 + 1] = v.y; out[i / 2 + 2] = v.z;
        }
        writeFile(katOut, out); return 0;
    }
#endif

    // ---- main.cpp:163-218: one substep per frame ----
    char name[256];
    auto t0 = std::chrono::steady_clock::now();
    double stageMs[7] = { 0, 0, 0, 0, 0, 0, 0 };
    for (int s = 1; s <= steps; ++s) {
        const bool st = !dumpDir.empty() && stageSteps.count(s);
        auto stageDumpP = [&](const char* tag) { if (st) { snprintf(name, sizeof name, "%s/stage%04d_%s.f32", dumpDir.c_str(), s, tag); dumpParticles(sim, name); } };
        auto stageDumpG = [&](const char* tag) { if (st) { snprintf(name, sizeof name, "%s/stage%04d_%s.f32", dumpDir.c_str(), s, tag); dumpGrid(sim, name); } };
        auto tick = [&]() { return std::chrono::steady_clock::now(); };
        auto acc = [&](int k, std::chrono::steady_clock::time_point a) { stageMs[k] += std::chrono::duration<double, std::milli>(tick() - a).count(); };
        stageDumpP("pre");
        for (auto& box : solidObjects) box.move(dt);          // main.cpp:187-190 (velocity 0)
        auto a = tick(); sim.rasterizeParticlesToGrid();      acc(0, a); stageDumpG("p2g");
        a = tick(); sim.computeExplicitGridForces();          acc(1, a); stageDumpG("forces");
        a = tick(); sim.gridVelocitiesUpdate(dt);             acc(2, a); stageDumpG("gridvel");
        a = tick(); sim.gridBasedCollisions(dt, solidObjects); acc(3, a); stageDumpG("collide");
        a = tick(); sim.updateDeformationGradient(dt);        acc(4, a); stageDumpP("fupdate");
        a = tick(); sim.updateParticleVelocities();           acc(5, a); stageDumpP("g2p");
        a = tick(); sim.updateParticlePositions(dt);          acc(6, a); stageDumpP("advect");
        if (!dumpDir.empty() && dumpSteps.count(s)) {
            snprintf(name, sizeof name, "%s/particles_step%04d.f32", dumpDir.c_str(), s);
            dumpParticles(sim, name);
        }
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!dumpDir.empty()) {
        // the colliders once more, after `steps` calls of MeshCollider::move (hpp:90-92: applyMatrix4 + glm::decompose)
        std::vector<float> c;
        for (auto& o : solidObjects) {
            const auto& m = o.mesh;
            c.insert(c.end(), { m.scale.x, m.scale.y, m.scale.z, m.rotation.w, m.rotation.x, m.rotation.y, m.rotation.z,
                                m.translation.x, m.translation.y, m.translation.z, o.velocity.x, o.velocity.y, o.velocity.z });
            const glm::mat4 inv = glm::inverse(glm::translate(glm::mat4(), m.translation) * glm::toMat4(m.rotation));
            for (int cc = 0; cc < 4; ++cc) for (int rr = 0; rr < 4; ++rr) c.push_back(inv[cc][rr]);
        }
        writeFile(dumpDir + "/colliders_final.f32", c);
    }
    if (bench || !quiet) {
        printf("{\"impl\": \"reference\", \"n_particles\": %d, \"grid\": [%d, %d, %d], \"steps\": %d, \"seconds\": %.6f, "
               "\"particle_updates_per_s\": %.3f, \"stage_ms\": [%.3f, %.3f, %.3f, %.3f, %.3f, %.3f, %.3f]}\n",
               n, I, J, K, steps, sec, steps > 0 ? (double)n * steps / sec : 0.0,
               stageMs[0], stageMs[1], stageMs[2], stageMs[3], stageMs[4], stageMs[5], stageMs[6]);
    }
    return 0;
}

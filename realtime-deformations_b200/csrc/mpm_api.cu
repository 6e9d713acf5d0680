// libmpm_b200.so — host side of the C ABI declared in include/mpm_b200.h.
// Each entry point names the reference member it replaces (material_point_method.{hpp,cpp}); the kernels are in
// mpm_kernels.cuh / mpm_tile_kernels.cuh. No CPU fallback: every compute call needs a CUDA device.
#include "../../include/mpm_b200.h"
#include "mpm_kernels.cuh"
#include "mpm_tile_kernels.cuh"
#include "mpm_implicit.cuh"

#ifndef MPM_HOST_EMU
#include <nvtx3/nvToolsExt.h>        // header-only; no-ops unless a profiler (nsys / ncu --nvtx) is attached
#define MPM_NVTX_PUSH(name) nvtxRangePushA(name)
#define MPM_NVTX_POP() nvtxRangePop()
#else
#define MPM_NVTX_PUSH(name)
#define MPM_NVTX_POP()
#endif

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace mpm;

static thread_local std::string g_last_error;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(MPM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKLAUNCH() CK(cudaGetLastError())
#define NEED(s) do { if (!(s)) return fail(MPM_ERR_INVALID, "null handle"); CK(cudaSetDevice((s)->device)); } while (0)
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

constexpr size_t PEER_FLAG_BYTES = 256;      // flag words: [0]/[1] cleared by lower/upper neighbour, [2]/[3] P2G done, [4]/[5] migration packed, [6]/[7] migration consumed

struct mpm_sim {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    MpmParams prm;
    GridDims gd;
    SimConst sc;
    int64_t capacity = 0;      // slots per plane
    int64_t n_uploaded = 0;    // pid space (upload order)
    int64_t n_bound = 0;       // host-side upper bound of occupied slots
    float4* buf[2] = { nullptr, nullptr };   // two particle buffers, NPLANES * capacity float4 each
    int cur = 0;
    int *key = nullptr, *sorted_ids = nullptr;
    int *blk_count = nullptr, *blk_start = nullptr, *blk_cursor = nullptr;
    int4* pblock_list = nullptr;      // occupied particle blocks as work items (block id, first sorted rank, count, -)
    int *gflag = nullptr, *gblock_list = nullptr;
    int2* partial = nullptr;
    int n_buckets = 0, n_chunks = 0;
    float4 *grid = nullptr, *gforce = nullptr;
    DevCounters* dc = nullptr;
    float4* out_buf[2] = { nullptr, nullptr };   // migration: packed outgoing particles (down, up)
    int64_t out_cap = 0;
    int64_t pid_base = 0;
    float4* render_stage = nullptr; int64_t render_cap = 0;      // async render path: device staging + copy stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t render_ready = nullptr, copy_done = nullptr;
    bool copy_pending = false;
    // opt-in (MPM_B200_GRAPH=1; bench.py --config 1 uses it: 96 us per substep at 2147 particles): a CUDA graph of two consecutive fused
    // substeps (the particle buffers ping-pong, so two substeps return every host-side pointer to where it started)
    bool graph_enabled = false, capturing = false;
    cudaGraphExec_t graph_exec = nullptr;
    float graph_dt = 0.0f; int graph_nc = -1; int64_t graph_n_bound = -1; int graph_cur = -1; int graph_launches = 0;
    ColliderSet graph_cols;
    void* pinned = nullptr; size_t pinned_bytes = 0;
    bool tau_valid = false, binned = false;
    bool hist_valid = false;   // key[] and blk_count[] already describe the current buffer (written by the fused substep's gather)
    bool hist_fuse = true;     // MPM_B200_FUSE_HIST=0 restores the separate k_bin_count pass (A/B)
    // implicit time integration (mpm_implicit.cuh): node vectors in grid layout, allocated on first use
    enum { IMP_X = 0, IMP_G, IMP_Q, IMP_XOLD, IMP_GOLD, IMP_TMP, IMP_S0, IMP_Y0 = IMP_S0 + 8, IMP_NVEC = IMP_Y0 + 8 };
    float4* imp_vec[IMP_NVEC] = {};
    float4* imp_aux = nullptr;        // 3 float4 per sorted rank: I + dt grad v, then the stress matrix Gm (block-tile form)
    double* imp_acc = nullptr;        // [0] inertia energy, [1] elastic energy, [2] dot product, [3] (int) |x|_inf bits
    bool mig_packed = false;   // slab handles: the gather of the last substep has already packed the leavers into out_buf (fused migration)
    // peer-memory halo (mpm_peer_connect*, mpm_substep_begin_peer): neighbours' shared layers and flag words
    PeerLayers peer = { nullptr, nullptr };
    int* peer_flags_dn = nullptr; int* peer_flags_up = nullptr;      // the neighbours' flag words (peer-mapped)
    void* ipc_dn = nullptr; void* ipc_up = nullptr;                  // mappings opened by mpm_peer_connect (closed in destroy)
    bool peer_connected = false;
    int peer_epoch = 0;
    const float4* peer_in_dn = nullptr; const float4* peer_in_up = nullptr;   // lower neighbour's UP buffer, upper neighbour's DOWN buffer
    void* ipc_mig_dn = nullptr; void* ipc_mig_up = nullptr;
    bool peer_mig_connected = false;
    int peer_mig_epoch = 0;
    int* slot_of_pid = nullptr; // p2g_variant 9 (deterministic debug mode): binned slot of every particle id
    bool fupd_pending = false;  // P2G has put the F-update results into the idle buffer (until substep_end)
    int num_sms = 148;
    cudaEvent_t ev[8];
    bool ev_ok = false;
    SideStream side = { nullptr, nullptr, nullptr, 1, nullptr, false };   // F-update || gather: opt-in, MPM_B200_OVERLAP=1|2
    MpmStats stats;
    ColliderSet colliders; int n_colliders = 0;

    Planes planes(int which) const {
        Planes P;
        for (int k = 0; k < NPLANES; ++k) P.p[k] = buf[which] + (size_t)k * capacity;
        return P;
    }
};

void mpm_default_params(MpmParams* p) {
    memset(p, 0, sizeof *p);
    p->h = 0.05f; p->youngs_modulus = 1.4e5f; p->poisson_ratio = 0.2f; p->hardening_xi = 10.0f;
    p->theta_c = (float)(2.5f * 1e-2); p->theta_s = (float)(5.0f * 1e-3);
    p->gravity[0] = 0.0f; p->gravity[1] = (float)-9.8; p->gravity[2] = 0.0f;
    p->friction_mu = 0.5f;
}
const char* mpm_last_error(void) { return g_last_error.c_str(); }
int mpm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// DpInverse exactly as the reference builds it (hpp:177): inverse(mat3(1) * (1.0f/3.0f) * h * h), glm formulas
// (quadratic stencil: D = h^2/4, the same expression with 1/4)
static float dp_inverse_scalar(float h, bool quadratic) {
    volatile float d = 1.0f * (quadratic ? 1.0f / 4.0f : 1.0f / 3.0f); d = d * h; d = d * h;
    volatile float dd = d * d;                       // m11*m22 - 0*0
    volatile float det = d * dd;                     // + d*(dd - 0) - 0 + 0
    volatile float ood = 1.0f / det;
    volatile float r = dd * ood;
    return r;
}

static void fill_consts(mpm_sim* s) {
    const MpmParams& p = s->prm;
    SimConst& c = s->sc;
    c.h = p.h; c.dinv = dp_inverse_scalar(p.h, p.stencil == 1); c.E = p.youngs_modulus; c.nu = p.poisson_ratio; c.xi = p.hardening_xi;
    c.clamp_lo = (float)(1.0 - (double)p.theta_c); c.clamp_hi = (float)(1.0 + (double)p.theta_s);   // cpp:320
    c.friction = p.friction_mu;
    c.mu0 = c.E / (2.0f * (1.0f + c.nu)); c.lambda0 = (c.E * c.nu) / ((1.0f + c.nu) * (1.0f - 2.0f * c.nu));   // cpp:237-238 (same float divisions as lame())
    for (int a = 0; a < 3; ++a) c.g[a] = p.gravity[a];
    c.pos_lo = (float)(3 * p.h);                                                                     // cpp:383
    c.pos_hi[0] = (float)((s->gd.I - 3) * p.h); c.pos_hi[1] = (float)((s->gd.J - 3) * p.h); c.pos_hi[2] = (float)((s->gd.K - 3) * p.h);
    // P2G accumulation loop: rotated record walk (see k_p2g_tile). On by default; MPM_B200_P2G_ROTATE=0 restores the
    // aligned walk of the round-1 measurements for A/B timing.
    { const char* r = getenv("MPM_B200_P2G_ROTATE"); c.p2g_rotate = (r && atoi(r) == 0) ? 0 : 1; }
    c.pd.h = p.h; c.pd.rh = 1.0f / p.h; c.pd.quadratic = p.stencil == 1 ? 1 : 0;
    // the fast quotient is validated (validate_pos_div) on [2h, (max dim + 2) h]: every position whose particle is not
    // parked anyway lies in there (cell >= 2), see particle_key
    const int md = std::max(s->gd.I, std::max(s->gd.J, s->gd.K));
    c.pd.lo = 2.0f * p.h; c.pd.hi = (float)(md + 2) * p.h;
}

// runs once per handle: exhaustive comparison of the pos/h shortcut with __fdiv_rn over its whole range
static int validate_pos_div(mpm_sim* s) {
    PosDiv& d = s->sc.pd;
    d.fast = 0;
    unsigned lo, hi;
    memcpy(&lo, &d.lo, 4); memcpy(&hi, &d.hi, 4);
    if (!(d.lo > 0.0f) || !(d.hi > d.lo)) return MPM_OK;
    unsigned long long* bad = nullptr;
    CK(cudaMalloc(&bad, sizeof *bad));
    CK(cudaMemsetAsync(bad, 0, sizeof *bad, s->stream));
    k_validate_pos_div<<<s->num_sms * 8, 256, 0, s->stream>>>(d, lo, hi, bad);
    unsigned long long h_bad = 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, sizeof h_bad, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(bad);
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "validate_pos_div: %s", cudaGetErrorString(e));
    d.fast = (h_bad == 0) ? 1 : 0;
    s->stats.kernel_launches++;
    return MPM_OK;
}

static bool valid_variants(const MpmParams& p) {
    return (p.p2g_variant == 0 || p.p2g_variant == 1 || p.p2g_variant == 2 || p.p2g_variant == 9) && (p.g2p_variant == 0 || p.g2p_variant == 1) &&
           p.fupdate_exact >= 0 && p.fupdate_exact <= 2 && (p.stencil == 0 || p.stencil == 1);
}
static int grid_for(int64_t n, int threads) { return (int)std::max<int64_t>(1, (n + threads - 1) / threads); }

static int create_impl(const MpmParams* params, int max_i, int max_j, int max_k, int block_lo, int block_hi,
                       int64_t n_particles, int64_t capacity, mpm_sim*& s);
int mpm_create_slab(const MpmParams* params, int max_i, int max_j, int max_k, int block_lo, int block_hi,
                    int64_t n_particles, int64_t capacity, mpm_t** out) {
    if (!out) return fail(MPM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    mpm_sim* s = nullptr;
    const int rc = create_impl(params, max_i, max_j, max_k, block_lo, block_hi, n_particles, capacity, s);
    if (rc != MPM_OK) {                 // release whatever was allocated before the failure (keeps the error message)
        const std::string msg = g_last_error;
        if (s) mpm_destroy(s);
        g_last_error = msg;
        return rc;
    }
    *out = s;
    return MPM_OK;
}
static int create_impl(const MpmParams* params, int max_i, int max_j, int max_k, int block_lo, int block_hi,
                       int64_t n_particles, int64_t capacity, mpm_sim*& s) {
    if (max_i < 8 || max_j < 8 || max_k < 8) return fail(MPM_ERR_INVALID, "grid must be at least 8^3 nodes");
    if (n_particles < 0 || capacity < n_particles) return fail(MPM_ERR_INVALID, "capacity < n_particles");
    if (capacity >= (int64_t)1 << 31) return fail(MPM_ERR_INVALID, "capacity must be < 2^31");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(MPM_ERR_CUDA, "no CUDA device: libmpm_b200 has no CPU fallback"); }
    s = new mpm_sim();
    CK(cudaGetDevice(&s->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s->device));
    s->num_sms = prop.multiProcessorCount;
    if (params) s->prm = *params; else mpm_default_params(&s->prm);
    GridDims& g = s->gd;
    g.I = max_i; g.J = max_j; g.K = max_k;
    g.npbi_global = (max_i + 3) / 4; g.npbj = (max_j + 3) / 4; g.npbk = (max_k + 3) / 4;
    g.nbj = g.npbj + 1; g.nbk = g.npbk + 1;
    if (block_lo < 0 || block_hi > g.npbi_global || block_lo >= block_hi) return fail(MPM_ERR_INVALID, "bad slab [%d,%d) of %d layers", block_lo, block_hi, g.npbi_global);
    g.lo = block_lo; g.hi = block_hi;
    const int64_t npb = (int64_t)(g.hi - g.lo) * g.npbj * g.npbk, ngb = (int64_t)(g.hi - g.lo + 1) * g.nbj * g.nbk;
    if (npb + 3 >= ((int64_t)1 << 31) / 64) return fail(MPM_ERR_INVALID, "grid too large");
    if (g.hi - g.lo > PB_COORD_MAX || g.npbj > PB_COORD_MAX || g.npbk > PB_COORD_MAX)       // work items carry 10-bit block coordinates
        return fail(MPM_ERR_INVALID, "grid too large: at most %d particle blocks (%d nodes) per axis and slab", PB_COORD_MAX, 4 * PB_COORD_MAX);
    g.n_pblocks = (int)npb; g.n_gblocks = (int)ngb;
    if (!valid_variants(s->prm)) return fail(MPM_ERR_INVALID, "p2g_variant must be 0, 1, 2 or 9, g2p_variant 0 or 1, fupdate_exact 0..2, stencil 0 or 1");
    fill_consts(s);
    s->capacity = std::max<int64_t>(capacity, 1);
    s->n_uploaded = n_particles; s->n_bound = n_particles;
    s->n_buckets = g.n_pblocks + 3;
    s->n_chunks = (s->n_buckets + SCAN_CHUNK - 1) / SCAN_CHUNK;
    CK(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    for (int b = 0; b < 2; ++b) CK(cudaMalloc(&s->buf[b], sizeof(float4) * NPLANES * (size_t)s->capacity));
    CK(cudaMalloc(&s->key, sizeof(int) * (size_t)s->capacity));
    CK(cudaMalloc(&s->sorted_ids, sizeof(int) * (size_t)s->capacity));
    CK(cudaMalloc(&s->blk_count, sizeof(int) * (size_t)s->n_buckets));
    CK(cudaMalloc(&s->blk_start, sizeof(int) * (size_t)s->n_buckets));
    CK(cudaMalloc(&s->blk_cursor, sizeof(int) * (size_t)s->n_buckets));
    CK(cudaMalloc(&s->pblock_list, sizeof(int4) * (size_t)g.n_pblocks));
    CK(cudaMalloc(&s->gflag, sizeof(int) * (size_t)g.n_gblocks));
    CK(cudaMalloc(&s->gblock_list, sizeof(int) * (size_t)g.n_gblocks));
    CK(cudaMalloc(&s->partial, sizeof(int2) * (size_t)s->n_chunks));
    CK(cudaMalloc(&s->grid, sizeof(float4) * 64 * (size_t)g.n_gblocks + PEER_FLAG_BYTES));     // + the peer-halo flag words (same IPC handle)
    CK(cudaMemsetAsync(s->grid + 64 * (size_t)g.n_gblocks, 0, PEER_FLAG_BYTES, s->stream));
    CK(cudaMalloc(&s->dc, sizeof(DevCounters)));
    CK(cudaMemsetAsync(s->gflag, 0, sizeof(int) * (size_t)g.n_gblocks, s->stream));
    CK(cudaMemsetAsync(s->grid, 0, sizeof(float4) * 64 * (size_t)g.n_gblocks, s->stream));
    CK(cudaMemsetAsync(s->dc, 0, sizeof(DevCounters), s->stream));
    for (int b = 0; b < 2; ++b) CK(cudaMemsetAsync(s->buf[b], 0, sizeof(float4) * NPLANES * (size_t)s->capacity, s->stream));
    for (auto& e : s->ev) CK(cudaEventCreate(&e));
    s->ev_ok = true;
    { const char* g = getenv("MPM_B200_GRAPH"); s->graph_enabled = g && atoi(g) > 0; }
    { const char* f = getenv("MPM_B200_FUSE_HIST"); s->hist_fuse = !(f && atoi(f) == 0); }
    {
        // Running the F-update on a side stream next to the gather was measured at <= 1.5 % (the gather's persistent CTAs
        // own the register file), so it is opt-in; the default keeps the two kernels back to back and times them apart.
        const char* ov = getenv("MPM_B200_OVERLAP");
        const int mode = ov ? atoi(ov) : 0;
        CK(cudaEventCreate(&s->side.mid));
        if (mode > 0) {
            CK(cudaStreamCreateWithFlags(&s->side.stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s->side.fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->side.join, cudaEventDisableTiming));
            s->side.gather_ctas_per_sm = mode >= 2 ? 2 : 1;
        }
    }
    memset(&s->stats, 0, sizeof s->stats);
    memset(&s->colliders, 0, sizeof s->colliders);
    CK(tile_kernels_init());
    CK(implicit_kernels_init());
    CK(cudaStreamSynchronize(s->stream));
    { int rc = validate_pos_div(s); if (rc) return rc; }
    return MPM_OK;
}

int mpm_create(const MpmParams* params, int max_i, int max_j, int max_k, int64_t n_particles, mpm_t** out) {
    return mpm_create_slab(params, max_i, max_j, max_k, 0, (max_i + 3) / 4, n_particles, n_particles, out);
}

int mpm_destroy(mpm_t* s) {
    if (!s) return MPM_OK;
    cudaSetDevice(s->device);           // the handle's allocations live on the device it was created on
    cudaStreamSynchronize(s->stream);
    for (int b = 0; b < 2; ++b) { cudaFree(s->buf[b]); cudaFree(s->out_buf[b]); }
    cudaFree(s->key); cudaFree(s->sorted_ids); cudaFree(s->blk_count); cudaFree(s->blk_start); cudaFree(s->blk_cursor);
    cudaFree(s->pblock_list); cudaFree(s->gflag); cudaFree(s->gblock_list); cudaFree(s->partial);
    cudaFree(s->grid); cudaFree(s->gforce);
    for (auto& v : s->imp_vec) cudaFree(v);
    cudaFree(s->imp_acc); cudaFree(s->imp_aux); cudaFree(s->dc); cudaFree(s->slot_of_pid);
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->ev_ok) for (auto& e : s->ev) cudaEventDestroy(e);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); cudaEventDestroy(s->render_ready); cudaEventDestroy(s->copy_done); }
    cudaFree(s->render_stage);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    if (s->ipc_dn) cudaIpcCloseMemHandle(s->ipc_dn);
    if (s->ipc_up) cudaIpcCloseMemHandle(s->ipc_up);
    if (s->ipc_mig_dn) cudaIpcCloseMemHandle(s->ipc_mig_dn);
    if (s->ipc_mig_up) cudaIpcCloseMemHandle(s->ipc_mig_up);
    if (s->side.stream) { cudaStreamDestroy(s->side.stream); cudaEventDestroy(s->side.fork); cudaEventDestroy(s->side.join); }
    if (s->side.mid) cudaEventDestroy(s->side.mid);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return MPM_OK;
}

static void drop_graph(mpm_sim* s) {     // a captured substep pair bakes in the stream, SimConst and the kernel variants
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
}
int mpm_set_stream(mpm_t* s, void* st) {
    NEED(s);
    CK(cudaStreamSynchronize(s->stream));
    s->stream = st ? (cudaStream_t)st : s->own_stream;
    drop_graph(s);
    return MPM_OK;
}

int mpm_set_params(mpm_t* s, const MpmParams* p) {
    if (!s || !p) return fail(MPM_ERR_INVALID, "null argument");
    if (p->h != s->prm.h || p->stencil != s->prm.stencil) return fail(MPM_ERR_INVALID, "h and stencil cannot change after creation");
    if (!valid_variants(*p)) return fail(MPM_ERR_INVALID, "p2g_variant must be 0, 1, 2 or 9, g2p_variant 0 or 1, fupdate_exact 0..2, stencil 0 or 1");
    s->prm = *p;
    drop_graph(s);                  // material constants and kernel variants are baked into a captured substep pair
    const int fast = s->sc.pd.fast;
    fill_consts(s);
    s->sc.pd.fast = fast;           // h is unchanged, so the validation still holds
    s->tau_valid = false;
    return MPM_OK;
}

static int ensure_pinned(mpm_sim* s, size_t bytes) {
    if (s->pinned_bytes >= bytes) return MPM_OK;
    if (s->pinned) cudaFreeHost(s->pinned);
    s->pinned = nullptr; s->pinned_bytes = 0;
    CK(cudaMallocHost(&s->pinned, bytes));
    s->pinned_bytes = bytes;
    return MPM_OK;
}

// ---- upload / download -----------------------------------------------------------------------------------
struct HostFieldPtrs {   // strided views over the caller's particle memory (AoS or SoA), NULL = default
    const char *mass, *vel, *vol, *pos, *FE, *FP, *B;
    size_t s_mass, s_vel, s_vol, s_pos, s_FE, s_FP, s_B;
};
static const float ID9[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
static const float Z9[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };

static int upload_common(mpm_sim* s, int64_t n, const HostFieldPtrs& f) {
    NEED(s);
    if (n < 0 || n > s->capacity) return fail(MPM_ERR_CAPACITY, "n = %lld exceeds capacity %lld", (long long)n, (long long)s->capacity);
    if (n > 0 && (!f.pos || !f.vel || !f.mass)) return fail(MPM_ERR_INVALID, "pos, vel and mass are required");
    const int64_t CH = 1 << 20;
    int rc = ensure_pinned(s, sizeof(float4) * NPLANES * (size_t)std::min<int64_t>(CH, std::max<int64_t>(n, 1)));
    if (rc) return rc;
    Planes P = s->planes(s->cur);
    for (int64_t base = 0; base < n; base += CH) {
        const int64_t m = std::min(CH, n - base);
        CK(cudaStreamSynchronize(s->stream));      // staging buffer reuse
        float4* st = (float4*)s->pinned;
        for (int64_t i = 0; i < m; ++i) {
            const int64_t p = base + i;
            const float* pos = (const float*)(f.pos + p * f.s_pos);
            const float* vel = (const float*)(f.vel + p * f.s_vel);
            const float mass = *(const float*)(f.mass + p * f.s_mass);
            const float vol = f.vol ? *(const float*)(f.vol + p * f.s_vol) : 0.0f;
            const float* FE = f.FE ? (const float*)(f.FE + p * f.s_FE) : ID9;
            const float* FP = f.FP ? (const float*)(f.FP + p * f.s_FP) : ID9;
            const float* B = f.B ? (const float*)(f.B + p * f.s_B) : Z9;
            int pid = (int)(p + s->pid_base); float pidf; memcpy(&pidf, &pid, 4);
            st[0 * m + i] = make_float4(pos[0], pos[1], pos[2], mass);
            st[1 * m + i] = make_float4(B[0], B[1], B[2], B[3]);
            st[2 * m + i] = make_float4(B[4], B[5], B[6], B[7]);
            st[3 * m + i] = make_float4(B[8], vel[0], vel[1], vel[2]);
            st[4 * m + i] = make_float4(0, 0, 0, 0);
            st[5 * m + i] = make_float4(0, 0, 0, 0);
            st[6 * m + i] = make_float4(vol, pidf, FE[0], FE[1]);
            st[7 * m + i] = make_float4(FE[2], FE[3], FE[4], FE[5]);
            st[8 * m + i] = make_float4(FE[6], FE[7], FE[8], FP[0]);
            st[9 * m + i] = make_float4(FP[1], FP[2], FP[3], FP[4]);
            st[10 * m + i] = make_float4(FP[5], FP[6], FP[7], FP[8]);
        }
        for (int k = 0; k < NPLANES; ++k)
            CK(cudaMemcpyAsync(P.p[k] + base, st + (size_t)k * m, sizeof(float4) * m, cudaMemcpyHostToDevice, s->stream));
    }
    const int ni = (int)n;
    CK(cudaMemcpyAsync(&s->dc->n_slots, &ni, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->n_uploaded = n; s->n_bound = n;
    s->tau_valid = false; s->binned = false; s->hist_valid = false;
    return MPM_OK;
}

struct HostFieldPtrsW { char *mass, *vel, *vol, *pos, *FE, *FP, *B; size_t s_mass, s_vel, s_vol, s_pos, s_FE, s_FP, s_B; };

static int download_common(mpm_sim* s, int64_t n, const HostFieldPtrsW& f) {
    NEED(s);
    if (s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global)
        return fail(MPM_ERR_INVALID, "slab handles exchange particles: use mpm_download_live_particles");
    if (n != s->n_uploaded) return fail(MPM_ERR_INVALID, "download of %lld particles but %lld were uploaded", (long long)n, (long long)s->n_uploaded);
    if (s->fupd_pending) return fail(MPM_ERR_INVALID, "mpm_substep_begin is pending: call mpm_substep_end before downloading (the idle buffer is in use)");
    const int64_t CH = 1 << 20;
    int rc = ensure_pinned(s, sizeof(float4) * NPLANES * (size_t)std::min<int64_t>(CH, std::max<int64_t>(n, 1)));
    if (rc) return rc;
    Planes C = s->planes(s->cur), O = s->planes(s->cur ^ 1);    // the idle buffer is the staging area
    k_unsort<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(C, O, s->dc);
    CKLAUNCH(); s->stats.kernel_launches++;
    for (int64_t base = 0; base < n; base += CH) {
        const int64_t m = std::min(CH, n - base);
        float4* st = (float4*)s->pinned;
        for (int k = 0; k < NPLANES; ++k)
            CK(cudaMemcpyAsync(st + (size_t)k * m, O.p[k] + base, sizeof(float4) * m, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        for (int64_t i = 0; i < m; ++i) {
            const int64_t p = base + i;
            const float4 a0 = st[0 * m + i], a1 = st[1 * m + i], a2 = st[2 * m + i], a3 = st[3 * m + i], a6 = st[6 * m + i],
                         a7 = st[7 * m + i], a8 = st[8 * m + i], a9 = st[9 * m + i], a10 = st[10 * m + i];
            if (f.pos) { float* o = (float*)(f.pos + p * f.s_pos); o[0] = a0.x; o[1] = a0.y; o[2] = a0.z; }
            if (f.mass) *(float*)(f.mass + p * f.s_mass) = a0.w;
            if (f.vel) { float* o = (float*)(f.vel + p * f.s_vel); o[0] = a3.y; o[1] = a3.z; o[2] = a3.w; }
            if (f.vol) *(float*)(f.vol + p * f.s_vol) = a6.x;
            if (f.B) { float* o = (float*)(f.B + p * f.s_B); o[0] = a1.x; o[1] = a1.y; o[2] = a1.z; o[3] = a1.w; o[4] = a2.x; o[5] = a2.y; o[6] = a2.z; o[7] = a2.w; o[8] = a3.x; }
            if (f.FE) { float* o = (float*)(f.FE + p * f.s_FE); o[0] = a6.z; o[1] = a6.w; o[2] = a7.x; o[3] = a7.y; o[4] = a7.z; o[5] = a7.w; o[6] = a8.x; o[7] = a8.y; o[8] = a8.z; }
            if (f.FP) { float* o = (float*)(f.FP + p * f.s_FP); o[0] = a8.w; o[1] = a9.x; o[2] = a9.y; o[3] = a9.z; o[4] = a9.w; o[5] = a10.x; o[6] = a10.y; o[7] = a10.z; o[8] = a10.w; }
        }
    }
    return MPM_OK;
}

int mpm_upload_particles_aos(mpm_t* s, const void* particles, int64_t n, size_t stride, size_t off_mass, size_t off_velocity,
                             size_t off_volume, size_t off_pos, size_t off_FE, size_t off_FP, size_t off_B) {
    if (!s || (!particles && n > 0)) return fail(MPM_ERR_INVALID, "null argument");
    const char* b = (const char*)particles;
    HostFieldPtrs f = { b + off_mass, b + off_velocity, b + off_volume, b + off_pos, b + off_FE, b + off_FP, b + off_B,
                        stride, stride, stride, stride, stride, stride, stride };
    return upload_common(s, n, f);
}
int mpm_upload_particles_soa(mpm_t* s, int64_t n, const float* pos, const float* vel, const float* mass, const float* volume,
                             const float* FE, const float* FP, const float* B) {
    if (!s) return fail(MPM_ERR_INVALID, "null handle");
    HostFieldPtrs f = { (const char*)mass, (const char*)vel, (const char*)volume, (const char*)pos, (const char*)FE, (const char*)FP, (const char*)B,
                        4, 12, 4, 12, 36, 36, 36 };
    return upload_common(s, n, f);
}
int mpm_download_particles_aos(mpm_t* s, void* particles, int64_t n, size_t stride, size_t off_mass, size_t off_velocity,
                               size_t off_volume, size_t off_pos, size_t off_FE, size_t off_FP, size_t off_B) {
    if (!s || (!particles && n > 0)) return fail(MPM_ERR_INVALID, "null argument");
    char* b = (char*)particles;
    HostFieldPtrsW f = { b + off_mass, b + off_velocity, b + off_volume, b + off_pos, b + off_FE, b + off_FP, b + off_B,
                         stride, stride, stride, stride, stride, stride, stride };
    return download_common(s, n, f);
}
int mpm_download_particles_soa(mpm_t* s, int64_t n, float* pos, float* vel, float* mass, float* volume, float* FE, float* FP, float* B) {
    if (!s) return fail(MPM_ERR_INVALID, "null handle");
    HostFieldPtrsW f = { (char*)mass, (char*)vel, (char*)volume, (char*)pos, (char*)FE, (char*)FP, (char*)B, 4, 12, 4, 12, 36, 36, 36 };
    return download_common(s, n, f);
}

int mpm_download_render_buffers(mpm_t* s, int64_t n, float* xyzs, unsigned char* rgba, float size) {
    NEED(s);
    const bool slab = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
    if (!slab && n != s->n_uploaded) return fail(MPM_ERR_INVALID, "n mismatch");
    if (slab && (n < 0 || n > s->capacity)) return fail(MPM_ERR_INVALID, "n exceeds the slab capacity");
    if (xyzs) {
        float4* stage = s->buf[s->cur ^ 1];     // plane 0 of the idle buffer
        if (slab) k_render_slots<<<grid_for(n, 256), 256, 0, s->stream>>>(s->planes(s->cur), stage, s->dc, size, (int)n);
        else k_render<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), stage, s->dc, size);
        CKLAUNCH(); s->stats.kernel_launches++;
        CK(cudaMemcpyAsync(xyzs, stage, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    if (rgba) memset(rgba, 255, (size_t)n * 4);     // initializeParticles sets r=g=b=a=255 (cpp:48-51)
    return MPM_OK;
}

int mpm_write_render_buffers_device(mpm_t* s, int64_t n, void* d_xyzs, void* d_rgba, float size) {
    NEED(s);
    const bool slab = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
    if (!slab && n != s->n_uploaded) return fail(MPM_ERR_INVALID, "n mismatch");
    if (slab && (n < 0 || n > s->capacity)) return fail(MPM_ERR_INVALID, "n exceeds the slab capacity");
    if (n == 0) return MPM_OK;
    if (d_xyzs) {
        if (slab) k_render_slots<<<grid_for(n, 256), 256, 0, s->stream>>>(s->planes(s->cur), (float4*)d_xyzs, s->dc, size, (int)n);
        else k_render<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), (float4*)d_xyzs, s->dc, size);
        CKLAUNCH(); s->stats.kernel_launches++;
    }
    if (d_rgba) CK(cudaMemsetAsync(d_rgba, 255, (size_t)n * 4, s->stream));     // initializeParticles: r = g = b = a = 255 (cpp:48-51)
    return MPM_OK;
}

int mpm_download_render_buffers_async(mpm_t* s, int64_t n, float* xyzs, float size) {
    NEED(s);
    if (!xyzs || n < 0) return fail(MPM_ERR_INVALID, "bad argument");
    const bool slab = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
    if (!slab && n != s->n_uploaded) return fail(MPM_ERR_INVALID, "n mismatch");
    if (slab && n > s->capacity) return fail(MPM_ERR_INVALID, "n exceeds the slab capacity");
    if (!s->copy_stream) {
        CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->render_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming));
    }
    if (s->render_cap < n) {
        if (s->copy_pending) CK(cudaEventSynchronize(s->copy_done));
        cudaFree(s->render_stage);
        s->render_stage = nullptr;
        CK(cudaMalloc(&s->render_stage, sizeof(float4) * (size_t)std::max<int64_t>(n, 1)));
        s->render_cap = n;
    }
    if (n == 0) return MPM_OK;
    // the staging buffer is reused: the previous copy must have drained before it is overwritten (device-side wait)
    if (s->copy_pending) CK(cudaStreamWaitEvent(s->stream, s->copy_done, 0));
    if (slab) k_render_slots<<<grid_for(n, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->render_stage, s->dc, size, (int)n);
    else k_render<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->render_stage, s->dc, size);
    CKLAUNCH(); s->stats.kernel_launches++;
    CK(cudaEventRecord(s->render_ready, s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, s->render_ready, 0));
    CK(cudaMemcpyAsync(xyzs, s->render_stage, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->copy_done, s->copy_stream));
    s->copy_pending = true;
    return MPM_OK;
}
int mpm_wait_render_buffers(mpm_t* s) {
    NEED(s);
    if (s->copy_pending) { CK(cudaEventSynchronize(s->copy_done)); s->copy_pending = false; }
    return MPM_OK;
}

// ---- binning / sort -----------------------------------------------------------------------------------------
static int do_binning(mpm_sim* s) {
    const GridDims& g = s->gd;
    const int nb = grid_for(s->n_bound, BIN_T * BIN_E);
    Planes C = s->planes(s->cur);
    if (!s->hist_valid) {      // (the fused substep's gather has already written next substep's keys and histogram)
        CK(cudaMemsetAsync(s->blk_count, 0, sizeof(int) * (size_t)s->n_buckets, s->stream));
        k_bin_count<<<nb, BIN_T, 0, s->stream>>>(C.p[0], (int)s->n_bound, s->dc, g, s->sc.pd, s->key, s->blk_count);
        CKLAUNCH(); s->stats.kernel_launches++;
    }
    s->hist_valid = false;
    k_scan_reduce<<<s->n_chunks, SCAN_T, 0, s->stream>>>(s->blk_count, s->n_buckets, g.n_pblocks, s->partial);
    CKLAUNCH();
    k_scan_partials<<<1, 1024, 0, s->stream>>>(s->partial, s->n_chunks, s->dc, s->blk_count, g.n_pblocks);
    CKLAUNCH();
    k_scan_apply<<<s->n_chunks, SCAN_T, 0, s->stream>>>(s->blk_count, s->n_buckets, g.n_pblocks, s->partial, s->blk_start,
                                                      s->blk_cursor, s->pblock_list, s->gflag, s->gblock_list, s->dc, g);
    CKLAUNCH();
    s->stats.kernel_launches += 3;
    const int layer_threads = g.nbj * g.nbk;
    if (g.lo > 0) { k_mark_layer<<<grid_for(layer_threads, 256), 256, 0, s->stream>>>(0, g, s->gflag, s->gblock_list, s->dc); CKLAUNCH(); s->stats.kernel_launches++; }
    if (g.hi < g.npbi_global) { k_mark_layer<<<grid_for(layer_threads, 256), 256, 0, s->stream>>>(g.hi - g.lo, g, s->gflag, s->gblock_list, s->dc); CKLAUNCH(); s->stats.kernel_launches++; }
    k_bin_scatter<<<nb, BIN_T, 0, s->stream>>>((int)s->n_bound, s->dc, s->key, s->blk_cursor, s->sorted_ids);
    CKLAUNCH(); s->stats.kernel_launches++;
    s->binned = true;
    return MPM_OK;
}

static int ensure_tau(mpm_sim* s) {
    if (s->tau_valid) return MPM_OK;
    k_stress<<<grid_for(s->n_bound, 128), 128, 0, s->stream>>>(s->planes(s->cur), s->dc, s->sc);
    CKLAUNCH(); s->stats.kernel_launches++;
    s->tau_valid = true;
    return MPM_OK;
}
static int ensure_gforce(mpm_sim* s) {
    if (s->gforce) return MPM_OK;
    CK(cudaMalloc(&s->gforce, sizeof(float4) * 64 * (size_t)s->gd.n_gblocks));
    CK(cudaMemsetAsync(s->gforce, 0, sizeof(float4) * 64 * (size_t)s->gd.n_gblocks, s->stream));
    return MPM_OK;
}
static int persistent_grid(const mpm_sim* s, int per_sm) { return s->num_sms * per_sm; }

static int launch_clear(mpm_sim* s) {
    k_grid_clear<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->grid, s->gforce);
    CKLAUNCH(); s->stats.kernel_launches++;
    return MPM_OK;
}
// p2g_variant 0 (auto): the fused substep runs the F-update inside the P2G kernel; 2: as a kernel of its own (k_fupdate), which is
// also what the side-stream overlap (MPM_B200_OVERLAP) needs
static bool p2g_fupd(const mpm_sim* s) { return s->prm.p2g_variant == 0 && s->prm.g2p_variant == 0 && !s->side.stream; }
template <int MODE>
static int launch_p2g(mpm_sim* s, float4* target, float dt) {
    if (s->prm.p2g_variant == 1) {
        k_p2g_atomic<MODE><<<grid_for(s->n_bound, 128), 128, 0, s->stream>>>(s->planes(s->cur), s->sorted_ids, s->dc, target, s->gd, s->sc, dt);
        CKLAUNCH();
    } else if (s->prm.p2g_variant == 9) {
        // deterministic debug mode: one thread, ascending particle id, plain additions (bitwise reproducible run to run)
        const bool slab = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
        if (slab) return fail(MPM_ERR_INVALID, "p2g_variant 9 (deterministic debug mode) is for single-domain handles");
        const int n_pid = (int)s->n_uploaded;
        if (!s->slot_of_pid) CK(cudaMalloc(&s->slot_of_pid, sizeof(int) * (size_t)std::max<int64_t>(s->capacity, 1)));
        CK(cudaMemsetAsync(s->slot_of_pid, 0xff, sizeof(int) * (size_t)std::max(n_pid, 1), s->stream));       // -1: not binned
        k_slot_of_pid<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->key, s->dc, s->gd, s->slot_of_pid, n_pid);
        CKLAUNCH();
        k_p2g_serial<MODE><<<1, 1, 0, s->stream>>>(s->planes(s->cur), s->slot_of_pid, n_pid, target, s->gd, s->sc, dt);
        CKLAUNCH(); s->stats.kernel_launches++;
    } else if (MODE == P2G_FUSED && p2g_fupd(s)) {
        const Planes nxt = s->planes(s->cur ^ 1);
        CK((launch_p2g_tile<MODE>(s->planes(s->cur), s->sorted_ids, s->pblock_list, s->dc, target, s->gd, s->sc, dt,
                                  s->num_sms, (int)s->n_bound, s->stream, &nxt, s->prm.fupdate_exact != 1)));
    } else {
        CK((launch_p2g_tile<MODE>(s->planes(s->cur), s->sorted_ids, s->pblock_list, s->dc, target, s->gd, s->sc, dt,
                                  s->num_sms, (int)s->n_bound, s->stream)));
    }
    s->stats.kernel_launches++;
    return MPM_OK;
}
template <int FLAGS>
static int launch_grid_update(mpm_sim* s, float dt) {
    k_grid_update<FLAGS><<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->grid, s->gforce, s->gd, s->sc, dt,
                                                                     s->colliders, s->n_colliders);
    CKLAUNCH(); s->stats.kernel_launches++;
    return MPM_OK;
}
static int ensure_out_buffers(mpm_sim* s);
__global__ void k_after_reorder(DevCounters* dc) { dc->n_slots = dc->n_sorted; }      // baseline gather: the re-sorted buffer holds n_sorted contiguous slots
__global__ void k_peer_wait(const int* flag_a, const int* flag_b, int epoch, DevCounters* dc);
template <int FLAGS>
static int launch_g2p(mpm_sim* s, float dt) {
    Planes C = s->planes(s->cur), N = s->planes(s->cur ^ 1);
    // Next substep's keys + histogram are produced inside the gather, where the advected position is in registers. On a slab
    // handle the same pass packs the particles that left the slab into the migration buffers and retires their slots, so
    // that the particle set the histogram describes is the one the next binning sees (incoming particles add their own keys
    // in k_append_incoming_hdr). Not with the side-stream F-update: a packed record needs this substep's planes 4..10.
    const bool slab_handle = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
    const bool fuse_hist = s->hist_fuse && s->prm.g2p_variant != 1 && !(slab_handle && s->side.stream) &&
                           (FLAGS & G2P_REORDER) && (FLAGS & G2P_ADVECT) && (FLAGS & G2P_GATHER);
    if ((FLAGS & G2P_ADVECT) && !fuse_hist) s->hist_valid = false;       // positions change without new keys
    MigOut mo = { nullptr, nullptr, 0 };
    if (fuse_hist) {
        CK(cudaMemsetAsync(s->blk_count, 0, sizeof(int) * (size_t)s->n_buckets, s->stream));
        if (slab_handle) {
            TRY(ensure_out_buffers(s));
            if (s->peer_mig_connected) {      // the neighbours must have consumed the buffers of the previous substep before they are refilled
                int* mine = (int*)(s->grid + 64 * (size_t)s->gd.n_gblocks);
                k_peer_wait<<<1, 1, 0, s->stream>>>(s->gd.lo > 0 ? mine + 6 : nullptr, s->gd.hi < s->gd.npbi_global ? mine + 7 : nullptr, s->peer_mig_epoch, s->dc);
                CKLAUNCH(); s->stats.kernel_launches++;
            }
            for (int d = 0; d < 2; ++d) CK(cudaMemsetAsync(s->out_buf[d], 0, sizeof(float4), s->stream));
            mo.dn = s->gd.lo > 0 ? s->out_buf[0] : nullptr;
            mo.up = s->gd.hi < s->gd.npbi_global ? s->out_buf[1] : nullptr;
            mo.cap = (int)s->out_cap;
        }
    }
    if (s->prm.g2p_variant == 1 || !(FLAGS & (G2P_GATHER | G2P_F))) {
        k_g2p_direct<FLAGS><<<grid_for(s->n_bound, 128), 128, 0, s->stream>>>(C, N, s->sorted_ids, s->dc, s->grid, s->gd, s->sc, dt);
        CKLAUNCH();
    } else {
        CK((launch_g2p_tile<FLAGS>(C, N, s->sorted_ids, s->pblock_list, s->dc, s->grid, s->gd, s->sc, dt,
                                   s->num_sms, (int)s->n_bound, s->stream, &s->side,
                                   s->prm.fupdate_exact == 2 || ((FLAGS & G2P_REORDER) && s->prm.fupdate_exact == 0),      // tolerance-form F-update: fused substep (or forced)
                                   fuse_hist ? s->key : nullptr, fuse_hist ? s->blk_count : nullptr, mo)));
    }
    s->stats.kernel_launches += (s->prm.g2p_variant == 1) ? 1 : ((FLAGS & G2P_F) ? 1 : 0) + ((FLAGS & G2P_GATHER) ? 1 : 0) + ((FLAGS & G2P_REORDER) ? 1 : 0);
    if (FLAGS & G2P_REORDER) {
        if (s->prm.g2p_variant == 1) {      // (the tile path sets the slot count in k_copy_parked)
            k_after_reorder<<<1, 1, 0, s->stream>>>(s->dc);
            CKLAUNCH(); s->stats.kernel_launches++;
        }
        s->cur ^= 1;
        s->binned = false;      // sorted_ids referred to the old buffer
        s->hist_valid = fuse_hist;
        s->mig_packed = fuse_hist && slab_handle;
    }
    return MPM_OK;
}

static int set_colliders(mpm_sim* s, const MpmBoxCollider* c, int n) {
    if (n < 0 || n > MPM_MAX_COLLIDERS) return fail(MPM_ERR_INVALID, "n_colliders must be in [0, %d]", MPM_MAX_COLLIDERS);
    if (n > 0 && !c) return fail(MPM_ERR_INVALID, "colliders is NULL");
    static_assert(sizeof(MpmBoxCollider) == sizeof(BoxCollider), "collider POD mismatch");
    if (n) memcpy(s->colliders.c, c, sizeof(BoxCollider) * n);
    s->n_colliders = n;
    return MPM_OK;
}

// ---- staged API: one entry per reference stage ----------------------------------------------------------------
int mpm_rasterize_particles_to_grid(mpm_t* s) {               // cpp:94-129
    NEED(s);
    TRY(do_binning(s));
    TRY(launch_clear(s));
    TRY((launch_p2g<P2G_MOMENTUM>(s, s->grid, 0.0f)));
    TRY((launch_grid_update<GU_NORMALIZE | GU_COUNT>(s, 0.0f)));
    return MPM_OK;
}
int mpm_compute_particle_volumes_and_densities(mpm_t* s) {     // cpp:131-142
    NEED(s);
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first (the reference reuses its neighbour lists)");
    k_volumes<<<grid_for(s->n_bound, 128), 128, 0, s->stream>>>(s->planes(s->cur), s->sorted_ids, s->dc, s->grid, s->gd, s->sc);
    CKLAUNCH(); s->stats.kernel_launches++;
    s->tau_valid = false;
    return MPM_OK;
}
int mpm_compute_explicit_grid_forces(mpm_t* s) {               // cpp:235-254
    NEED(s);
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first");
    const bool fresh = !s->gforce;
    TRY(ensure_gforce(s));
    (void)fresh;
    TRY(ensure_tau(s));
    TRY((launch_p2g<P2G_FORCE>(s, s->gforce, 0.0f)));
    return MPM_OK;
}
int mpm_grid_velocities_update(mpm_t* s, float dt) {           // cpp:256-262
    NEED(s);
    TRY(ensure_gforce(s));
    TRY((launch_grid_update<GU_FORCE | GU_GRAVITY>(s, dt)));
    return MPM_OK;
}
int mpm_grid_based_collisions(mpm_t* s, float dt, const MpmBoxCollider* c, int n) {   // cpp:264-304
    NEED(s);
    TRY(set_colliders(s, c, n));
    TRY((launch_grid_update<GU_COLLIDE>(s, dt)));
    return MPM_OK;
}
int mpm_update_deformation_gradient(mpm_t* s, float dt) {      // cpp:306-330
    NEED(s);
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first");
    TRY((launch_g2p<G2P_F>(s, dt)));
    s->tau_valid = true;
    return MPM_OK;
}
int mpm_update_particle_velocities(mpm_t* s) {                 // cpp:332-342
    NEED(s);
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first");
    TRY((launch_g2p<G2P_GATHER>(s, 0.0f)));
    return MPM_OK;
}
int mpm_update_particle_positions(mpm_t* s, float dt) {        // cpp:344-350
    NEED(s);
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first");
    TRY((launch_g2p<G2P_ADVECT>(s, dt)));
    return MPM_OK;
}

// ---- implicit (optimisation-based) time integration: LagrangeEulerView::timeIntegration, cpp:211-233 ------------------------
void mpm_default_implicit_params(MpmImplicitParams* q) {
    memset(q, 0, sizeof *q);
    q->mu0 = 1.0f; q->lambda0 = 1.0f; q->xi = 10.0f;                       // hpp:212-214
    q->hardening = 0;                                                      // exp(xi*1 - det FP), cpp:187-191 as written
    q->max_iters = 50;                                                     // LBFGS.hpp:48
    q->ls_decrease = 1e-4f; q->ls_tau = 0.7f; q->ls_max_iters = 100000;    // Minimizer.hpp:62-63, Backtracking.hpp:46
    q->tol_grad = 1e-2f; q->tol_step = 1e-2f;                              // mathy.hpp:31-35
}
static int implicit_ready(mpm_sim* s, const MpmImplicitParams* q, int n_vec) {
    if (!q) return fail(MPM_ERR_INVALID, "params is NULL");
    if (!s->binned) return fail(MPM_ERR_INVALID, "call mpm_rasterize_particles_to_grid first (the objective is defined on its grid and used cells)");
    if (s->prm.stencil != 0) return fail(MPM_ERR_INVALID, "implicit time integration is defined for the reference's cubic stencil only");
    if (s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global) return fail(MPM_ERR_INVALID, "implicit time integration: single-domain handles only");
    if (!(q->mu0 >= 0.0f) || !(q->lambda0 >= 0.0f) || q->max_iters < 0 || q->ls_max_iters < 1 || !(q->ls_tau > 0.0f && q->ls_tau < 1.0f) || (q->hardening != 0 && q->hardening != 1))
        return fail(MPM_ERR_INVALID, "bad implicit parameters");
    const size_t bytes = sizeof(float4) * 64 * (size_t)s->gd.n_gblocks;
    for (int i = 0; i < n_vec; ++i)
        if (!s->imp_vec[i]) {
            CK(cudaMalloc(&s->imp_vec[i], bytes));
            CK(cudaMemsetAsync(s->imp_vec[i], 0, bytes, s->stream));
        }
    if (!s->imp_acc) { CK(cudaMalloc(&s->imp_acc, 32 * sizeof(double))); CK(cudaMemsetAsync(s->imp_acc, 0, 32 * sizeof(double), s->stream)); }
    if (!s->imp_aux) CK(cudaMalloc(&s->imp_aux, sizeof(float4) * 3 * (size_t)s->capacity));
    return MPM_OK;
}
static ImplicitConst implicit_const(const MpmImplicitParams* q) { return ImplicitConst{ q->mu0, q->lambda0, q->xi, q->hardening }; }
// device scalars of the implicit solve (doubles): [0..1] energies of a line-search evaluation, [2..3] energies of the
// value + gradient evaluation, [4..6] dot products read back together, [7] |x|_inf bits, [8] the recursion's running dot product;
// floats at (float*)(imp_acc + 16): alpha[8] and the current coefficient
enum { ACC_LS = 0, ACC_VG = 2, ACC_DOT = 4, ACC_ABS = 7, ACC_RUN = 8, ACC_FLOATS = 16, ACC_DOUBLES = 32 };
// E(x) (and its gradient into vector g when g >= 0) accumulated into imp_acc[slot], imp_acc[slot + 1]; no host read-back
static int implicit_eval_async(mpm_sim* s, float dt, const MpmImplicitParams* q, int x, int g, int slot) {
    CK(cudaMemsetAsync(s->imp_acc + slot, 0, 2 * sizeof(double), s->stream));
    double* acc = s->imp_acc + slot;
    float4* G = g >= 0 ? s->imp_vec[g] : nullptr;
    k_imp_nodes<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->grid, s->imp_vec[x], G, acc);
    CKLAUNCH();
    const int nb = grid_for(s->n_bound, 128);
    if (s->prm.p2g_variant == 1 || s->prm.g2p_variant == 1) {       // baseline: thread per particle, 64 gathers and 64 vector reds each
        if (G) k_imp_particles<true><<<nb, 128, 0, s->stream>>>(s->planes(s->cur), s->sorted_ids, s->dc, s->imp_vec[x], G, s->gd, s->sc, dt, implicit_const(q), acc);
        else k_imp_particles<false><<<nb, 128, 0, s->stream>>>(s->planes(s->cur), s->sorted_ids, s->dc, s->imp_vec[x], nullptr, s->gd, s->sc, dt, implicit_const(q), acc);
        CKLAUNCH();
        s->stats.kernel_launches += 2;
    } else {                                                        // block tiles: TMA-staged gather | per-particle stress | register-accumulated scatter
        const Planes C = s->planes(s->cur);
        CK(cudaMemsetAsync(&s->dc->work_b, 0, sizeof(int), s->stream));
        k_g2p_tile<G2P_GATHER | G2P_GRADW, 4><<<s->num_sms * G2P_MIN_CTAS, G2P_T, sizeof(G2PSmem), s->stream>>>(C, C, s->sorted_ids, s->pblock_list, s->dc, s->imp_vec[x], s->gd, s->sc, dt,
                                                                                                         nullptr, nullptr, MigOut{ nullptr, nullptr, 0 }, s->imp_aux);
        CKLAUNCH();
        if (G) k_imp_stress<true><<<nb, 128, 0, s->stream>>>(C, s->sorted_ids, s->dc, s->imp_aux, dt, implicit_const(q), acc);
        else k_imp_stress<false><<<nb, 128, 0, s->stream>>>(C, s->sorted_ids, s->dc, s->imp_aux, dt, implicit_const(q), acc);
        CKLAUNCH();
        s->stats.kernel_launches += 3;
        if (G) {
            CK(cudaMemsetAsync(&s->dc->work_a, 0, sizeof(int), s->stream));
            k_imp_scatter_tile<<<s->num_sms * 2, P2G_T, sizeof(ImpScatterSmem), s->stream>>>(C, s->sorted_ids, s->pblock_list, s->dc, s->imp_aux, G, s->gd, s->sc);
            CKLAUNCH();
            s->stats.kernel_launches++;
        }
    }
    return MPM_OK;
}
static int implicit_read(mpm_sim* s, int first, int count, double* h) {      // the one place the solve waits for the device
    CK(cudaMemcpyAsync(h, s->imp_acc + first, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return MPM_OK;
}
static int implicit_eval(mpm_sim* s, float dt, const MpmImplicitParams* q, int x, int g, double* inertia, double* elastic) {
    TRY(implicit_eval_async(s, dt, q, x, g, ACC_LS));
    double h[2];
    TRY(implicit_read(s, ACC_LS, 2, h));
    *inertia = h[0]; *elastic = h[1];
    return MPM_OK;
}
// imp_acc[slot] += a . b (the slot must be zero); optionally |a|_inf into ACC_ABS
static int implicit_dot_async(mpm_sim* s, int a, int b, int slot, bool absmax_a = false) {
    k_vec_dot<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->imp_vec[a], s->imp_vec[b], s->imp_acc + slot, absmax_a ? (int*)(s->imp_acc + ACC_ABS) : nullptr);
    CKLAUNCH(); s->stats.kernel_launches++;
    return MPM_OK;
}
static int implicit_lin(mpm_sim* s, int z, float a, int x, float b, int y, bool b_on_device = false) {      // z = a x + b y (y < 0: z = a x)
    k_vec_lin<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->imp_vec[z], a, s->imp_vec[x], b, y >= 0 ? s->imp_vec[y] : nullptr,
                                                           b_on_device ? (const float*)(s->imp_acc + ACC_FLOATS) + 8 : nullptr);
    CKLAUNCH(); s->stats.kernel_launches++;
    return MPM_OK;
}
static int implicit_load_trial(mpm_sim* s, const float* trial_velocity, int add) {
    const size_t n = (size_t)s->gd.I * s->gd.J * s->gd.K;
    if (!trial_velocity) {
        k_imp_from_grid<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->grid, s->imp_vec[mpm_sim::IMP_X]);
        CKLAUNCH(); s->stats.kernel_launches++;
        return MPM_OK;
    }
    float* d = nullptr;
    CK(cudaMalloc(&d, n * 3 * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d, trial_velocity, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
        k_imp_import<<<grid_for((int64_t)n, 256), 256, 0, s->stream>>>(s->imp_vec[mpm_sim::IMP_X], s->grid, s->gd, d, add);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "implicit: trial velocity upload: %s", cudaGetErrorString(e));
    return MPM_OK;
}
int mpm_energy(mpm_t* s, float dt, const MpmImplicitParams* q, const float* trial_velocity, int relative, double* energy, double* elastic) {
    NEED(s);
    if (!energy) return fail(MPM_ERR_INVALID, "energy is NULL");
    TRY(implicit_ready(s, q, 1));
    TRY(implicit_load_trial(s, trial_velocity, relative));
    double ein, eel;
    TRY(implicit_eval(s, dt, q, mpm_sim::IMP_X, -1, &ein, &eel));
    *energy = ein + eel;
    if (elastic) *elastic = eel;
    return MPM_OK;
}
int mpm_energy_gradient(mpm_t* s, float dt, const MpmImplicitParams* q, const float* trial_velocity, int relative, float* gradient) {
    NEED(s);
    if (!gradient) return fail(MPM_ERR_INVALID, "gradient is NULL");
    TRY(implicit_ready(s, q, 2));
    TRY(implicit_load_trial(s, trial_velocity, relative));
    double ein, eel;
    // nodes outside the active blocks: zero (the vector was zero-initialised and only active blocks are ever written)
    TRY(implicit_eval(s, dt, q, mpm_sim::IMP_X, mpm_sim::IMP_G, &ein, &eel));
    const size_t n = (size_t)s->gd.I * s->gd.J * s->gd.K;
    float* d = nullptr;
    CK(cudaMalloc(&d, n * 3 * sizeof(float)));
    k_imp_export<<<grid_for((int64_t)n, 256), 256, 0, s->stream>>>(s->imp_vec[mpm_sim::IMP_G], s->gd, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(gradient, d, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "energy_gradient: %s", cudaGetErrorString(e));
    return MPM_OK;
}
// mcl::optlib::LBFGS<float, Dynamic, 8>::minimize (LBFGS.hpp:52-152) with Backtracking::search (Backtracking.hpp:38-69) and the
// convergence rule of mathy.hpp:31-35, driven from the host over device vectors; the objective's gradient is analytic. The
// scalars of the two-loop recursion stay on the device (k_lbfgs_coef); the host waits three to four times per iteration: for
// the direction's dot products, for each line-search value, and for the new point's energy together with s.y and y.y.
int mpm_time_integration(mpm_t* s, float dt, const MpmImplicitParams* q, MpmImplicitStats* st) {
    NEED(s);
    TRY(implicit_ready(s, q, mpm_sim::IMP_NVEC));
    enum { M = 8 };
    const int X = mpm_sim::IMP_X, G = mpm_sim::IMP_G, Q = mpm_sim::IMP_Q, XOLD = mpm_sim::IMP_XOLD, GOLD = mpm_sim::IMP_GOLD, TMP = mpm_sim::IMP_TMP,
              S0 = mpm_sim::IMP_S0, Y0 = mpm_sim::IMP_Y0;
    int slot[M];                         // history column i lives in vectors S0 + slot[i], Y0 + slot[i] (rotated instead of copied)
    for (int i = 0; i < M; ++i) slot[i] = i;
    float rho[M] = { 0 };                // 1 / (s_i . y_i): the library recomputes it every iteration from the same vectors
    double* sc = s->imp_acc;
    float* fc = (float*)(s->imp_acc + ACC_FLOATS);
    CK(cudaMemsetAsync(sc, 0, ACC_DOUBLES * sizeof(double), s->stream));
    TRY(implicit_load_trial(s, nullptr, 0));
    double h[4];
    int evals = 0, result = 0;
    TRY(implicit_eval_async(s, dt, q, X, G, ACC_VG)); ++evals;
    TRY(implicit_read(s, ACC_VG, 2, h));
    const double e_start = h[0] + h[1];
    double e_cur = e_start;
    float gamma_k = 1.0f, alpha_init = 1.0f;
    int global_iter = 0, max_iters = q->max_iters;
    for (int k = 0; k < max_iters; ++k) {
        TRY(implicit_lin(s, XOLD, 1.0f, X, 0.0f, -1));
        TRY(implicit_lin(s, GOLD, 1.0f, G, 0.0f, -1));
        TRY(implicit_lin(s, Q, 1.0f, G, 0.0f, -1));
        global_iter++;
        const int iter = k < M ? k : M;
        for (int i = iter - 1; i >= 0; --i) {          // alpha_i = rho_i (s_i . q); q -= alpha_i y_i
            TRY(implicit_dot_async(s, S0 + slot[i], Q, ACC_RUN));
            k_lbfgs_coef<<<1, 1, 0, s->stream>>>(sc + ACC_RUN, fc, i, rho[i], 0);
            CKLAUNCH();
            TRY(implicit_lin(s, Q, 1.0f, Q, 0.0f, Y0 + slot[i], true));
        }
        TRY(implicit_lin(s, Q, gamma_k, Q, 0.0f, -1));
        for (int i = 0; i < iter; ++i) {               // beta = rho_i (y_i . q); q += (alpha_i - beta) s_i
            TRY(implicit_dot_async(s, Q, Y0 + slot[i], ACC_RUN));
            k_lbfgs_coef<<<1, 1, 0, s->stream>>>(sc + ACC_RUN, fc, i, rho[i], 1);
            CKLAUNCH();
            TRY(implicit_lin(s, Q, 1.0f, Q, 0.0f, S0 + slot[i], true));
        }
        s->stats.kernel_launches += 2 * iter;
        // g . q, |g|_inf, q . q, g . g in one read-back
        CK(cudaMemsetAsync(sc + ACC_DOT, 0, 4 * sizeof(double), s->stream));
        TRY(implicit_dot_async(s, G, Q, ACC_DOT, true));
        TRY(implicit_dot_async(s, Q, Q, ACC_DOT + 1));
        TRY(implicit_dot_async(s, G, G, ACC_DOT + 2));
        TRY(implicit_read(s, ACC_DOT, 4, h));
        double gq = h[0], qq = h[1];
        const double gg = h[2];
        if ((float)gq <= 0) {                          // not a descent direction: steepest descent, restart the count (LBFGS.hpp:101-106)
            float ginf; { int bits; memcpy(&bits, &h[3], sizeof bits); memcpy(&ginf, &bits, sizeof bits); }
            TRY(implicit_lin(s, Q, 1.0f, G, 0.0f, -1));
            gq = gg; qq = gg;
            max_iters -= k;
            k = 0;
            alpha_init = (float)std::min(1.0, 1.0 / (double)ginf);
        }
        float rate;
        {   // Backtracking::search(x, p = -q): (the library re-evaluates value and gradient at x here; they are unchanged, so they are reused)
            if ((float)std::sqrt(qq) <= FLT_EPSILON) rate = q->ls_decrease;
            else {
                float a = alpha_init;
                const float fx0 = (float)e_cur;
                const float gtp = -(float)gq;
                int it = 0;
                for (; it < q->ls_max_iters; ++it) {
                    TRY(implicit_lin(s, TMP, 1.0f, X, -a, Q));
                    double ein, eel;
                    TRY(implicit_eval(s, dt, q, TMP, -1, &ein, &eel)); ++evals;
                    if ((float)(ein + eel) <= fx0 + a * q->ls_decrease * gtp) break;
                    a *= q->ls_tau;
                }
                rate = it >= q->ls_max_iters ? -1.0f : a;
            }
        }
        if (rate <= 0) { result = -1; break; }
        TRY(implicit_lin(s, X, 1.0f, X, -rate, Q));
        // Objective::converged(x_last, x, grad): |grad| < tol or |x_last - x| = rate |q| < tol
        if ((float)std::sqrt(gg) < q->tol_grad || (float)(rate * std::sqrt(qq)) < q->tol_step) { result = global_iter; break; }
        TRY(implicit_eval_async(s, dt, q, X, G, ACC_VG)); ++evals;
        int col;
        if (k < M) col = slot[k];
        else { col = slot[0]; for (int i = 0; i + 1 < M; ++i) slot[i] = slot[i + 1]; slot[M - 1] = col; }
        TRY(implicit_lin(s, S0 + col, 1.0f, X, -1.0f, XOLD));
        TRY(implicit_lin(s, Y0 + col, 1.0f, G, -1.0f, GOLD));
        CK(cudaMemsetAsync(sc + ACC_DOT, 0, 2 * sizeof(double), s->stream));
        TRY(implicit_dot_async(s, Y0 + col, Y0 + col, ACC_DOT));
        TRY(implicit_dot_async(s, S0 + col, Y0 + col, ACC_DOT + 1));
        TRY(implicit_read(s, ACC_VG, 4, h));           // energies of the new point, y.y, s.y
        e_cur = h[0] + h[1];
        const double yy = h[2], sy = h[3];
        {   // rho of the column as the next iterations' recursion will see it (columns shift when the window is full)
            const int pos = k < M ? k : M - 1;
            if (k >= M) for (int i = 0; i + 1 < M; ++i) rho[i] = rho[i + 1];
            rho[pos] = (float)(1.0 / sy);
        }
        result = global_iter;
        if (std::fabs((float)yy) <= 0) break;
        gamma_k = (float)sy / (float)yy;
        alpha_init = 1.0f;
    }
    // the minimiser becomes the grid velocity of the used cells (cpp:227-232)
    TRY(implicit_eval_async(s, dt, q, X, G, ACC_VG));
    CK(cudaMemsetAsync(sc + ACC_DOT, 0, sizeof(double), s->stream));
    TRY(implicit_dot_async(s, G, G, ACC_DOT));
    k_imp_to_grid<<<persistent_grid(s, 8), 256, 0, s->stream>>>(s->gblock_list, s->dc, s->grid, s->imp_vec[X]);
    CKLAUNCH(); s->stats.kernel_launches++;
    TRY(implicit_read(s, ACC_VG, 3, h));
    if (st) {
        memset(st, 0, sizeof *st);
        st->iterations = result; st->evaluations = evals;
        st->energy_start = e_start; st->energy_end = h[0] + h[1]; st->grad_norm_end = std::sqrt(h[2]);
    }
    if (result < 0) return fail(MPM_ERR_INVALID, "implicit time integration: line search failed (no step length satisfied the Armijo condition)");
    return MPM_OK;
}

// ---- fused fast path --------------------------------------------------------------------------------------------
#define EV(i) do { if (!s->capturing) CK(cudaEventRecord(s->ev[i], s->stream)); } while (0)   // timing events are not graph nodes
// NVTX ranges name the phases of a substep for nsys / ncu --nvtx (tracing); the ranges cover the host-side
// enqueue, the per-phase device times are in MpmStats::last_ms
struct NvtxRange { explicit NvtxRange(const char* n) { MPM_NVTX_PUSH(n); (void)n; } ~NvtxRange() { MPM_NVTX_POP(); } };
int mpm_substep_begin(mpm_t* s, float dt) {
    NEED(s);
    NvtxRange r("mpm_substep_begin: bin/sort, clear, P2G");
    EV(0);
    TRY(ensure_tau(s));
    { NvtxRange r1("bin/sort"); TRY(do_binning(s)); }
    EV(1);
    TRY(launch_clear(s));
    EV(2);
    { NvtxRange r2("P2G"); TRY((launch_p2g<P2G_FUSED>(s, s->grid, dt))); }
    s->fupd_pending = p2g_fupd(s);
    EV(3);
    return MPM_OK;
}
int mpm_substep_end(mpm_t* s, float dt, const MpmBoxCollider* c, int n) {
    NEED(s);
    NvtxRange r("mpm_substep_end: grid update, F-update, G2P");
    TRY(set_colliders(s, c, n));
    EV(4);
    TRY((launch_grid_update<GU_NORMALIZE | GU_GRAVITY | GU_COLLIDE | GU_COUNT>(s, dt)));
    EV(5);
    if (s->fupd_pending) TRY((launch_g2p<G2P_GATHER | G2P_ADVECT | G2P_REORDER>(s, dt)));      // the F-update already ran inside P2G
    else TRY((launch_g2p<G2P_F | G2P_GATHER | G2P_ADVECT | G2P_REORDER>(s, dt)));
    s->fupd_pending = false;
    EV(6);
    s->tau_valid = true;
    s->stats.substeps_done++;
    return MPM_OK;
}
#undef EV

// graph path: (re)capture two substeps when dt / colliders / particle bound / buffer parity changed
static int graph_prepare(mpm_sim* s, float dt, const MpmBoxCollider* c, int n) {
    const bool same = s->graph_exec && s->graph_dt == dt && s->graph_nc == n && s->graph_n_bound == s->n_bound && s->graph_cur == s->cur &&
                      (n == 0 || memcmp(s->graph_cols.c, c, sizeof(BoxCollider) * n) == 0);
    if (same) return MPM_OK;
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
    TRY(ensure_tau(s));                                     // the one lazily launched kernel stays outside the capture
    s->hist_valid = false;                                  // the captured pair starts with a full binning, whatever ran before
    const int64_t launches0 = s->stats.kernel_launches, steps0 = s->stats.substeps_done;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    s->capturing = true;
    int rc = MPM_OK;
    for (int i = 0; i < 2 && rc == MPM_OK; ++i) {
        rc = mpm_substep_begin(s, dt);
        if (rc == MPM_OK) rc = mpm_substep_end(s, dt, c, n);
    }
    s->capturing = false;
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    s->graph_launches = (int)(s->stats.kernel_launches - launches0);
    s->stats.kernel_launches = launches0; s->stats.substeps_done = steps0;      // nothing has run yet
    if (rc != MPM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&s->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { s->graph_exec = nullptr; return fail(MPM_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e2)); }
    s->graph_dt = dt; s->graph_nc = n; s->graph_n_bound = s->n_bound; s->graph_cur = s->cur;
    if (n) memcpy(s->graph_cols.c, c, sizeof(BoxCollider) * n);
    return MPM_OK;
}
int mpm_substep(mpm_t* s, float dt, const MpmBoxCollider* c, int n, int n_substeps) {
    NEED(s);
    if (n_substeps < 0) return fail(MPM_ERR_INVALID, "n_substeps < 0");
    const bool slab = s->pid_base != 0 || s->gd.lo != 0 || s->gd.hi != s->gd.npbi_global;
    if (s->graph_enabled && !slab && !s->side.stream && n_substeps >= 2 && n >= 0 && n <= MPM_MAX_COLLIDERS && (n == 0 || c)) {
        TRY(graph_prepare(s, dt, c, n));
        const int pairs = n_substeps / 2;
        for (int i = 0; i < pairs; ++i) CK(cudaGraphLaunch(s->graph_exec, s->stream));
        s->stats.substeps_done += 2 * pairs;
        s->stats.kernel_launches += (int64_t)s->graph_launches * pairs;
        s->tau_valid = true; s->binned = false; s->hist_valid = false;
        n_substeps -= 2 * pairs;                            // an odd leftover runs through the plain path below
    }
    for (int i = 0; i < n_substeps; ++i) {
        TRY(mpm_substep_begin(s, dt));
        TRY(mpm_substep_end(s, dt, c, n));
    }
    return MPM_OK;
}

#ifdef MPM_HOST_EMU
// only the host-emulation build of tests/emu has this symbol; capi refuses such a library unless a test asks for it
extern "C" int mpm_emulated_build(void) { return 1; }
#endif

// ---- diagnostics ---------------------------------------------------------------------------------------------------
int mpm_synchronize(mpm_t* s) { NEED(s); CK(cudaStreamSynchronize(s->stream)); return MPM_OK; }

// ---- scene front-end (host only; include/mpm_b200.h) ---------------------------------------------------------------
// glm 0.9.7.1 value arithmetic restated on plain arrays, one rounding per source-level operation (volatile keeps the host
// compiler from contracting or re-associating). m[c][r] is glm's m[c][r] (column c, row r).
namespace scene_fe {
typedef volatile float vf;
struct V4 { float v[4]; };
struct M4 { V4 c[4]; };
static V4 mul(const V4& a, float s) { V4 r; for (int i = 0; i < 4; ++i) { vf t = a.v[i] * s; r.v[i] = t; } return r; }
static V4 mul(const V4& a, const V4& b) { V4 r; for (int i = 0; i < 4; ++i) { vf t = a.v[i] * b.v[i]; r.v[i] = t; } return r; }
static V4 add(const V4& a, const V4& b) { V4 r; for (int i = 0; i < 4; ++i) { vf t = a.v[i] + b.v[i]; r.v[i] = t; } return r; }
static V4 sub(const V4& a, const V4& b) { V4 r; for (int i = 0; i < 4; ++i) { vf t = a.v[i] - b.v[i]; r.v[i] = t; } return r; }
static float mm(float a, float b) { vf t = a * b; return t; }
static float ss(float a, float b) { vf t = a - b; return t; }
static float aa(float a, float b) { vf t = a + b; return t; }
static M4 identity() { M4 m; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m.c[c].v[r] = c == r ? 1.0f : 0.0f; return m; }
// gtc/quaternion.inl:598-622 (mat3_cast) widened to mat4 (mat4_cast)
static M4 to_mat4(const float q[4] /* w x y z */) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float qxx = mm(x, x), qyy = mm(y, y), qzz = mm(z, z), qxz = mm(x, z), qxy = mm(x, y), qyz = mm(y, z), qwx = mm(w, x), qwy = mm(w, y), qwz = mm(w, z);
    M4 m = identity();
    m.c[0].v[0] = ss(1.0f, mm(2.0f, aa(qyy, qzz))); m.c[0].v[1] = mm(2.0f, aa(qxy, qwz)); m.c[0].v[2] = mm(2.0f, ss(qxz, qwy));
    m.c[1].v[0] = mm(2.0f, ss(qxy, qwz)); m.c[1].v[1] = ss(1.0f, mm(2.0f, aa(qxx, qzz))); m.c[1].v[2] = mm(2.0f, aa(qyz, qwx));
    m.c[2].v[0] = mm(2.0f, aa(qxz, qwy)); m.c[2].v[1] = mm(2.0f, ss(qyz, qwx)); m.c[2].v[2] = ss(1.0f, mm(2.0f, aa(qxx, qyy)));
    return m;
}
// gtc/matrix_transform.inl:40-49
static M4 translate(const M4& m, const float t[3]) {
    M4 r = m;
    r.c[3] = add(add(add(mul(m.c[0], t[0]), mul(m.c[1], t[1])), mul(m.c[2], t[2])), m.c[3]);
    return r;
}
// detail/type_mat4x4.inl:704-722
static M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int c = 0; c < 4; ++c)
        r.c[c] = add(add(add(mul(a.c[0], b.c[c].v[0]), mul(a.c[1], b.c[c].v[1])), mul(a.c[2], b.c[c].v[2])), mul(a.c[3], b.c[c].v[3]));
    return r;
}
// detail/type_mat4x4.inl:37-92 (cofactor expansion; the 2x2 sub-determinants are shared between the columns)
static M4 inverse(const M4& M) {
#define E(col_, row_) M.c[col_].v[row_]
    const float c00 = ss(mm(E(2,2), E(3,3)), mm(E(3,2), E(2,3))), c02 = ss(mm(E(1,2), E(3,3)), mm(E(3,2), E(1,3))), c03 = ss(mm(E(1,2), E(2,3)), mm(E(2,2), E(1,3)));
    const float c04 = ss(mm(E(2,1), E(3,3)), mm(E(3,1), E(2,3))), c06 = ss(mm(E(1,1), E(3,3)), mm(E(3,1), E(1,3))), c07 = ss(mm(E(1,1), E(2,3)), mm(E(2,1), E(1,3)));
    const float c08 = ss(mm(E(2,1), E(3,2)), mm(E(3,1), E(2,2))), c10 = ss(mm(E(1,1), E(3,2)), mm(E(3,1), E(1,2))), c11 = ss(mm(E(1,1), E(2,2)), mm(E(2,1), E(1,2)));
    const float c12 = ss(mm(E(2,0), E(3,3)), mm(E(3,0), E(2,3))), c14 = ss(mm(E(1,0), E(3,3)), mm(E(3,0), E(1,3))), c15 = ss(mm(E(1,0), E(2,3)), mm(E(2,0), E(1,3)));
    const float c16 = ss(mm(E(2,0), E(3,2)), mm(E(3,0), E(2,2))), c18 = ss(mm(E(1,0), E(3,2)), mm(E(3,0), E(1,2))), c19 = ss(mm(E(1,0), E(2,2)), mm(E(2,0), E(1,2)));
    const float c20 = ss(mm(E(2,0), E(3,1)), mm(E(3,0), E(2,1))), c22 = ss(mm(E(1,0), E(3,1)), mm(E(3,0), E(1,1))), c23 = ss(mm(E(1,0), E(2,1)), mm(E(2,0), E(1,1)));
    const V4 f0 = { { c00, c00, c02, c03 } }, f1 = { { c04, c04, c06, c07 } }, f2 = { { c08, c08, c10, c11 } };
    const V4 f3 = { { c12, c12, c14, c15 } }, f4 = { { c16, c16, c18, c19 } }, f5 = { { c20, c20, c22, c23 } };
    const V4 v0 = { { E(1,0), E(0,0), E(0,0), E(0,0) } }, v1 = { { E(1,1), E(0,1), E(0,1), E(0,1) } };
    const V4 v2 = { { E(1,2), E(0,2), E(0,2), E(0,2) } }, v3 = { { E(1,3), E(0,3), E(0,3), E(0,3) } };
    const V4 i0 = add(sub(mul(v1, f0), mul(v2, f1)), mul(v3, f2)), i1 = add(sub(mul(v0, f0), mul(v2, f3)), mul(v3, f4));
    const V4 i2 = add(sub(mul(v0, f1), mul(v1, f3)), mul(v3, f5)), i3 = add(sub(mul(v0, f2), mul(v1, f4)), mul(v2, f5));
    const V4 sa = { { 1.0f, -1.0f, 1.0f, -1.0f } }, sb = { { -1.0f, 1.0f, -1.0f, 1.0f } };
    M4 inv;
    inv.c[0] = mul(i0, sa); inv.c[1] = mul(i1, sb); inv.c[2] = mul(i2, sa); inv.c[3] = mul(i3, sb);
    const V4 row0 = { { inv.c[0].v[0], inv.c[1].v[0], inv.c[2].v[0], inv.c[3].v[0] } };
    const V4 d0 = mul(M.c[0], row0);
    const float det = aa(aa(d0.v[0], d0.v[1]), aa(d0.v[2], d0.v[3]));
    vf ood = 1.0f / det;
    for (int c = 0; c < 4; ++c) inv.c[c] = mul(inv.c[c], (float)ood);
#undef E
    return inv;
}
}  // namespace scene_fe

// utils.h:110-112: the arguments are floats, the arithmetic is double (the literal 2000.0), the result a float
static float fe_rand_float(MpmRandFn rnd, void* user, float low, float high) {
    const int r = rnd ? rnd(user) : rand();
    volatile double t = (double)(r % 2000) / 2000.0;
    volatile double d = (double)(volatile float)(high - low);
    volatile double v = (double)low + t * d;
    return (float)v;
}
int mpm_fill_ball(const float origin[3], float radius, float h, MpmRandFn rnd, void* user,
                  float* pos_xyz, int64_t capacity, int64_t* n_written, int64_t* n_missing) {
    if (!origin || !(h > 0.0f) || !(radius >= 0.0f) || capacity < 0 || (capacity > 0 && !pos_xyz)) return fail(MPM_ERR_INVALID, "bad argument");
    using scene_fe::mm; using scene_fe::aa; using scene_fe::ss;
    int centre[3];
    for (int a = 0; a < 3; ++a) { volatile float q = origin[a] / h; centre[a] = (int)q; }      // cpp:20: ivec3(origin / h)
    volatile float reach_f = radius / h;                                                       // cpp:25
    const int reach = (int)reach_f;
    static const float sites[8][3] = { {1, 1, 1}, {1, 1, 3}, {1, 3, 1}, {1, 3, 3}, {3, 1, 1}, {3, 1, 3}, {3, 3, 1}, {3, 3, 3} };
    int64_t stored = 0, missing = 0;
    for (int i = centre[0] - reach; i < centre[0] + reach; ++i)
        for (int j = centre[1] - reach; j < centre[1] + reach; ++j)
            for (int k = centre[2] - reach; k < centre[2] + reach; ++k)
                for (int d = 0; d < 8; ++d) {
                    // utils.h:114-127, generateRandomInsideUnitBall(0.25): phi and u are floats, cos(theta) - 1.0 is a double
                    const float phi = fe_rand_float(rnd, user, 0.0f, (float)(2.0 * 3.1415));
                    volatile double costheta = (double)fe_rand_float(rnd, user, 0.0f, 2.0f) - 1.0;
                    const float u = fe_rand_float(rnd, user, 0.0f, 1.0f);
                    const double theta = acos(costheta);
                    const float r = mm(0.25f, cbrtf(u));
                    volatile double rs = (double)r * sin(theta);
                    volatile double bx = rs * cos((double)phi), by = rs * sin((double)phi), bz = (double)r * cos(theta);   // unqualified cos/sin: the double versions
                    const float ball[3] = { (float)bx, (float)by, (float)bz };
                    const int cell[3] = { i, j, k };
                    float cand[3], dist2 = 0.0f;
                    for (int a = 0; a < 3; ++a) {
                        cand[a] = mm(aa(aa((float)cell[a], mm(sites[d][a], 0.25f)), ball[a]), h);      // cpp:35
                        const float df = ss(cand[a], origin[a]);
                        dist2 = a == 0 ? mm(df, df) : aa(dist2, mm(df, df));
                    }
                    volatile float dist = sqrtf(dist2);
                    if (dist > radius) continue;                                                       // cpp:36
                    if (stored == capacity) { ++missing; continue; }                                   // cpp:38-41
                    for (int a = 0; a < 3; ++a) pos_xyz[3 * stored + a] = cand[a];
                    ++stored;
                    for (int c = 0; c < 3; ++c) { if (rnd) rnd(user); else rand(); }                   // cpp:45-47: r, g, b (then overwritten)
                }
    if (n_written) *n_written = stored;
    if (n_missing) *n_missing = missing;
    return MPM_OK;
}

// ---- bodies from triangle meshes (the reference ships common/objloader.hpp:4 loadOBJ but never calls it) ----
// Minimal Wavefront OBJ reader: "v x y z" and "f a b c ..." (indices may carry /vt/vn, may be negative, polygons are fanned).
int mpm_load_obj(const char* path, float** tri_xyz, int64_t* n_tri) {
    if (!path || !tri_xyz || !n_tri) return fail(MPM_ERR_INVALID, "null argument");
    *tri_xyz = nullptr; *n_tri = 0;
    FILE* f = fopen(path, "r");
    if (!f) return fail(MPM_ERR_INVALID, "cannot open %s", path);
    std::vector<float> v, tri;
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            float x, y, z;
            if (sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3) { v.push_back(x); v.push_back(y); v.push_back(z); }
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            std::vector<long> idx;
            for (char* tok = strtok(line + 2, " \t\r\n"); tok; tok = strtok(nullptr, " \t\r\n")) {
                long i = strtol(tok, nullptr, 10);
                const long nv = (long)(v.size() / 3);
                if (i < 0) i = nv + 1 + i;
                if (i < 1 || i > nv) { fclose(f); return fail(MPM_ERR_INVALID, "%s: face index %ld out of range", path, i); }
                idx.push_back(i - 1);
            }
            for (size_t k = 2; k < idx.size(); ++k)
                for (long i : { idx[0], idx[k - 1], idx[k] }) { tri.push_back(v[3 * i]); tri.push_back(v[3 * i + 1]); tri.push_back(v[3 * i + 2]); }
        }
    }
    fclose(f);
    if (tri.empty()) return fail(MPM_ERR_INVALID, "%s holds no faces", path);
    float* out = (float*)malloc(tri.size() * sizeof(float));
    if (!out) return fail(MPM_ERR_INVALID, "out of memory");
    memcpy(out, tri.data(), tri.size() * sizeof(float));
    *tri_xyz = out; *n_tri = (int64_t)(tri.size() / 9);
    return MPM_OK;
}
void mpm_free(void* p) { free(p); }

// point-in-closed-mesh by ray parity along +x (double precision; the ray is nudged off edges / vertices by an irrational offset)
static bool inside_mesh(const float* tri, int64_t n_tri, const float p[3]) {
    const double py = (double)p[1] + 1.2345678912e-7, pz = (double)p[2] + 2.7182818284e-7, px = p[0];
    int crossings = 0;
    for (int64_t t = 0; t < n_tri; ++t) {
        const float* a = tri + 9 * t; const float* b = a + 3; const float* c = a + 6;
        // barycentric test of (py, pz) in the triangle's yz projection
        const double d = ((double)b[1] - a[1]) * ((double)c[2] - a[2]) - ((double)c[1] - a[1]) * ((double)b[2] - a[2]);
        if (d == 0.0) continue;
        const double u = ((py - a[1]) * ((double)c[2] - a[2]) - ((double)c[1] - a[1]) * (pz - a[2])) / d;
        const double w = (((double)b[1] - a[1]) * (pz - a[2]) - (py - a[1]) * ((double)b[2] - a[2])) / d;
        if (u < 0.0 || w < 0.0 || u + w > 1.0) continue;
        const double x = a[0] + u * ((double)b[0] - a[0]) + w * ((double)c[0] - a[0]);
        if (x > px) ++crossings;
    }
    return crossings & 1;
}
// initializeParticles' fill rule (cpp:29-54: 8 jittered sites per cell, the same random stream as mpm_fill_ball) over the
// bounding box of a closed triangle mesh, a candidate being kept if it lies inside the mesh
int mpm_fill_mesh(const float* tri_xyz, int64_t n_tri, float h, MpmRandFn rnd, void* user,
                  float* pos_xyz, int64_t capacity, int64_t* n_written, int64_t* n_missing) {
    if (!tri_xyz || n_tri < 1 || !(h > 0.0f) || capacity < 0 || (capacity > 0 && !pos_xyz)) return fail(MPM_ERR_INVALID, "bad argument");
    using scene_fe::mm; using scene_fe::aa;
    float lo[3] = { tri_xyz[0], tri_xyz[1], tri_xyz[2] }, hi[3] = { tri_xyz[0], tri_xyz[1], tri_xyz[2] };
    for (int64_t i = 0; i < 3 * n_tri; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], tri_xyz[3 * i + a]); hi[a] = std::max(hi[a], tri_xyz[3 * i + a]); }
    int c0[3], c1[3];
    for (int a = 0; a < 3; ++a) { c0[a] = (int)floorf(lo[a] / h) - 1; c1[a] = (int)floorf(hi[a] / h) + 2; }
    static const float sites[8][3] = { {1, 1, 1}, {1, 1, 3}, {1, 3, 1}, {1, 3, 3}, {3, 1, 1}, {3, 1, 3}, {3, 3, 1}, {3, 3, 3} };
    int64_t stored = 0, missing = 0;
    for (int i = c0[0]; i < c1[0]; ++i)
        for (int j = c0[1]; j < c1[1]; ++j)
            for (int k = c0[2]; k < c1[2]; ++k)
                for (int d = 0; d < 8; ++d) {
                    const float phi = fe_rand_float(rnd, user, 0.0f, (float)(2.0 * 3.1415));          // utils.h:114-127, as in mpm_fill_ball
                    volatile double costheta = (double)fe_rand_float(rnd, user, 0.0f, 2.0f) - 1.0;
                    const float u = fe_rand_float(rnd, user, 0.0f, 1.0f);
                    const double theta = acos(costheta);
                    const float r = mm(0.25f, cbrtf(u));
                    volatile double rs = (double)r * sin(theta);
                    volatile double bx = rs * cos((double)phi), by = rs * sin((double)phi), bz = (double)r * cos(theta);
                    const float ball[3] = { (float)bx, (float)by, (float)bz };
                    const int cell[3] = { i, j, k };
                    float cand[3];
                    for (int a = 0; a < 3; ++a) cand[a] = mm(aa(aa((float)cell[a], mm(sites[d][a], 0.25f)), ball[a]), h);
                    if (!inside_mesh(tri_xyz, n_tri, cand)) continue;
                    if (stored == capacity) { ++missing; continue; }
                    for (int a = 0; a < 3; ++a) pos_xyz[3 * stored + a] = cand[a];
                    ++stored;
                }
    if (n_written) *n_written = stored;
    if (n_missing) *n_missing = missing;
    return MPM_OK;
}
// MeshCollider::sdf is a std::function (hpp:87): a host may install other shapes. A sphere travels through the same POD with
// half_extent = (radius, -1, -1): world_to_local = translate(-centre), sdf = |p_local| - radius; everything after the sdf
// (central-difference normal, friction rule, cpp:264-296) is the reference's bodyCollision unchanged.
int mpm_sphere_collider(const float centre[3], float radius, const float velocity[3], MpmBoxCollider* out) {
    if (!centre || !out || !(radius > 0.0f)) return fail(MPM_ERR_INVALID, "bad argument");
    memset(out, 0, sizeof *out);
    out->world_to_local[0] = out->world_to_local[5] = out->world_to_local[10] = out->world_to_local[15] = 1.0f;
    for (int a = 0; a < 3; ++a) { out->world_to_local[12 + a] = -centre[a]; out->velocity[a] = velocity ? velocity[a] : 0.0f; }
    out->half_extent[0] = radius; out->half_extent[1] = -1.0f; out->half_extent[2] = -1.0f;
    return MPM_OK;
}
int mpm_box_collider_from_transform(const MpmBoxTransform* t, MpmBoxCollider* out) {
    if (!t || !out) return fail(MPM_ERR_INVALID, "null argument");
    using namespace scene_fe;
    const M4 w2l = inverse(mul(translate(identity(), t->translation), to_mat4(t->rotation_wxyz)));    // hpp:82
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) out->world_to_local[c * 4 + r] = w2l.c[c].v[r];
    // hpp:80-81: b = (scale(mat4(), mesh.scale) * vec4(1,1,1,1)).xyz = mesh.scale (products with 0 and 1 are exact)
    for (int a = 0; a < 3; ++a) { out->half_extent[a] = t->scale[a]; out->velocity[a] = t->velocity[a]; }
    return MPM_OK;
}
int mpm_box_transform_move(MpmBoxTransform* t, float time_delta) {
    if (!t) return fail(MPM_ERR_INVALID, "null argument");
    // matrixWorld = translate(mat4(1), timeDelta * velocity) * matrixWorld changes only the last column:
    // x' = ((1*x + 0*y) + 0*z) + (timeDelta*vx)*1; glm::decompose then reads the translation back from that column and
    // re-derives scale / rotation from the untouched 3x3 part
    for (int a = 0; a < 3; ++a) t->translation[a] = scene_fe::aa(t->translation[a], scene_fe::mm(time_delta, t->velocity[a]));
    return MPM_OK;
}
int mpm_box_transform_flip_velocity(MpmBoxTransform* t) {
    if (!t) return fail(MPM_ERR_INVALID, "null argument");
    for (int a = 0; a < 3; ++a) t->velocity[a] = scene_fe::mm(t->velocity[a], -1.0f);     // main.cpp:176-177: velocity *= -1
    return MPM_OK;
}

int mpm_get_stats(mpm_t* s, MpmStats* out) {
    NEED(s);
    if (!out) return fail(MPM_ERR_INVALID, "out is NULL");
    CK(cudaStreamSynchronize(s->stream));
    DevCounters h;
    CK(cudaMemcpy(&h, s->dc, sizeof h, cudaMemcpyDeviceToHost));
    MpmStats& st = s->stats;
    st.n_particles = h.n_slots; st.n_out_of_grid = h.n_out_of_grid; st.n_active_nodes = h.n_active_nodes;
    st.n_particle_blocks = h.n_active_pblocks; st.n_grid_blocks = h.n_active_gblocks; st.svd_failed = h.svd_failed;
    st.reserved[0] = s->sc.pd.fast;      // 1: the pos/h FMA shortcut passed its exhaustive check against __fdiv_rn
    st.reserved[1] = h.mig_overflow;
    st.reserved[2] = h.peer_timeout;     // peer-memory halo: a neighbour's flag never arrived
    if (st.substeps_done > 0) {
        float ms;
        const int pairs[6][2] = { { 0, 1 }, { 1, 2 }, { 2, 3 }, { 4, 5 }, { 5, 6 }, { 3, 4 } };
        for (int k = 0; k < 6; ++k) st.last_ms[k] = cudaEventElapsedTime(&ms, s->ev[pairs[k][0]], s->ev[pairs[k][1]]) == cudaSuccess ? ms : -1.0f;
        st.last_ms[6] = cudaEventElapsedTime(&ms, s->ev[0], s->ev[6]) == cudaSuccess ? ms : -1.0f;
        st.last_ms[7] = (s->side.mid_recorded && cudaEventElapsedTime(&ms, s->ev[5], s->side.mid) == cudaSuccess) ? ms : -1.0f;   // F-update alone
        cudaGetLastError();
    }
    *out = st;
    return MPM_OK;
}

int mpm_download_grid(mpm_t* s, float* grid7) {
    NEED(s);
    if (!grid7) return fail(MPM_ERR_INVALID, "grid7 is NULL");
    const size_t n = (size_t)s->gd.I * s->gd.J * s->gd.K;
    float* d = nullptr;
    CK(cudaMalloc(&d, n * 7 * sizeof(float)));
    k_grid_export<<<grid_for((int64_t)n, 256), 256, 0, s->stream>>>(s->grid, s->gforce, s->gd, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(grid7, d, n * 7 * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "download_grid: %s", cudaGetErrorString(e));
    return MPM_OK;
}
int mpm_upload_grid(mpm_t* s, const float* grid7) {
    NEED(s);
    if (!grid7) return fail(MPM_ERR_INVALID, "grid7 is NULL");
    TRY(ensure_gforce(s));
    const size_t n = (size_t)s->gd.I * s->gd.J * s->gd.K;
    float* d = nullptr;
    CK(cudaMalloc(&d, n * 7 * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d, grid7, n * 7 * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->grid, 0, sizeof(float4) * 64 * (size_t)s->gd.n_gblocks, s->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->gforce, 0, sizeof(float4) * 64 * (size_t)s->gd.n_gblocks, s->stream);
    if (e == cudaSuccess) {
        k_grid_import<<<grid_for((int64_t)n, 256), 256, 0, s->stream>>>(s->grid, s->gforce, s->gd, d);
        k_activate_all<<<grid_for(s->gd.n_gblocks, 256), 256, 0, s->stream>>>(s->gd.n_gblocks, s->gflag, s->gblock_list, s->dc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    s->stats.kernel_launches += 2;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "upload_grid: %s", cudaGetErrorString(e));
    return MPM_OK;
}
int mpm_download_binning(mpm_t* s, int64_t n, int32_t* cells3, int32_t* block_key, int32_t* sorted_ids) {
    NEED(s);
    if (n != s->n_uploaded) return fail(MPM_ERR_INVALID, "n mismatch");
    if (!s->binned) return fail(MPM_ERR_INVALID, "no binning available (call mpm_rasterize_particles_to_grid)");
    int *dc3 = nullptr, *dk = nullptr;
    CK(cudaMalloc(&dc3, sizeof(int) * 3 * (size_t)std::max<int64_t>(n, 1)));
    CK(cudaMalloc(&dk, sizeof(int) * (size_t)std::max<int64_t>(n, 1)));
    k_binning_debug<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->key, s->dc, s->sc.pd, dc3, dk);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && cells3) e = cudaMemcpyAsync(cells3, dc3, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && block_key) e = cudaMemcpyAsync(block_key, dk, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && sorted_ids) e = cudaMemcpyAsync(sorted_ids, s->sorted_ids, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(dc3); cudaFree(dk);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "download_binning: %s", cudaGetErrorString(e));
    return MPM_OK;
}

// ---- slab plumbing ---------------------------------------------------------------------------------------------------
__global__ void k_halo_add(float4* __restrict__ layer, const float4* __restrict__ buf, size_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float4 a = layer[t]; const float4 b = buf[t];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    layer[t] = a;
}
size_t mpm_halo_bytes(const mpm_t* s) { return s ? sizeof(float4) * 64 * (size_t)s->gd.nbj * s->gd.nbk : 0; }
int mpm_halo_pack(mpm_t* s, int upper, void* dev_buf) {
    NEED(s);
    const size_t n = (size_t)64 * s->gd.nbj * s->gd.nbk;
    const float4* src = s->grid + (upper ? (size_t)(s->gd.hi - s->gd.lo) * n : 0);
    CK(cudaMemcpyAsync(dev_buf, src, n * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream));
    return MPM_OK;
}
int mpm_halo_add(mpm_t* s, int upper, const void* dev_buf) {
    NEED(s);
    const size_t n = (size_t)64 * s->gd.nbj * s->gd.nbk;
    float4* dst = s->grid + (upper ? (size_t)(s->gd.hi - s->gd.lo) * n : 0);
    k_halo_add<<<grid_for((int64_t)n, 256), 256, 0, s->stream>>>(dst, (const float4*)dev_buf, n);
    CKLAUNCH(); s->stats.kernel_launches++;
    return MPM_OK;
}
// ---- peer-memory halo: the ghost-layer reduction inside P2G over NVLink-mapped neighbour grids ----------------------------
// Protocol of substep number e (identical on every rank), all on the handle's stream, no host synchronisation:
//   phase 0:  bin, clear (the shared layers are always in the active list)        -> signal "cleared(e)" to both neighbours
//   phase 1:  wait for the neighbours' "cleared(e)"; P2G with remote reds (k_p2g_tile<.., PEER>) -> signal "p2g done(e)"
//   phase 2:  wait for the neighbours' "p2g done(e)"; from here mpm_substep_end runs unchanged (both copies of a shared layer
//             hold the complete sums, both ranks update it redundantly as with the message-based halo)
// A neighbour clears its copy of a shared layer for substep e+1 only after phase 2 of substep e, i.e. after my remote reds of
// substep e are complete, and my reds of substep e+1 wait for its "cleared(e+1)". The flag kernels are one thread each;
// a wait gives up after ~30 s and raises DevCounters::peer_timeout (reported by mpm_sync_counts / mpm_get_stats)
// instead of hanging the device.
__global__ void k_peer_signal(int* flag_a, int* flag_b, int epoch) {
    __threadfence_system();                       // everything this stream did before is visible system-wide first
    if (flag_a) *(volatile int*)flag_a = epoch;
    if (flag_b) *(volatile int*)flag_b = epoch;
    __threadfence_system();
}
__global__ void k_peer_wait(const int* flag_a, const int* flag_b, int epoch, DevCounters* dc) {
    const int* flags[2] = { flag_a, flag_b };
    if (dc->peer_timeout) return;                 // a neighbour is already known to be gone: do not wait for it again and again
    for (int f = 0; f < 2; ++f) {
        if (!flags[f]) continue;
        bool ok = false;
#ifndef MPM_HOST_EMU
        // bounded by time, not by polls: ranks legitimately drift apart by seconds between timed regions (pinned allocations,
        // host-side bookkeeping); ~30 s of SM clock is far below every watchdog above us (NCCL, the job's own timeout)
        const long long t0 = clock64();
        do { ok = *(const volatile int*)flags[f] >= epoch; } while (!ok && clock64() - t0 < 60000000000ll);
#else   // tests/emu: the neighbour is another PROCESS whose emulated kernels take seconds; poll politely for up to 5 minutes
        for (int poll = 0; poll < 300000 && !ok; ++poll) { ok = *(const volatile int*)flags[f] >= epoch; if (!ok) emu_sleep_ms(1); }
#endif
        if (!ok) dc->peer_timeout = 1;
    }
    __threadfence_system();
}
int mpm_peer_export(mpm_t* s, unsigned char* handle) {
    NEED(s);
    if (!handle) return fail(MPM_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MPM_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->grid));
    memcpy(handle, &h, sizeof h);
    return MPM_OK;
}
int mpm_peer_connect_ptr(mpm_t* s, void* lower_grid, int lower_layers, void* upper_grid, int upper_layers) {
    NEED(s);
    const size_t layer_nodes = (size_t)64 * s->gd.nbj * s->gd.nbk;
    if ((lower_grid && lower_layers < 1) || (upper_grid && upper_layers < 1)) return fail(MPM_ERR_INVALID, "a neighbour slab has at least one block layer");
    if ((lower_grid != nullptr) != (s->gd.lo > 0) || (upper_grid != nullptr) != (s->gd.hi < s->gd.npbi_global))
        return fail(MPM_ERR_INVALID, "peer grids must be given exactly for the neighbours this slab has");
    // a neighbour's grid holds its block layers [lo', hi'] (hi'-lo'+1 layers) followed by its flag words
    s->peer.dn = lower_grid ? (float4*)lower_grid + (size_t)lower_layers * layer_nodes : nullptr;           // its ghost layer == my first layer
    s->peer_flags_dn = lower_grid ? (int*)((float4*)lower_grid + (size_t)(lower_layers + 1) * layer_nodes) : nullptr;
    s->peer.up = upper_grid ? (float4*)upper_grid : nullptr;                                                // its first layer == my ghost layer
    s->peer_flags_up = upper_grid ? (int*)((float4*)upper_grid + (size_t)(upper_layers + 1) * layer_nodes) : nullptr;
    s->peer_connected = true;
    s->peer_epoch = 0;
    return MPM_OK;
}
int mpm_peer_connect(mpm_t* s, const unsigned char* lower_handle, int lower_layers, const unsigned char* upper_handle, int upper_layers) {
    NEED(s);
    void *lo_ptr = nullptr, *up_ptr = nullptr;
    cudaIpcMemHandle_t h;
    if (lower_handle) { memcpy(&h, lower_handle, sizeof h); CK(cudaIpcOpenMemHandle(&lo_ptr, h, cudaIpcMemLazyEnablePeerAccess)); s->ipc_dn = lo_ptr; }
    if (upper_handle) { memcpy(&h, upper_handle, sizeof h); CK(cudaIpcOpenMemHandle(&up_ptr, h, cudaIpcMemLazyEnablePeerAccess)); s->ipc_up = up_ptr; }
    return mpm_peer_connect_ptr(s, lo_ptr, lower_layers, up_ptr, upper_layers);
}
int mpm_grid_device_ptr(mpm_t* s, void** grid) {
    NEED(s);
    if (!grid) return fail(MPM_ERR_INVALID, "null argument");
    *grid = s->grid;
    return MPM_OK;
}
#define EVP(i) CK(cudaEventRecord(s->ev[i], s->stream))
int mpm_substep_begin_peer(mpm_t* s, float dt, int phase) {
    NEED(s);
    if (!s->peer_connected) return fail(MPM_ERR_INVALID, "mpm_peer_connect first");
    if (s->prm.p2g_variant == 1) return fail(MPM_ERR_INVALID, "the peer-memory halo needs the tile P2G kernel (p2g_variant != 1)");
    int* mine = (int*)(s->grid + 64 * (size_t)s->gd.n_gblocks);          // my flag words, written by the neighbours
    // my lower neighbour sees me as ITS upper neighbour: I write its flags [1] / [3]; my upper neighbour's [0] / [2]
    if (phase == 0) {
        ++s->peer_epoch;
        EVP(0);
        TRY(ensure_tau(s));
        TRY(do_binning(s));
        EVP(1);
        TRY(launch_clear(s));
        EVP(2);
        k_peer_signal<<<1, 1, 0, s->stream>>>(s->peer_flags_dn ? s->peer_flags_dn + 1 : nullptr, s->peer_flags_up ? s->peer_flags_up + 0 : nullptr, s->peer_epoch);
        CKLAUNCH(); s->stats.kernel_launches++;
    } else if (phase == 1) {
        k_peer_wait<<<1, 1, 0, s->stream>>>(s->peer.dn ? mine + 0 : nullptr, s->peer.up ? mine + 1 : nullptr, s->peer_epoch, s->dc);
        CKLAUNCH(); s->stats.kernel_launches++;
        const Planes nxt = s->planes(s->cur ^ 1);
        CK((launch_p2g_tile<P2G_FUSED>(s->planes(s->cur), s->sorted_ids, s->pblock_list, s->dc, s->grid, s->gd, s->sc, dt,
                                       s->num_sms, (int)s->n_bound, s->stream, p2g_fupd(s) ? &nxt : nullptr, s->prm.fupdate_exact != 1, &s->peer)));
        s->stats.kernel_launches++;
        s->fupd_pending = p2g_fupd(s);
        k_peer_signal<<<1, 1, 0, s->stream>>>(s->peer_flags_dn ? s->peer_flags_dn + 3 : nullptr, s->peer_flags_up ? s->peer_flags_up + 2 : nullptr, s->peer_epoch);
        CKLAUNCH(); s->stats.kernel_launches++;
    } else if (phase == 2) {
        k_peer_wait<<<1, 1, 0, s->stream>>>(s->peer.dn ? mine + 2 : nullptr, s->peer.up ? mine + 3 : nullptr, s->peer_epoch, s->dc);
        CKLAUNCH(); s->stats.kernel_launches++;
        EVP(3);
    } else return fail(MPM_ERR_INVALID, "phase must be 0, 1 or 2");
    return MPM_OK;
}
#undef EVP

static int ensure_out_buffers(mpm_sim* s);
int mpm_migrate_outgoing(mpm_t* s, int64_t* n_down, int64_t* n_up, const void** dev_down, const void** dev_up) {
    NEED(s);
    if (!n_down || !n_up || !dev_down || !dev_up) return fail(MPM_ERR_INVALID, "null argument");
    TRY(ensure_out_buffers(s));
    if (s->mig_packed) {
        // the last gather has packed the leavers already (header + records): hand those out, counts from the headers
        int hdr[2] = { 0, 0 };
        DevCounters h;
        for (int d = 0; d < 2; ++d) CK(cudaMemcpyAsync(&hdr[d], s->out_buf[d], sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaMemcpyAsync(&h, s->dc, sizeof h, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        *n_down = std::min<int64_t>(hdr[0], s->out_cap); *n_up = std::min<int64_t>(hdr[1], s->out_cap);
        *dev_down = s->out_buf[0] + 1; *dev_up = s->out_buf[1] + 1;
        s->n_bound = h.n_slots;
        s->mig_packed = false; s->binned = false;
        return MPM_OK;
    }
    CK(cudaMemsetAsync(s->dc->n_mig, 0, 3 * sizeof(int), s->stream));
    k_mark_outgoing<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->dc, s->gd, s->sc.pd, s->out_buf[0], s->out_buf[1], (int)s->out_cap);
    CKLAUNCH(); s->stats.kernel_launches++;
    DevCounters h;
    CK(cudaMemcpyAsync(&h, s->dc, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *n_down = h.n_mig[0]; *n_up = h.n_mig[1];
    *dev_down = s->out_buf[0]; *dev_up = s->out_buf[1];
    s->n_bound = h.n_slots;          // exact after the sync
    s->binned = false; s->hist_valid = false; s->mig_packed = false;
    return MPM_OK;
}
int mpm_migrate_append(mpm_t* s, const void* dev_buf, int64_t n) {
    NEED(s);
    if (n < 0 || (n > 0 && !dev_buf)) return fail(MPM_ERR_INVALID, "bad argument");
    if (n == 0) return MPM_OK;
    if (s->n_bound + n > s->capacity) return fail(MPM_ERR_CAPACITY, "migration overflows the slab capacity (%lld + %lld > %lld)", (long long)s->n_bound, (long long)n, (long long)s->capacity);
    k_append_incoming<<<grid_for(n, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->dc, (const float4*)dev_buf, (int)s->n_bound, (int)n);
    CKLAUNCH(); s->stats.kernel_launches++;
    s->n_bound += n;
    s->binned = false; s->hist_valid = false;
    return MPM_OK;
}
static int64_t default_migrate_capacity(const mpm_sim* s) { return std::min<int64_t>(std::max<int64_t>(1 << 14, s->capacity / 128), 1 << 18); }
int mpm_set_migrate_capacity(mpm_t* s, int64_t records) {
    NEED(s);
    if (records < 1 || records > ((int64_t)1 << 24)) return fail(MPM_ERR_INVALID, "migrate capacity out of range");
    if (s->out_buf[0]) return fail(MPM_ERR_INVALID, "migration buffers already allocated");
    s->out_cap = records;
    return MPM_OK;
}
static int ensure_out_buffers(mpm_sim* s) {
    if (s->out_buf[0]) return MPM_OK;
    if (s->out_cap <= 0) s->out_cap = default_migrate_capacity(s);
    for (int d = 0; d < 2; ++d) CK(cudaMalloc(&s->out_buf[d], sizeof(float4) * (1 + NPLANES * (size_t)s->out_cap)));
    return MPM_OK;
}
size_t mpm_migrate_buffer_bytes(const mpm_t* s) {
    if (!s) return 0;
    const int64_t cap = s->out_cap > 0 ? s->out_cap : default_migrate_capacity(s);
    return sizeof(float4) * (1 + NPLANES * (size_t)cap);
}
int mpm_migrate_pack(mpm_t* s, const void** dev_down, const void** dev_up) {
    NEED(s);
    if (!dev_down || !dev_up) return fail(MPM_ERR_INVALID, "null argument");
    TRY(ensure_out_buffers(s));
    *dev_down = s->out_buf[0]; *dev_up = s->out_buf[1];
    if (s->mig_packed) {                // the gather has packed the leavers already (and kept keys + histogram consistent)
        s->mig_packed = false;
        s->binned = false;
        return MPM_OK;
    }
    for (int d = 0; d < 2; ++d) CK(cudaMemsetAsync(s->out_buf[d], 0, sizeof(float4), s->stream));
    k_mark_outgoing_hdr<<<grid_for(s->n_bound, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->dc, s->gd, s->sc.pd, s->out_buf[0], s->out_buf[1], (int)s->out_cap);
    CKLAUNCH(); s->stats.kernel_launches++;
    s->binned = false; s->hist_valid = false;
    return MPM_OK;
}
int mpm_migrate_append_packed(mpm_t* s, const void* dev_buf) {
    NEED(s);
    if (!dev_buf) return fail(MPM_ERR_INVALID, "null buffer");
    TRY(ensure_out_buffers(s));
    const int cap = (int)s->out_cap;
    k_append_incoming_hdr<<<grid_for(cap, 256), 256, 0, s->stream>>>(s->planes(s->cur), s->dc, (const float4*)dev_buf, cap, (int)s->capacity,
                                                                   s->gd, s->sc.pd, s->hist_valid ? s->key : nullptr, s->hist_valid ? s->blk_count : nullptr);
    CKLAUNCH();
    k_bump_slots<<<1, 1, 0, s->stream>>>(s->dc, (const float4*)dev_buf, cap, (int)s->capacity);
    CKLAUNCH(); s->stats.kernel_launches += 2;
    s->n_bound = std::min<int64_t>(s->capacity, s->n_bound + cap);      // upper bound; mpm_sync_counts tightens it
    s->binned = false;          // (keys + histogram stay valid: the appended particles have added theirs)
    return MPM_OK;
}
// peer-memory migration (same flag words as the peer-memory halo, [4..7]): a rank packs its leavers into its own
// two buffers as before; the neighbours READ them through their IPC mappings (pull), so no message is sent.
//   phase 0: wait until both neighbours have consumed my buffers of the previous substep; pack; signal "packed(e)"
//   phase 1: wait for the neighbours' "packed(e)"; append from the lower neighbour's UP and the upper neighbour's DOWN buffer;
//            signal "consumed(e)"
int mpm_peer_export_migration(mpm_t* s, unsigned char* handle_down, unsigned char* handle_up) {
    NEED(s);
    if (!handle_down || !handle_up) return fail(MPM_ERR_INVALID, "null argument");
    TRY(ensure_out_buffers(s));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->out_buf[0])); memcpy(handle_down, &h, sizeof h);
    CK(cudaIpcGetMemHandle(&h, s->out_buf[1])); memcpy(handle_up, &h, sizeof h);
    return MPM_OK;
}
int mpm_peer_connect_migration_ptr(mpm_t* s, const void* lower_up_buf, const void* upper_down_buf) {
    NEED(s);
    if (!s->peer_connected) return fail(MPM_ERR_INVALID, "mpm_peer_connect first (the flag words live behind the neighbours' grids)");
    if ((lower_up_buf != nullptr) != (s->gd.lo > 0) || (upper_down_buf != nullptr) != (s->gd.hi < s->gd.npbi_global))
        return fail(MPM_ERR_INVALID, "neighbour buffers must be given exactly for the neighbours this slab has");
    TRY(ensure_out_buffers(s));
    s->peer_in_dn = (const float4*)lower_up_buf; s->peer_in_up = (const float4*)upper_down_buf;
    s->peer_mig_connected = true;
    s->peer_mig_epoch = 0;
    return MPM_OK;
}
int mpm_peer_connect_migration(mpm_t* s, const unsigned char* lower_up_handle, const unsigned char* upper_down_handle) {
    NEED(s);
    void *lo_ptr = nullptr, *up_ptr = nullptr;
    cudaIpcMemHandle_t h;
    if (lower_up_handle) { memcpy(&h, lower_up_handle, sizeof h); CK(cudaIpcOpenMemHandle(&lo_ptr, h, cudaIpcMemLazyEnablePeerAccess)); s->ipc_mig_dn = lo_ptr; }
    if (upper_down_handle) { memcpy(&h, upper_down_handle, sizeof h); CK(cudaIpcOpenMemHandle(&up_ptr, h, cudaIpcMemLazyEnablePeerAccess)); s->ipc_mig_up = up_ptr; }
    return mpm_peer_connect_migration_ptr(s, lo_ptr, up_ptr);
}
int mpm_migrate_peer(mpm_t* s, int phase) {
    NEED(s);
    if (!s->peer_mig_connected) return fail(MPM_ERR_INVALID, "mpm_peer_connect_migration first");
    int* mine = (int*)(s->grid + 64 * (size_t)s->gd.n_gblocks);
    const bool has_dn = s->gd.lo > 0, has_up = s->gd.hi < s->gd.npbi_global;
    if (phase == 0) {
        ++s->peer_mig_epoch;
        if (!s->mig_packed) {           // (a gather that packed the leavers has waited for "consumed" itself, before refilling the buffers)
            k_peer_wait<<<1, 1, 0, s->stream>>>(has_dn ? mine + 6 : nullptr, has_up ? mine + 7 : nullptr, s->peer_mig_epoch - 1, s->dc);
            CKLAUNCH(); s->stats.kernel_launches++;
        }
        const void *d0, *d1;
        TRY(mpm_migrate_pack(s, &d0, &d1));
        k_peer_signal<<<1, 1, 0, s->stream>>>(has_dn ? s->peer_flags_dn + 5 : nullptr, has_up ? s->peer_flags_up + 4 : nullptr, s->peer_mig_epoch);
        CKLAUNCH(); s->stats.kernel_launches++;
    } else if (phase == 1) {
        k_peer_wait<<<1, 1, 0, s->stream>>>(has_dn ? mine + 4 : nullptr, has_up ? mine + 5 : nullptr, s->peer_mig_epoch, s->dc);
        CKLAUNCH(); s->stats.kernel_launches++;
        if (has_dn) TRY(mpm_migrate_append_packed(s, s->peer_in_dn));
        if (has_up) TRY(mpm_migrate_append_packed(s, s->peer_in_up));
        k_peer_signal<<<1, 1, 0, s->stream>>>(has_dn ? s->peer_flags_dn + 7 : nullptr, has_up ? s->peer_flags_up + 6 : nullptr, s->peer_mig_epoch);
        CKLAUNCH(); s->stats.kernel_launches++;
    } else return fail(MPM_ERR_INVALID, "phase must be 0 or 1");
    return MPM_OK;
}
int mpm_sync_counts(mpm_t* s) {
    NEED(s);
    DevCounters h;
    CK(cudaMemcpyAsync(&h, s->dc, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->n_bound = h.n_slots;
    if (h.peer_timeout) return fail(MPM_ERR_CUDA, "peer-memory halo: a neighbour's flag never arrived (the grids of this substep are incomplete)");
    if (h.mig_overflow) {
        const int what = h.mig_overflow;
        CK(cudaMemsetAsync(&s->dc->mig_overflow, 0, sizeof(int), s->stream));
        return fail(MPM_ERR_CAPACITY, what == 2 ? "slab capacity exceeded by incoming particles" : "more particles left the slab in one substep than the migration buffer holds");
    }
    return MPM_OK;
}
int mpm_set_pid_base(mpm_t* s, int64_t pid_base) {
    NEED(s);
    if (pid_base < 0 || pid_base >= ((int64_t)1 << 31)) return fail(MPM_ERR_INVALID, "pid_base out of range");
    s->pid_base = pid_base;
    return MPM_OK;
}
int mpm_debug_p2g_profile(mpm_t* s, int64_t* ticks8, int reset) {      /* profiling builds (-DMPM_P2G_PROFILE): per-phase clock64 sums */
    NEED(s);
    CK(cudaStreamSynchronize(s->stream));
    if (ticks8) CK(cudaMemcpy(ticks8, s->dc->prof, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(s->dc->prof, 0, sizeof(long long) * 8));
    return MPM_OK;
}
int mpm_reduce_invariants(mpm_t* s, double* fsum5, uint64_t* isum3) {
    NEED(s);
    if (!fsum5 || !isum3) return fail(MPM_ERR_INVALID, "bad argument");
    unsigned char* d = nullptr;
    CK(cudaMalloc(&d, 64));
    cudaError_t e = cudaMemsetAsync(d, 0, 64, s->stream);
    if (e == cudaSuccess) {
        k_invariants<<<s->num_sms * 4, 256, 0, s->stream>>>(s->planes(s->cur), s->dc, reinterpret_cast<double*>(d), reinterpret_cast<unsigned long long*>(d + 40));
        e = cudaGetLastError();
    }
    unsigned char h[64];
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "reduce_invariants: %s", cudaGetErrorString(e));
    memcpy(fsum5, h, 40); memcpy(isum3, h + 40, 24);
    return MPM_OK;
}
int mpm_download_live_particles(mpm_t* s, int64_t capacity, int64_t* n_out, float* state35, int32_t* pid) {
    NEED(s);
    if (!n_out || !state35 || !pid || capacity < 0) return fail(MPM_ERR_INVALID, "bad argument");
    float* d35 = nullptr; int *dpid = nullptr, *dcnt = nullptr;
    const size_t cap = (size_t)std::max<int64_t>(capacity, 1);
    CK(cudaMalloc(&d35, cap * 35 * sizeof(float)));
    CK(cudaMalloc(&dpid, cap * sizeof(int)));
    CK(cudaMalloc(&dcnt, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(dcnt, 0, sizeof(int), s->stream);
    int cnt = 0;
    if (e == cudaSuccess) {
        k_export_live<<<grid_for(s->n_bound, 128), 128, 0, s->stream>>>(s->planes(s->cur), s->dc, d35, dpid, dcnt, (int)capacity);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, dcnt, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess && cnt <= capacity && cnt > 0) {
        e = cudaMemcpy(state35, d35, (size_t)cnt * 35 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(pid, dpid, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(d35); cudaFree(dpid); cudaFree(dcnt);
    s->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(MPM_ERR_CUDA, "download_live_particles: %s", cudaGetErrorString(e));
    *n_out = cnt;
    if (cnt > capacity) return fail(MPM_ERR_CAPACITY, "%d live particles exceed the caller's capacity %lld", cnt, (long long)capacity);
    return MPM_OK;
}

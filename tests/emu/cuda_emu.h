// TEST INFRASTRUCTURE ONLY -- never part of the product (libmpm_b200.so is built by nvcc from the same kernel sources
// and has no CPU path). This header lets g++ compile the CUDA kernel sources under csrc/ for the HOST so that their
// index arithmetic, barrier structure and copy/transaction bookkeeping can be executed without a GPU:
//   * every CUDA thread is a fiber (own stack, hand-written context switch) of the launching OS thread; the CTAs of a
//     launch run one after the other; fibers are resumed in a seeded pseudo-random order (EMU_SEED) at every barrier /
//     wait, so that a missing barrier shows up as a poisoned or stale read for some seed;
//   * __syncthreads / __syncwarp / __shfl_* are real barriers (every live thread of the CTA / warp must arrive); if no
//     fiber can make progress the launch dies with "deadlock" (e.g. an mbarrier whose byte count never completes);
//   * cp.async.bulk + mbarrier are emulated with exact transaction-byte accounting and 16-byte alignment checks;
//   * the _rn intrinsics map to plain IEEE operations (compile with -ffp-contract=off).
// What it cannot show: performance, the hardware memory model, real concurrency between CTAs.
// Used by tests/emu/emu_harness.cpp (differential tests: experimental kernel variants against the hardware-validated ones).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <tuple>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define MPM_HOST_EMU 1
#undef __shared__
#define __shared__ static          /* CTAs run one at a time: a static is shared by the CTA's threads */
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

namespace emu {

[[noreturn]] inline void die(const char* what) {
    std::fprintf(stderr, "cuda_emu: %s\n", what);
    std::fflush(stderr);
    std::abort();
}

void yield();                       // hand the OS thread to another fiber of the running CTA
void note_progress();
// Resume the fibers in thread order instead of a random order, and let the thread that completes a barrier wait for its
// turn too: same-address shared atomics of a warp are then served in lane order, as on the hardware (used by the
// bank-model profile; the differential tests keep the random order, which is what finds missing barriers).
inline thread_local bool lane_order = false;
struct Barrier {                    // n = live participants; a participant that exits stops counting (leave())
    int n = 0, arrived = 0;
    unsigned gen = 0;
    void wait() {
        const unsigned g = gen;
        if (++arrived >= n) { arrived = 0; ++gen; note_progress(); if (lane_order) yield(); return; }
        while (gen == g) yield();
    }
    void leave() {
        --n;
        if (n > 0 && arrived >= n) { arrived = 0; ++gen; note_progress(); }
    }
};

struct Cta {
    int nthreads = 0;
    Barrier cta_bar;
    std::vector<Barrier> warp_bar;
    std::vector<unsigned long long> slot;       // shuffle exchange, one per thread
    std::vector<unsigned char> smem;            // dynamic shared memory
};

inline thread_local Cta* cta = nullptr;
inline thread_local int tid = 0;
struct Idx { unsigned x = 0, y = 0, z = 0; };

inline std::mutex& mbar_mutex() { static std::mutex m; return m; }

// mbarrier state packed into the 64-bit shared-memory word: [0,32) pending tx bytes (signed), [32,48) pending arrivals,
// [48,63) arrival count, bit 63 phase parity
struct Mbar {
    static long long tx(unsigned long long w) { return (int)(w & 0xffffffffull); }
    static int pending(unsigned long long w) { return (int)((w >> 32) & 0xffff); }
    static int count(unsigned long long w) { return (int)((w >> 48) & 0x7fff); }
    static int phase(unsigned long long w) { return (int)(w >> 63); }
    static unsigned long long pack(long long tx, int pending, int count, int phase) {
        return ((unsigned long long)(unsigned)(int)tx) | ((unsigned long long)(pending & 0xffff) << 32) |
               ((unsigned long long)(count & 0x7fff) << 48) | ((unsigned long long)(phase & 1) << 63);
    }
    static void update(unsigned long long* bar, long long dtx, int darrive) {
        std::lock_guard<std::mutex> lk(mbar_mutex());
        unsigned long long w = __atomic_load_n(bar, __ATOMIC_ACQUIRE);
        long long t = tx(w) + dtx;
        int p = pending(w) - darrive, c = count(w), ph = phase(w);
        if (p < 0) die("mbarrier: more arrivals than the barrier was initialised for");
        if (p == 0 && t == 0) { ph ^= 1; p = c; }
        __atomic_store_n(bar, pack(t, p, c, ph), __ATOMIC_RELEASE);
    }
};
inline void mbar_init(unsigned long long* bar, int count) { __atomic_store_n(bar, Mbar::pack(0, count, count, 0), __ATOMIC_RELEASE); }
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { Mbar::update(bar, bytes, 1); }
inline long long& tma_bytes_total() { static long long b = 0; return b; }
inline void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15) || bytes == 0) die("cp.async.bulk: address / size not a multiple of 16 bytes");
    Cta* c = cta;
    const unsigned char* lo = c->smem.data();
    if ((const unsigned char*)dst < lo || (const unsigned char*)dst + bytes > lo + c->smem.size()) die("cp.async.bulk: destination outside the CTA's shared memory");
    std::memcpy(dst, src, bytes);
    { std::lock_guard<std::mutex> lk(mbar_mutex()); tma_bytes_total() += bytes; }
    Mbar::update(bar, -(long long)bytes, 0);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (Mbar::phase(__atomic_load_n(bar, __ATOMIC_ACQUIRE)) == (int)(parity & 1)) yield();    // a wait that can never complete ends in "deadlock"
    note_progress();
}

// ---- shared-memory wavefront model -------------------------------------------------------------------------------
// The kernels mark shared-memory instructions with MPM_SMEM_PROBE(site, key, ptr, bytes) (a no-op in the nvcc build). With
// EMU_SMEM_PROFILE set (or SmemProbe::on), the emulator groups the lanes of a warp that execute the same (site, key) in the
// same epoch -- the fibers run one after the other, so the grouping is by key, not by time -- and counts the wavefronts that
// request costs under this bank model: 32 banks of 4 bytes; a request is served in phases of 32 lanes (<= 4-byte accesses)
// or 16 lanes (8- and 16-byte accesses); a phase costs max over banks of the number of DISTINCT words it needs from that
// bank (same word = broadcast). Calibration: ncu counts 2.01 wavefronts per LDS.128 and no bank conflicts for the gather
// kernel at 64 Mi particles, 8 per cell (lanes of a half-warp read 2 different 16-byte nodes), which this model reproduces;
// a strict quarter-warp model would give 4.
struct SmemProbe {
    struct Req { int lane; long long off; int bytes; };
    struct Site { long long requests = 0, wavefronts = 0, lanes = 0; };
    bool on = std::getenv("EMU_SMEM_PROFILE") != nullptr;
    std::map<std::tuple<int, int, long long, int>, std::vector<Req>> open;     // (warp, site, key, epoch) of the running CTA
    std::map<int, Site> sites;
    std::vector<int> epoch;                                                     // per thread of the running CTA
    static int wavefronts_of(const std::vector<Req>& rq) {
        if (rq.empty()) return 0;
        const int per_phase = rq[0].bytes > 4 ? 16 : 32;
        int total = 0;
        for (int l0 = 0; l0 < 32; l0 += per_phase) {
            std::vector<long long> words;
            for (const Req& r : rq)
                if (r.lane >= l0 && r.lane < l0 + per_phase)
                    for (long long w = r.off / 4; w <= (r.off + r.bytes - 1) / 4; ++w) words.push_back(w);
            std::sort(words.begin(), words.end());
            words.erase(std::unique(words.begin(), words.end()), words.end());
            int per_bank[32] = { 0 }, mx = 0;
            for (long long w : words) mx = std::max(mx, ++per_bank[(int)(w & 31)]);
            total += mx;
        }
        return total;
    }
    void flush() {
        for (auto& kv : open) {
            Site& st = sites[std::get<1>(kv.first)];
            st.requests++; st.lanes += (long long)kv.second.size(); st.wavefronts += wavefronts_of(kv.second);
        }
        open.clear();
    }
    void reset() { open.clear(); sites.clear(); }
};
inline SmemProbe& smem_probe_state() { static thread_local SmemProbe s; return s; }
void smem_probe(int site, long long key, const void* p, int bytes);
void smem_epoch();

// run `body` as a grid of CTAs, one CTA at a time, `threads` OS threads per CTA
inline void launch(unsigned grid, unsigned threads, size_t smem_bytes, const std::function<void()>& body, const char* name = "");

}  // namespace emu

inline thread_local emu::Idx threadIdx, blockIdx, blockDim, gridDim;

// ---- fibers ------------------------------------------------------------------------------------------------------------
#if !defined(__x86_64__)
#error "tests/emu: the fiber context switch is written for x86-64"
#endif
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {
struct Sched {
    static constexpr size_t STACK = 256 * 1024;
    struct Fiber { void* sp = nullptr; char* stack = nullptr; bool done = true; };
    std::vector<Fiber> fib;
    void* main_sp = nullptr;
    int cur = -1, live = 0;
    bool progress = false;
    Cta* c = nullptr;
    const std::function<void()>* body = nullptr;
    Idx block, bdim, gdim;
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    Sched() { if (const char* e = std::getenv("EMU_SEED")) rng ^= (unsigned long long)std::atoll(e) * 0xD1B54A32D192ED03ull; }
    ~Sched() { for (Fiber& f : fib) std::free(f.stack); }
    static void entry();
    void prepare(int i) {
        Fiber& f = fib[i];
        if (!f.stack) f.stack = (char*)std::aligned_alloc(64, STACK);
        uintptr_t top = ((uintptr_t)(f.stack + STACK)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                    // return address of entry(): never used
        *--sp = (void*)&Sched::entry;       // popped by emu_switch's ret
        for (int r = 0; r < 6; ++r) *--sp = nullptr;
        f.sp = sp; f.done = false;
    }
    void resume(int i) {
        cur = i;
        cta = c; tid = i;
        threadIdx = Idx{ (unsigned)i, 0, 0 }; blockIdx = block; blockDim = bdim; gridDim = gdim;
        emu_switch(&main_sp, fib[i].sp);
    }
    void run_cta(Cta* cta_, int threads, Idx b, Idx bd, Idx gd, const std::function<void()>& f, const char* name) {
        if (cur >= 0) die("nested launch");
        if ((int)fib.size() < threads) fib.resize(threads);
        c = cta_; body = &f; block = b; bdim = bd; gdim = gd;
        for (int i = 0; i < threads; ++i) prepare(i);
        live = threads;
        std::vector<int> order(threads);
        for (int i = 0; i < threads; ++i) order[i] = i;
        int idle_rounds = 0;
        while (live > 0) {
            progress = false;
            // a fresh pseudo-random resume order every round (xorshift, Fisher-Yates)
            if (lane_order) std::sort(order.begin(), order.end());
            else
            for (int i = threads - 1; i > 0; --i) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                std::swap(order[i], order[(int)(rng % (unsigned long long)(i + 1))]);
            }
            for (int k = 0; k < threads; ++k) if (!fib[order[k]].done) resume(order[k]);
            cur = -1;
            idle_rounds = progress ? 0 : idle_rounds + 1;
            if (idle_rounds > 3) {
                std::fprintf(stderr, "cuda_emu: deadlock in %s, block %u: %d threads blocked, none can make progress\n", name, b.x, live);
                die("deadlock");
            }
        }
    }
};
inline Sched& sched() { static thread_local Sched s; return s; }
inline void Sched::entry() {
    Sched& s = sched();
    (*s.body)();
    const int i = s.cur;
    s.fib[i].done = true;
    --s.live;
    s.progress = true;
    s.c->cta_bar.leave();
    s.c->warp_bar[i >> 5].leave();
    void* dummy;
    emu_switch(&dummy, s.main_sp);
    die("resumed a finished fiber");
}
inline void yield() { Sched& s = sched(); if (s.cur < 0) die("yield outside a kernel"); emu_switch(&s.fib[s.cur].sp, s.main_sp); }
inline void note_progress() { sched().progress = true; }
}  // namespace emu

#include <chrono>
#include <map>
#include <string>
namespace emu {
struct Profile {       // EMU_PROFILE=1: seconds per kernel name, printed at exit
    std::map<std::string, std::pair<double, long>> t;
    bool on = std::getenv("EMU_PROFILE") != nullptr;
    ~Profile() { if (on) for (auto& kv : t) std::fprintf(stderr, "emu profile %-28s %8.3f s %6ld launches\n", kv.first.c_str(), kv.second.first, kv.second.second); }
};
inline Profile& profile() { static Profile p; return p; }
}
inline void emu::launch(unsigned grid, unsigned threads, size_t smem_bytes, const std::function<void()>& body, const char* name) {
    if (threads == 0 || threads > 1024) die("launch: bad block size");
    struct Timer {
        const char* n; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~Timer() { if (profile().on) { auto& e = profile().t[n]; e.first += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); e.second++; } }
    } timer{ name };
    for (unsigned b = 0; b < grid; ++b) {
        Cta c;
        c.nthreads = (int)threads;
        c.cta_bar.n = (int)threads;
        c.warp_bar = std::vector<Barrier>((threads + 31) / 32);
        for (unsigned w = 0; w < c.warp_bar.size(); ++w) c.warp_bar[w].n = (int)std::min(32u, threads - 32 * w);
        c.slot.assign(threads, 0);
        c.smem.assign(smem_bytes + 256, 0xcd);          // poison: uninitialised reads show up as garbage
        sched().run_cta(&c, (int)threads, Idx{ b, 0, 0 }, Idx{ threads, 1, 1 }, Idx{ grid, 1, 1 }, body, name);
        if (smem_probe_state().on) { smem_probe_state().flush(); smem_probe_state().epoch.clear(); }
    }
}

template <class F> inline cudaError_t emu_func_set_attribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// dynamic shared memory of the running CTA, 128-byte aligned
inline unsigned char* emu_dyn_smem() {
    unsigned char* p = emu::cta->smem.data();
    return p + ((128 - ((uintptr_t)p & 127)) & 127);
}

inline void emu::smem_probe(int site, long long key, const void* p, int bytes) {
    SmemProbe& s = smem_probe_state();
    if (!s.on) return;
    const long long off = (const unsigned char*)p - emu_dyn_smem();
    if (off < 0 || off + bytes > (long long)cta->smem.size()) die("smem probe: address outside the CTA's dynamic shared memory");
    if ((int)s.epoch.size() < cta->nthreads) s.epoch.assign(cta->nthreads, 0);
    s.open[std::make_tuple(tid >> 5, site, key, s.epoch[tid])].push_back(SmemProbe::Req{ tid & 31, off, bytes });
}
inline void emu::smem_epoch() {
    SmemProbe& s = smem_probe_state();
    if (!s.on) return;
    if ((int)s.epoch.size() < cta->nthreads) s.epoch.assign(cta->nthreads, 0);
    ++s.epoch[tid];
}

// ---- synchronisation and warp collectives -------------------------------------------------------------------------
inline void __syncthreads() { emu::cta->cta_bar.wait(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    if (mask != 0xffffffffu) emu::die("__syncwarp with a partial mask is not emulated");
    emu::cta->warp_bar[emu::tid >> 5].wait();
}
template <class T> inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
    emu::Cta* c = emu::cta;
    const int w = emu::tid >> 5;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    c->slot[emu::tid] = bits;
    c->warp_bar[w].wait();
    const unsigned long long got = c->slot[w * 32 + (src_lane & 31)];
    c->warp_bar[w].wait();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    if (mask != 0xffffffffu || width != 32) emu::die("__shfl_sync: only full-warp shuffles are emulated");
    return emu_exchange(v, src);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu || width != 32) emu::die("__shfl_up_sync: only full-warp shuffles are emulated");
    const int lane = emu::tid & 31;
    const T got = emu_exchange(v, lane >= (int)delta ? lane - (int)delta : lane);
    return lane >= (int)delta ? got : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu || width != 32) emu::die("__shfl_down_sync: only full-warp shuffles are emulated");
    const int lane = emu::tid & 31;
    const T got = emu_exchange(v, lane + (int)delta < 32 ? lane + (int)delta : lane);
    return lane + (int)delta < 32 ? got : v;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    if (mask != 0xffffffffu || width != 32) emu::die("__shfl_xor_sync: only full-warp shuffles are emulated");
    return emu_exchange(v, (emu::tid & 31) ^ x);
}
// full-warp collectives (every lane of the warp must call them, as the kernels do); partial masks are not emulated
inline unsigned __activemask() { emu::die("__activemask is not emulated"); }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    if (mask != 0xffffffffu) emu::die("__match_any_sync: only the full mask is emulated");
    emu::Cta* c = emu::cta;
    const int w = emu::tid >> 5, nl = c->warp_bar[w].n;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    c->slot[emu::tid] = bits;
    c->warp_bar[w].wait();
    unsigned peers = 0;
    for (int l = 0; l < nl; ++l) if (c->slot[w * 32 + l] == bits) peers |= 1u << l;
    c->warp_bar[w].wait();
    return peers;
}
inline unsigned __ballot_sync(unsigned, int) { emu::die("__ballot_sync is not emulated"); }
template <class T> inline T emu_reduce_add(unsigned mask, T v) {
    if (mask != 0xffffffffu) emu::die("__reduce_add_sync: only the full mask is emulated");
    emu::Cta* c = emu::cta;
    const int w = emu::tid >> 5, nl = c->warp_bar[w].n;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    c->slot[emu::tid] = bits;
    c->warp_bar[w].wait();
    T sum = 0;
    for (int l = 0; l < nl; ++l) { T x; std::memcpy(&x, &c->slot[w * 32 + l], sizeof(T)); sum += x; }
    c->warp_bar[w].wait();
    return sum;
}
inline int __reduce_add_sync(unsigned mask, int v) { return emu_reduce_add(mask, v); }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emu_reduce_add(mask, v); }

inline void emu_sleep_ms(int ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- atomics -------------------------------------------------------------------------------------------------------
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (o > v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicCAS(int* p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
inline float atomicAdd(float* p, float v) {
    unsigned* u = reinterpret_cast<unsigned*>(p);
    unsigned o = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        float f; std::memcpy(&f, &o, 4);
        const float nf = f + v;
        unsigned n; std::memcpy(&n, &nf, 4);
        if (__atomic_compare_exchange_n(u, &o, n, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
    }
}
inline double atomicAdd(double* p, double v) {
    unsigned long long* u = reinterpret_cast<unsigned long long*>(p);
    for (;;) {
        unsigned long long o = __atomic_load_n(u, __ATOMIC_RELAXED);
        double f; std::memcpy(&f, &o, 8);
        const double nf = f + v;
        unsigned long long n; std::memcpy(&n, &nf, 8);
        if (__atomic_compare_exchange_n(u, &o, n, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
    }
}
inline float4 atomicAdd(float4* p, float4 v) {       // red.global.add.v4.f32: four independent fp32 additions
    float4 o;
    o.x = atomicAdd(&p->x, v.x); o.y = atomicAdd(&p->y, v.y); o.z = atomicAdd(&p->z, v.z); o.w = atomicAdd(&p->w, v.w);
    return o;
}

// ---- arithmetic intrinsics (IEEE round-to-nearest; build with -ffp-contract=off) --------------------------------------
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline int __float2int_rz(float a) { return (int)a; }
inline int __float_as_int(float a) { int r; std::memcpy(&r, &a, 4); return r; }
inline unsigned __float_as_uint(float a) { unsigned r; std::memcpy(&r, &a, 4); return r; }
inline float __int_as_float(int a) { float r; std::memcpy(&r, &a, 4); return r; }
inline float __uint_as_float(unsigned a) { float r; std::memcpy(&r, &a, 4); return r; }
inline int __popc(unsigned a) { return __builtin_popcount(a); }
inline int __ffs(int a) { return __builtin_ffs(a); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
using std::isfinite;
using std::isnan;
using std::isinf;

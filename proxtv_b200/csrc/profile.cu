// profile.cu -- launch counting and optional per-kernel-class CUDA-event timing (used by bench.py for the roofline).
// Events are recorded on the stream the kernels are launched on; reading the totals synchronises those events.
#include "ptv_internal.h"
#include <atomic>
#include <mutex>
#include <vector>

namespace ptv {

static std::atomic<long long> g_launches[KC_COUNT];
static std::atomic<int> g_prof_on{0};
static std::mutex g_pm;
struct Span { cudaEvent_t a, b; int cls; };
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

KernelSpan::KernelSpan(int cls_, int nkernels, cudaStream_t st_) : cls(cls_), st(st_), a(nullptr) {
    g_launches[cls].fetch_add(nkernels, std::memory_order_relaxed);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        std::lock_guard<std::mutex> lk(g_pm);
        a = get_event();
        cudaEventRecord(a, st);
    }
}
void KernelSpan::cancel() {
    g_launches[cls].fetch_sub(1, std::memory_order_relaxed);
    if (a) { std::lock_guard<std::mutex> lk(g_pm); g_pool.push_back(a); a = nullptr; }
}
KernelSpan::~KernelSpan() {
    if (a) {
        std::lock_guard<std::mutex> lk(g_pm);
        cudaEvent_t b = get_event();
        cudaEventRecord(b, st);
        g_spans.push_back(Span{a, b, cls});
    }
}

void profile_enable(int on) { g_prof_on.store(on); }
void profile_reset() {
    std::lock_guard<std::mutex> lk(g_pm);
    for (auto& s : g_spans) { g_pool.push_back(s.a); g_pool.push_back(s.b); }
    g_spans.clear();
    for (int i = 0; i < KC_COUNT; i++) g_launches[i].store(0);
}
void profile_read(double* ms, long long* launches, long long* spans) {
    std::lock_guard<std::mutex> lk(g_pm);
    for (int i = 0; i < KC_COUNT; i++) { ms[i] = 0; spans[i] = 0; launches[i] = g_launches[i].load(); }
    for (auto& s : g_spans) {
        float t = 0;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
            ms[s.cls] += t; spans[s.cls]++;
        }
    }
}

}  // namespace ptv

namespace ptv {
// helpers for CUDA-graph replays (solver.cu): a replay launches the captured kernels without passing through KernelSpan
bool profile_is_enabled() { return g_prof_on.load() != 0; }
void profile_counters(long long* out) { for (int i = 0; i < KC_COUNT; i++) out[i] = g_launches[i].load(); }
void profile_add(const long long* delta) { for (int i = 0; i < KC_COUNT; i++) g_launches[i].fetch_add(delta[i]); }
}  // namespace ptv

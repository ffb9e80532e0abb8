// cuda_sim.cpp — CTA emulator runtime (see cuda_sim.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_sim.h"

#include <atomic>
#include <climits>
#include <memory>

#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace hbsim {

thread_local hbsim_dim3 t_threadIdx{0, 0, 0}, t_blockIdx{0, 0, 0}, t_blockDim{1, 1, 1}, t_gridDim{1, 1, 1};

constexpr size_t kMaxCluster = 8;
static std::vector<hbsim_u64> g_shared[kMaxCluster];
static thread_local size_t t_rank = 0; // CTA rank within its cluster
hbsim_u64 *shared_u64() { return g_shared[t_rank].data(); }
hbsim_u64 *shared_u64_of(size_t cluster_rank) { return g_shared[cluster_rank].data(); }

// A reusable sense-reversing barrier on a futex: arrivals are one atomic increment (no mutex to queue on — a CTA of the
// emulator is hundreds of host threads), waiters sleep on the generation word and the last arrival wakes them.
struct Barrier {
    std::atomic<unsigned> waiting{0};
    std::atomic<unsigned> gen{0};
    unsigned count = 0;
    void reset(size_t n) {
        count = (unsigned)n;
        waiting.store(0, std::memory_order_relaxed);
    }
    void wait() {
        const unsigned g = gen.load(std::memory_order_acquire);
        if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == count) {
            waiting.store(0, std::memory_order_relaxed);
            gen.store(g + 1, std::memory_order_release);
            syscall(SYS_futex, reinterpret_cast<unsigned *>(&gen), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
        } else {
            while (gen.load(std::memory_order_acquire) == g)
                syscall(SYS_futex, reinterpret_cast<unsigned *>(&gen), FUTEX_WAIT_PRIVATE, g, nullptr, nullptr, 0);
        }
    }
};

static Barrier g_sync[kMaxCluster]; // __syncthreads(), one per CTA of the running cluster
static Barrier g_cluster;           // barrier.cluster
static thread_local bool t_in_team = false;

void sync_threads() {
    if (t_in_team) g_sync[t_rank].wait();
}
void cluster_sync() {
    if (t_in_team) g_cluster.wait();
}

static inline void futex_wait(std::atomic<unsigned> &w, unsigned seen) {
    syscall(SYS_futex, reinterpret_cast<unsigned *>(&w), FUTEX_WAIT_PRIVATE, seen, nullptr, nullptr, 0);
}
static inline void futex_wake_all(std::atomic<unsigned> &w) {
    syscall(SYS_futex, reinterpret_cast<unsigned *>(&w), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
}

// Persistent team of worker threads, grown on demand; each launch hands them one body.  Start and completion are futex words
// too: a launch is one wake-all and one sleep of the launching thread, whatever the size of the team.
struct Team {
    std::vector<std::thread> workers;
    std::atomic<unsigned> epoch{0}, done{0}, finished{0};
    std::atomic<bool> quit{false};
    size_t active = 0, team_size = 0;
    const std::function<void()> *body = nullptr;
    size_t grid = 0, block = 0, cluster = 1;

    void worker(size_t id, unsigned seen) {
        for (;;) {
            while (epoch.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_acquire)) futex_wait(epoch, seen);
            if (quit.load(std::memory_order_acquire)) return;
            seen = epoch.load(std::memory_order_acquire);
            if (id < active) run_body(id);
            // every worker, active in this launch or not, reports: the launcher rewrites the fields only when nobody reads them
            if (done.fetch_add(1, std::memory_order_acq_rel) + 1 == team_size) {
                finished.store(seen, std::memory_order_release);
                futex_wake_all(finished);
            }
        }
    }
    void run_body(size_t id) {
        {
            t_in_team = true;
            t_rank = id / block;
            t_blockDim = {(unsigned)block, 1, 1};
            t_gridDim = {(unsigned)grid, 1, 1};
            t_threadIdx = {(unsigned)(id % block), 0, 0};
            for (size_t b = 0; b < grid; b += cluster) { // clusters run one after another
                t_blockIdx = {(unsigned)(b + t_rank), 0, 0};
                (*body)();
                g_cluster.wait();
            }
            t_in_team = false;
        }
    }

    void run(size_t g, size_t blk, size_t cl, const std::function<void()> &f) {
        const unsigned now = epoch.load(std::memory_order_relaxed);
        while (workers.size() < blk * cl) {
            const size_t id = workers.size();
            workers.emplace_back([this, id, now] { worker(id, now); });
        }
        grid = g;
        block = blk;
        body = &f;
        cluster = cl;
        active = blk * cl;
        team_size = workers.size();
        done.store(0, std::memory_order_relaxed);
        for (auto &b : g_sync) b.reset(blk);
        g_cluster.reset(blk * cl);
        epoch.store(now + 1, std::memory_order_release); // publishes the fields above
        futex_wake_all(epoch);
        while (finished.load(std::memory_order_acquire) != now + 1) futex_wait(finished, finished.load(std::memory_order_acquire));
    }

    ~Team() {
        quit.store(true, std::memory_order_release);
        futex_wake_all(epoch);
        for (auto &t : workers) t.join();
    }
};

static Team &team() {
    static Team *t = new Team(); // intentionally leaked: workers must outlive static destructors
    return *t;
}

void launch(size_t grid, size_t block, size_t smem_bytes, int sync, const std::function<void()> &body, size_t cluster) {
    if (grid == 0 || block == 0) return;
    if (cluster < 1 || cluster > kMaxCluster || grid % cluster) abort();
    for (auto &sh : g_shared)
        if (sh.size() * 8 < smem_bytes + 64) sh.assign(smem_bytes / 8 + 8, 0);
    if (sync || cluster > 1) {
        team().run(grid, block, cluster, body);
        return;
    }
    t_rank = 0;
    t_blockDim = {(unsigned)block, 1, 1};
    t_gridDim = {(unsigned)grid, 1, 1};
    for (size_t b = 0; b < grid; b++) {
        t_blockIdx = {(unsigned)b, 0, 0};
        for (size_t t = 0; t < block; t++) {
            t_threadIdx = {(unsigned)t, 0, 0};
            body();
        }
    }
}

} // namespace hbsim

// cuda_sim.cpp — CTA emulator runtime (see cuda_sim.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_sim.h"

#include <atomic>
#include <memory>

namespace hbsim {

thread_local hbsim_dim3 t_threadIdx{0, 0, 0}, t_blockIdx{0, 0, 0}, t_blockDim{1, 1, 1}, t_gridDim{1, 1, 1};

constexpr size_t kMaxCluster = 8;
static std::vector<hbsim_u64> g_shared[kMaxCluster];
static thread_local size_t t_rank = 0; // CTA rank within its cluster
hbsim_u64 *shared_u64() { return g_shared[t_rank].data(); }
hbsim_u64 *shared_u64_of(size_t cluster_rank) { return g_shared[cluster_rank].data(); }

// A reusable sense-reversing barrier.
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    size_t count = 0, waiting = 0;
    unsigned long gen = 0;
    void reset(size_t n) {
        count = n;
        waiting = 0;
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long g = gen;
        if (++waiting == count) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
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

// Persistent team of worker threads, grown on demand; each launch hands them one body.
struct Team {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    unsigned long epoch = 0;
    size_t active = 0, done = 0;
    const std::function<void()> *body = nullptr;
    size_t grid = 0, block = 0, cluster = 1;
    bool quit = false;

    void worker(size_t id) {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                if (id >= active) continue;
            }
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
            {
                std::lock_guard<std::mutex> lk(m);
                if (++done == active) cv_done.notify_all();
            }
        }
    }

    void run(size_t g, size_t blk, size_t cl, const std::function<void()> &f) {
        while (workers.size() < blk * cl) {
            const size_t id = workers.size();
            workers.emplace_back([this, id] { worker(id); });
        }
        {
            std::lock_guard<std::mutex> lk(m);
            grid = g;
            block = blk;
            body = &f;
            cluster = cl;
            active = blk * cl;
            done = 0;
            for (auto &b : g_sync) b.reset(blk);
            g_cluster.reset(blk * cl);
            epoch++;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return done == active; });
    }

    ~Team() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv_go.notify_all();
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

// multi.cu — one process, one host thread + one context + own streams per GPU.
//
// The pairs of a sequence are independent (pair i needs frames i and i + 1 only), so a sequence is cut into contiguous
// pair ranges, one per GPU, each with a one-frame halo. Every GPU has a persistent worker thread that owns its vb_ctx and
// drives it through vb_pairs_submit / vb_pairs_wait; the workers' downloads land directly in the caller's result arrays
// at the range's position — that IS the host gather. There is no collective and no NCCL: nothing is exchanged between
// GPUs. Pair i samples with std::mt19937(seed0 + i) wherever it runs, so the output is independent of the GPU count.
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "pairs_dev.cuh"

namespace vb {
int pairs_submit(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                 const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                 uint64_t cap, uint64_t base, int *ticket);
int pairs_wait(vb_ctx *ctx, int ticket, uint64_t *total_matches);

struct Worker {
    int device = 0;
    vb_ctx *ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> jobs;
    bool stop = false;

    void loop() {
        cudaSetDevice(device);
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;
                job = std::move(jobs.front());
                jobs.pop_front();
            }
            job();
        }
    }
    void post(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            jobs.push_back(std::move(f));
        }
        cv.notify_one();
    }
};

// Completion of one fan-out: every worker reports its status and error text.
struct Fan {
    std::mutex mu;
    std::condition_variable cv;
    uint32_t pending = 0;
    int rc = VB_OK;
    std::string err;
    void done(int r) {
        std::lock_guard<std::mutex> lk(mu);
        if (r != VB_OK && rc == VB_OK) { rc = r; err = vb_last_error(); }
        if (--pending == 0) cv.notify_all();
    }
    int join() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return pending == 0; });
        if (rc != VB_OK) set_error("%s", err.c_str());
        return rc;
    }
};

struct MultiTicket {
    bool live = false;
    std::vector<int> local;        // per worker: its context's ticket, or -1 when the worker got no pairs
    std::vector<uint64_t> totals;
};

}  // namespace vb

struct vb_multi {
    std::vector<vb::Worker *> workers;
    vb::MultiTicket tickets[vb::PAIRS_DEPTH];
    int next_ticket = 0;
};

using namespace vb;

extern "C" {

int vb_multi_destroy(vb_multi *m) {
    if (!m) return VB_OK;
    for (Worker *w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->stop = true;
        }
        w->cv.notify_one();
        if (w->th.joinable()) w->th.join();
        if (w->ctx) vb_destroy(w->ctx);
        delete w;
    }
    delete m;
    return VB_OK;
}

int vb_multi_create(const int *devices, uint32_t ndev, vb_multi **out) {
    VB_REQUIRE(devices && out && ndev > 0, VB_ERR_INVALID, "NULL argument or no devices");
    vb_multi *m = new vb_multi();
    for (uint32_t i = 0; i < ndev; i++) {
        Worker *w = new Worker();
        w->device = devices[i];
        m->workers.push_back(w);
        const int rc = vb_create(devices[i], &w->ctx);
        if (rc != VB_OK) {
            vb_multi_destroy(m);
            return rc;
        }
        w->th = std::thread([w] { w->loop(); });
    }
    *out = m;
    return VB_OK;
}

uint32_t vb_multi_device_count(const vb_multi *m) { return m ? (uint32_t)m->workers.size() : 0; }

int vb_multi_pairs_submit(vb_multi *m, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                          const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                          uint64_t cap_matches, int *ticket) {
    VB_REQUIRE(m && pts && desc && params && results && ticket, VB_ERR_INVALID, "NULL argument");
    MultiTicket &t = m->tickets[m->next_ticket % PAIRS_DEPTH];
    VB_REQUIRE(!t.live, VB_ERR_CAPACITY, "three submissions are already in flight: vb_multi_pairs_wait the oldest ticket first");
    const uint32_t P = nframes < 2 ? 0 : nframes - 1, nw = (uint32_t)m->workers.size();
    // every range keeps its matches in its own [first * k, (first + count) * k) window of matches16, so the window a GPU
    // writes is known before any GPU has finished
    VB_REQUIRE(matches16 == nullptr || cap_matches >= (uint64_t)P * k, VB_ERR_CAPACITY,
               "vb_multi needs room for (nframes - 1) * k matches (ranges are written in place, not packed across GPUs)");
    t.local.assign(nw, -1);
    t.totals.assign(nw, 0);
    Fan fan;
    fan.pending = nw;
    for (uint32_t w = 0; w < nw; w++) {
        const uint32_t first = (uint32_t)((uint64_t)P * w / nw), last = (uint32_t)((uint64_t)P * (w + 1) / nw);
        Worker *wk = m->workers[w];
        int *slot = &t.local[w];
        wk->post([=, &fan] {
            int rc = VB_OK;
            if (last > first) {
                vb_pair_params prm = *params;
                prm.seed0 = params->seed0 + first;
                rc = pairs_submit(wk->ctx, pts + (size_t)first * k * 2, desc + (size_t)first * k * bytes, last - first + 1, k, bytes,
                                  &prm, results + first, match_offsets ? match_offsets + first : nullptr, matches16,
                                  (uint64_t)(last - first) * k, (uint64_t)first * k, slot);
            }
            fan.done(rc);
        });
    }
    const int rc = fan.join();
    if (rc != VB_OK) {   // release whatever did get submitted
        for (uint32_t w = 0; w < nw; w++)
            if (t.local[w] >= 0) pairs_wait(m->workers[w]->ctx, t.local[w], nullptr);
        return rc;
    }
    t.live = true;
    *ticket = m->next_ticket++;
    return VB_OK;
}

int vb_multi_pairs_wait(vb_multi *m, int ticket, uint64_t *total_matches) {
    VB_REQUIRE(m && ticket >= 0, VB_ERR_INVALID, "unknown ticket");
    MultiTicket &t = m->tickets[ticket % PAIRS_DEPTH];
    VB_REQUIRE(t.live, VB_ERR_INVALID, "unknown or already completed ticket");
    t.live = false;
    const uint32_t nw = (uint32_t)m->workers.size();
    Fan fan;
    fan.pending = nw;
    for (uint32_t w = 0; w < nw; w++) {
        Worker *wk = m->workers[w];
        const int lt = t.local[w];
        uint64_t *tot = &t.totals[w];
        wk->post([=, &fan] { fan.done(lt >= 0 ? pairs_wait(wk->ctx, lt, tot) : VB_OK); });
    }
    const int rc = fan.join();
    uint64_t total = 0;
    for (uint64_t v : t.totals) total += v;
    if (total_matches) *total_matches = total;
    return rc;
}

int vb_multi_pairs_run(vb_multi *m, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                       const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                       uint64_t cap_matches, uint64_t *total_matches) {
    int ticket = -1;
    const int rc = vb_multi_pairs_submit(m, pts, desc, nframes, k, bytes, params, results, match_offsets, matches16, cap_matches,
                                         &ticket);
    if (rc) return rc;
    return vb_multi_pairs_wait(m, ticket, total_matches);
}

}  // extern "C"

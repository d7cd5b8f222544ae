// adapter_common.h — process-wide context shared by the KDTree / RansacFilter adapters.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/vslam_b200.h"

namespace vslam_b200_adapter {

// One lazily created context per process. The reference runs this path on a single thread
// (SURVEY §8b); the mutex only protects creation and the side tables.
struct Global {
    std::mutex mu;
    vb_ctx *ctx = nullptr;
    int device = -1;
};

inline Global &global() {
    static Global g;
    return g;
}

inline vb_ctx *context() {
    Global &g = global();
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.ctx) {
        int dev = g.device;
        if (dev < 0) {
            const char *e = std::getenv("VSLAM_B200_DEVICE");
            dev = e ? std::atoi(e) : 0;
        }
        if (vb_create(dev, &g.ctx) != VB_OK)   // no CPU fallback: fail loudly
            throw std::runtime_error(std::string("vslam_b200: cannot create GPU context: ") + vb_last_error());
        g.device = dev;
    }
    return g.ctx;
}

inline void check(int rc, const char *what) {
    if (rc != VB_OK) throw std::runtime_error(std::string("vslam_b200: ") + what + ": " + vb_last_error());
}

}  // namespace vslam_b200_adapter

// backend.h — process-wide handle on the B200 back end (the C ABI of include/hehub_b200.h).
//
// The reference is a single-threaded library with global caches (ntt.cpp:107-115); the mirror
// keeps that model: one lazily created context on device HEHUB_B200_DEVICE (default 0), used by
// every hehub:: call.  Status codes coming back over the C ABI are turned into the exception
// types the reference throws for the same condition.  There is no CPU path: without the CUDA
// library / a device the first call throws std::runtime_error.
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../../include/hehub_b200.h"

namespace hehub {
namespace b200 {

inline hehub_b200_ctx *&context_slot() {
    static hehub_b200_ctx *ctx = nullptr;
    return ctx;
}

inline hehub_b200_ctx *context() {
    hehub_b200_ctx *&ctx = context_slot();
    if (!ctx) {
        const char *dev = std::getenv("HEHUB_B200_DEVICE");
        const int rc = hehub_b200_ctx_create(&ctx, dev ? std::atoi(dev) : 0, nullptr);
        if (rc != HEHUB_B200_OK || !ctx)
            throw std::runtime_error("hehub_b200: cannot create a CUDA context (no device or driver); there is no CPU fallback");
    }
    return ctx;
}

/// Use an externally created context (e.g. one per GPU in a multi-device process).
inline void set_context(hehub_b200_ctx *ctx) { context_slot() = ctx; }

inline void check(int rc) {
    if (rc == HEHUB_B200_OK) return;
    const std::string msg = hehub_b200_last_error(context());
    switch (rc) {
    case HEHUB_B200_ERR_INVALID: throw std::invalid_argument(msg);
    case HEHUB_B200_ERR_UNSUPPORTED: {
        static thread_local std::string keep; // the reference throws bare `const char *`
        keep = msg;
        throw keep.c_str();
    }
    case HEHUB_B200_ERR_NOMEM: throw std::bad_alloc();
    default: throw std::runtime_error("hehub_b200: " + msg);
    }
}

inline void synchronize() { check(hehub_b200_ctx_synchronize(context())); }

} // namespace b200
} // namespace hehub

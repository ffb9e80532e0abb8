// cpp_api_latency.cpp — what an application written against the reference's C++ API sees per call: one
// hehub::ckks::mult (tensor + relinearize) on device-resident ciphertexts at the C3 shape, through the host
// mirror in hehub_b200/cpp/hehub, one ciphertext pair at a time.  Prints one JSON line.
//   g++ -std=c++17 -O2 -Ihehub_b200/cpp tools/cpp_api_latency.cpp hehub_b200/libhehub_b200.so -Wl,-rpath,$PWD/hehub_b200 -o tools/cpp_api_latency
#include <chrono>
#include <cstdio>
#include <vector>

#include "hehub/hehub.h"
using namespace hehub;

static RnsPolynomial filled(size_t n, const std::vector<u64> &mods, u64 seed, PolyRepForm form) {
    RnsPolynomial p(n, mods.size(), mods);
    for (size_t k = 0; k < mods.size(); k++) {
        u64 s = seed + k;
        for (size_t i = 0; i < n; i++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            p[(int)k][i] = s % mods[k];
        }
    }
    p.rep_form = form;
    return p;
}

int main() {
    const size_t n = 8192;
    const std::vector<u64> mods = {1099507695617ull, 1073479681ull, 1072496641ull, 1071513601ull};
    std::vector<u64> ext(mods);
    ext.push_back(1099510054913ull);
    ckks::CkksCt ct1(RlweCt{filled(n, mods, 100, PolyRepForm::value), filled(n, mods, 110, PolyRepForm::value)});
    ckks::CkksCt ct2(RlweCt{filled(n, mods, 200, PolyRepForm::value), filled(n, mods, 210, PolyRepForm::value)});
    ct1.scaling_factor = ct2.scaling_factor = 1073741824.0;
    RlweKsk key;
    for (size_t r = 0; r < mods.size(); r++)
        key.push_back(RlweCt{filled(n, ext, 1000 + 100 * r, PolyRepForm::value), filled(n, ext, 1010 + 100 * r, PolyRepForm::value)});
    for (int i = 0; i < 20; i++) auto warm = ckks::mult(ct1, ct2, key);
    b200::synchronize();
    const int reps = 500;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) {
        auto prod = ckks::mult(ct1, ct2, key);
    }
    b200::synchronize();
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    // operands that are themselves results of earlier calls: their halves already sit back to back in one slab
    const auto r1 = ckks::mult(ct1, ct2, key), r2 = ckks::mult(ct2, ct1, key);
    b200::synchronize();
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) {
        auto prod = ckks::mult(r1, r2, key);
    }
    b200::synchronize();
    const double us_chain = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) {
        ckks::CkksCt rs = ct1;
        rs.scaling_factor = ct1.scaling_factor;
        ckks::rescale_inplace(rs);
    }
    b200::synchronize();
    const double us_rs = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("{\"shape\": \"C3 (N=8192, L=4)\", \"ckks_mult_us_per_call\": %.2f, \"ckks_mult_on_earlier_results_us_per_call\": %.2f, "
                "\"copy_plus_rescale_inplace_us_per_call\": %.2f, \"calls\": %d}\n", us, us_chain, us_rs, reps);
    return 0;
}

// cpp_api_bench.cpp — the reference's own application-level benchmark (bench/benchmarks.cpp, bench/ckks_bm.cpp, README table:
// CKKS encode + encrypt, decrypt + decode, rotation) written against the C++ mirror in hehub_b200/cpp/hehub, same parameter
// sets: ckks::create_params(N, scaling_bits) for (2^12, 36), (2^13, 43), (2^14, 48), (2^15, 55).  One call at a time, host
// data in and out as an application would have it (the encoder's FFT and the samplers run on the host, the transforms and
// coefficient-wise arithmetic on the GPU).  Prints one JSON line per parameter set.
//   g++ -std=c++17 -O2 -Ihehub_b200/cpp tools/cpp_api_bench.cpp hehub_b200/libhehub_b200.so -Wl,-rpath,$PWD/hehub_b200 -o tools/cpp_api_bench
#include <chrono>
#include <cstdio>
#include <utility>
#include <vector>

#include "hehub/hehub.h"
using namespace hehub;

template <class F>
static double us_per_call(int reps, F f) {
    for (int i = 0; i < 3; i++) f();
    b200::synchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) f();
    b200::synchronize();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main() {
    const std::pair<int, int> sets[] = {{12, 36}, {13, 43}, {14, 48}, {15, 55}};
    for (const auto &[logn, scaling_bits] : sets) {
        const size_t n = (size_t)1 << logn;
        auto params = ckks::create_params(n, (size_t)scaling_bits);
        CkksSk sk(params);
        auto rot_key = get_rot_key(sk, params.additional_mod, 1);
        auto relin_key = get_relin_key(sk, params.additional_mod);
        std::vector<cc_double> data(n / 2);
        for (size_t i = 0; i < data.size(); i++) data[i] = cc_double(0.001 * (double)(i % 997), -0.002 * (double)(i % 499));
        auto ct = ckks::encrypt(ckks::simd_encode(data, params), sk);
        const int reps = logn <= 13 ? 200 : 50;
        const double enc = us_per_call(reps, [&] { auto c = ckks::encrypt(ckks::simd_encode(data, params), sk); });
        const double dec = us_per_call(reps, [&] { auto d = ckks::simd_decode<cc_double>(ckks::decrypt(ct, sk)); });
        const double rot = us_per_call(reps, [&] { auto r = ckks::rotate(ct, rot_key, 1); });
        const double mul = us_per_call(reps, [&] { auto m = ckks::mult(ct, ct, relin_key); });
        auto back = ckks::simd_decode<cc_double>(ckks::decrypt(ct, sk));
        double worst = 0;
        for (size_t i = 0; i < data.size(); i++) worst = std::max(worst, std::abs(back[i] - data[i]));
        std::printf("{\"N\": %zu, \"scaling_bits\": %d, \"limbs\": %zu, \"encode_encrypt_us\": %.1f, \"decrypt_decode_us\": %.1f, "
                    "\"rotate_us\": %.1f, \"mult_relin_us\": %.1f, \"decode_max_abs_error\": %.3g}\n",
                    n, scaling_bits, params.moduli.size(), enc, dec, rot, mul, worst);
    }
    return 0;
}

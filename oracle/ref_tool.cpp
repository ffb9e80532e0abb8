/*
 * ref_tool.cpp — stand-alone driver for the parts of the UNMODIFIED reference that cannot run inside the Python process.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's big integers go through std::stringstream (bigint.cpp:35-46, 295-302); inside a
 * Python interpreter that has already loaded another libstdc++ (numpy) that path crashes in locale code, so
 * oracle/make_golden.py runs these scenarios in this separate executable (built by `make ref` into oracle/_ref/, linked against
 * oracle/_ref/libhehub_ref.so) and reads one JSON line per scenario from its stdout.
 *
 *   ref_tool codec <logn> <additional_bits> <log2_scaling> <seed> <count> <bits...>
 *   ref_tool apibench <logn> <scaling_bits> <reps>     the reference's application benchmark on this host (tools/gpu_api_bench.sh)
 */
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int ref_ckks_codec_scenario(unsigned logn, size_t L, const unsigned *moduli_bits, unsigned additional_bits, double log2_scaling,
                                       uint64_t seed, size_t count, uint64_t *hash, double *decoded);

extern "C" int ref_api_bench(unsigned logn, unsigned scaling_bits, int reps, double *out_us);

int main(int argc, char **argv) {
    if (argc == 5 && !std::strcmp(argv[1], "apibench")) {
        double us[5] = {0, 0, 0, 0, 0};
        const unsigned logn = (unsigned)std::atoi(argv[2]), bits = (unsigned)std::atoi(argv[3]);
        const int rc = ref_api_bench(logn, bits, std::atoi(argv[4]), us);
        if (rc) return 10 + rc;
        std::printf("{\"impl\": \"reference (one host core)\", \"N\": %u, \"scaling_bits\": %u, \"limbs\": %d, \"encode_encrypt_us\": %.1f, "
                    "\"decrypt_decode_us\": %.1f, \"rotate_us\": %.1f, \"mult_relin_us\": %.1f}\n",
                    1u << logn, bits, (int)us[4], us[0], us[1], us[2], us[3]);
        return 0;
    }
    if (argc >= 8 && !std::strcmp(argv[1], "codec")) {
        const unsigned logn = (unsigned)std::atoi(argv[2]), add = (unsigned)std::atoi(argv[3]);
        const double log2s = std::atof(argv[4]);
        const uint64_t seed = std::strtoull(argv[5], nullptr, 10);
        const size_t count = (size_t)std::strtoull(argv[6], nullptr, 10);
        std::vector<unsigned> bits;
        for (int i = 7; i < argc; i++) bits.push_back((unsigned)std::atoi(argv[i]));
        uint64_t hash = 0;
        std::vector<double> dec((size_t)1 << logn, 0.0);
        const int rc = ref_ckks_codec_scenario(logn, bits.size(), bits.data(), add, log2s, seed, count, &hash, dec.data());
        if (rc) return 10 + rc;
        double abs_sum = 0;
        for (double v : dec) abs_sum += v < 0 ? -v : v;
        std::printf("{\"plaintext\": \"%016" PRIx64 "\", \"decoded_abs_sum\": %.17g, \"decoded_head\": [", hash, abs_sum);
        for (int i = 0; i < 16 && i < (int)dec.size(); i++) std::printf("%s%.17g", i ? ", " : "", dec[i]);
        std::printf("]}\n");
        return 0;
    }
    std::fprintf(stderr, "usage: ref_tool codec <logn> <additional_bits> <log2_scaling> <seed> <count> <bits...>\n");
    return 2;
}

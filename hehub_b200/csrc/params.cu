// params.cu — parameter selection of the reference, host-only (no device work).
//
//   prime_lists[bits][i]        src/fhe/common/primelists.cpp:5-192
//   hehub::create_params        src/fhe/primitives/rlwe.cpp:9-29
//   ckks::create_params (x2)    src/fhe/ckks/basics.cpp:14-64
//
// The reference ships its primes as a literal table.  The table is the output of one rule — row `bits`
// (27 <= bits <= 59) lists, in descending order, the 20 largest primes below 2^bits that are congruent to 1
// modulo 2^16 — so the rule is restated here instead of the 660 literals, together with the three places where
// the shipped table departs from its own rule (a drop-in must hand out the SAME moduli in the SAME order, or
// every ciphertext made from `create_params` differs from the reference's):
//   * row 45 has 19 entries: the rule's 18th prime (35184351313921) is absent, later entries move up by one;
//   * row 57, entry 12 and row 58, entry 16 lost their leading decimal digit (144115188062617601 is listed as
//     44115188062617601, 288230376128839681 as 88230376128839681).  Both listed values are composite; the
//     reference hands them out all the same, and so does this function — the transforms then reject them
//     when the tables are built (no primitive 2N-th root), which is also where the reference breaks.
#include <mutex>
#include <vector>

#include "../../include/hehub_b200.h"
#include "context.h"

namespace hb {

typedef unsigned __int128 u128;

static bool is_prime_u64(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37}; // deterministic for every n < 2^64
    if (n < 2) return false;
    for (u64 b : bases) {
        if (n == b) return true;
        if (n % b == 0) return false;
    }
    u64 d = n - 1;
    unsigned r = 0;
    while (!(d & 1)) {
        d >>= 1;
        r++;
    }
    for (u64 b : bases) {
        u64 x = host_pow_mod(n, b, d);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (unsigned j = 1; j < r && witness; j++) {
            x = (u64)((u128)x * x % n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

constexpr unsigned kFirstRow = 27, kLastRow = 59, kRowLength = 20;

static const std::vector<std::vector<u64>> &prime_rows() {
    static std::vector<std::vector<u64>> rows;
    static std::once_flag once;
    std::call_once(once, [] {
        rows.assign(kLastRow + 1, {});
        for (unsigned bits = kFirstRow; bits <= kLastRow; bits++) {
            auto &row = rows[bits];
            for (u64 c = ((u64)1 << bits) - 65536 + 1; row.size() < kRowLength && c > ((u64)1 << (bits - 1)); c -= 65536)
                if (is_prime_u64(c)) row.push_back(c);
        }
        // the shipped table's departures from the rule (see the header of this file)
        rows[45].erase(rows[45].begin() + 17);
        rows[57][12] -= 100000000000000000ull;
        rows[58][16] -= 200000000000000000ull;
    });
    return rows;
}

} // namespace hb

extern "C" {

int hehub_b200_prime_row(unsigned bits, size_t capacity, uint64_t *out) {
    const auto &rows = hb::prime_rows();
    if (bits >= rows.size()) return 0;
    const auto &row = rows[bits];
    for (size_t i = 0; i < row.size() && i < capacity; i++) out[i] = row[i];
    return (int)row.size();
}

int hehub_b200_pick_moduli(const unsigned *moduli_bits, size_t L, unsigned additional_bits, uint64_t *moduli_out,
                           uint64_t *additional_out) {
    // one cursor per row, shared by the additional modulus (drawn FIRST, basics.cpp:29) and the chain (:30-33);
    // additional_bits == 0 with additional_out == NULL gives hehub::create_params (rlwe.cpp:9-29), which has no such modulus
    if ((L && (!moduli_bits || !moduli_out))) return HEHUB_B200_ERR_INVALID;
    const auto &rows = hb::prime_rows();
    std::vector<size_t> cursor(64, 0);
    auto next = [&](unsigned bits, uint64_t *dst) {
        // the reference indexes its table unchecked here (undefined behaviour past the end of a row, its try/catch never
        // fires); the replacement reports what that catch was meant to: "No suitable primes in the library."
        if (bits >= rows.size() || cursor[bits] >= rows[bits].size()) return false;
        *dst = rows[bits][cursor[bits]++];
        return true;
    };
    if (additional_out && !next(additional_bits, additional_out)) return HEHUB_B200_ERR_UNSUPPORTED;
    for (size_t k = 0; k < L; k++)
        if (!next(moduli_bits[k], moduli_out + k)) return HEHUB_B200_ERR_UNSUPPORTED;
    return HEHUB_B200_OK;
}

} // extern "C"

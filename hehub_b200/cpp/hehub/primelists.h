// primelists.h — hehub::prime_lists (src/fhe/common/primelists.{h,cpp}): prime_lists[bits][i] is the i-th prime the
// parameter constructors draw for a `bits`-bit modulus.  The values come from the back end
// (hehub_b200_prime_row, csrc/params.cu: the table's generating rule plus the shipped table's three irregular
// entries), so rlwe / ckks `create_params` hand out exactly the reference's moduli.
#pragma once
#include <vector>

#include "backend.h"

namespace hehub {

namespace detail {
inline std::vector<std::vector<uint64_t>> build_prime_lists() {
    std::vector<std::vector<uint64_t>> lists(60); // primelists.cpp: 60 rows, 27..59 populated
    for (unsigned bits = 0; bits < 60; bits++) {
        uint64_t row[32];
        const int n = hehub_b200_prime_row(bits, 32, row);
        lists[bits].assign(row, row + (n < 32 ? n : 32));
    }
    return lists;
}
} // namespace detail

inline const std::vector<std::vector<uint64_t>> prime_lists = detail::build_prime_lists();

} // namespace hehub

// permutation.h — Galois permutations on NTT-form polynomials (src/fhe/common/permutation.{h,cpp}).
#pragma once
#include "rns.h"

namespace hehub {

/// permutation.cpp:28-60 — X -> X^(3^step) on a polynomial in NTT value form
inline RnsPolynomial cycle(const RnsPolynomial &poly_ntt, const size_t step) {
    if (poly_ntt.rep_form != PolyRepForm::value) throw std::invalid_argument("poly_ntt is expected to be in NTT value form");
    RnsPolynomial cycled(poly_ntt.dimension(), poly_ntt.component_count(), poly_ntt.modulus_vec());
    cycled.rep_form = PolyRepForm::value;
    b200::check(hehub_b200_galois_cycle(b200::context(), (unsigned)poly_ntt.log_dimension(), poly_ntt.component_count(),
                                        poly_ntt.dev(), cycled.dev_mut(), step, 1));
    return cycled;
}
/// permutation.cpp:62-75 — X -> X^(-1)
inline RnsPolynomial involution(const RnsPolynomial &poly_ntt) {
    if (poly_ntt.rep_form != PolyRepForm::value) throw std::invalid_argument("poly_ntt is expected to be in NTT value form");
    RnsPolynomial inv(poly_ntt.dimension(), poly_ntt.component_count(), poly_ntt.modulus_vec());
    inv.rep_form = PolyRepForm::value;
    b200::check(hehub_b200_galois_involution(b200::context(), (unsigned)poly_ntt.log_dimension(), poly_ntt.component_count(),
                                             poly_ntt.dev(), inv.dev_mut(), 1));
    return inv;
}

} // namespace hehub

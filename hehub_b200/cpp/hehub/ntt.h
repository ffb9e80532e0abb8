// ntt.h — negacyclic NTT / INTT entry points of hehub (src/fhe/common/ntt.h:27-92) on the B200 back end.
#pragma once
#include "mod_arith.h"

namespace hehub {

/// ntt.cpp:145-176.  coeffs: HOST pointer (reference signature); one synchronising round trip.
inline void ntt_negacyclic_inplace_lazy(const size_t log_dimension, const u64 modulus, u64 coeffs[]) {
    detail::Staged s(coeffs, (size_t)1 << log_dimension);
    b200::check(hehub_b200_ntt_fwd_lazy(b200::context(), (unsigned)log_dimension, &modulus, 1, s.dev, 1));
    s.download(coeffs);
}
/// ntt.h:41-51 — all limbs of the polynomial in ONE launch, on the device
inline void ntt_negacyclic_inplace_lazy(RnsPolynomial &rns_poly) {
    // like the reference (ntt.h:41-51), the representation tag is set, not checked
    b200::check(hehub_b200_ntt_fwd_lazy(b200::context(), (unsigned)rns_poly.log_dimension(), rns_poly.modulus_vec().data(),
                                        rns_poly.component_count(), rns_poly.dev_mut(), 1));
    rns_poly.rep_form = PolyRepForm::value;
}
/// ntt.cpp:178-223
inline void intt_negacyclic_inplace_lazy(const size_t log_dimension, const u64 modulus, u64 values[]) {
    detail::Staged s(values, (size_t)1 << log_dimension);
    b200::check(hehub_b200_intt_lazy(b200::context(), (unsigned)log_dimension, &modulus, 1, s.dev, 1, 0));
    s.download(values);
}
/// ntt.h:72-82
inline void intt_negacyclic_inplace_lazy(RnsPolynomial &rns_poly) {
    b200::check(hehub_b200_intt_lazy(b200::context(), (unsigned)rns_poly.log_dimension(), rns_poly.modulus_vec().data(),
                                     rns_poly.component_count(), rns_poly.dev_mut(), 1, 0));
    rns_poly.rep_form = PolyRepForm::coeff;
}
/// ntt.h:89-92 — INTT followed by reduce_strict, fused into the transform's epilogue
inline void intt_negacyclic_inplace(RnsPolynomial &rns_poly) {
    b200::check(hehub_b200_intt_lazy(b200::context(), (unsigned)rns_poly.log_dimension(), rns_poly.modulus_vec().data(),
                                     rns_poly.component_count(), rns_poly.dev_mut(), 1, 1));
    rns_poly.rep_form = PolyRepForm::coeff;
}
/// ntt.cpp:225-231 — build and upload the twiddle tables ahead of time
inline void cache_ntt_factors_strict(const u64 log_dimension, const std::vector<u64> &moduli) {
    b200::check(hehub_b200_tables_prepare(b200::context(), (unsigned)log_dimension, moduli.data(), moduli.size()));
}

} // namespace hehub

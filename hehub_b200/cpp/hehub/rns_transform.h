// rns_transform.h — rns_base_transform of hehub (src/fhe/common/rns_transform.{h,cpp}:11-126) on the B200 back end.
#pragma once
#include <stdexcept>
#include <vector>

#include "mod_arith.h"
#include "rns.h"

namespace hehub {

/// rns_transform.cpp:107-126: coefficient-form input, one modulus -> many or many -> one.  The
/// many -> one direction is built for the reference's small-coefficient path (:47-84); inputs that
/// need the big-integer CRT composition (:86-105) throw `const char *`, like the reference's own
/// unimplemented case (:123).
inline RnsPolynomial rns_base_transform(RnsPolynomial input_rns_poly, const std::vector<u64> &new_moduli) {
    if (input_rns_poly.rep_form == PolyRepForm::value)
        throw std::logic_error("Trying to perform RNS base transformation on NTT values."); // :109-112
    const size_t n = input_rns_poly.dimension();
    if (input_rns_poly.component_count() == 1) {
        RnsPolynomial result(n, new_moduli.size(), new_moduli);
        b200::check(hehub_b200_rns_base_transform_from_single(b200::context(), input_rns_poly.modulus_at(0), new_moduli.data(),
                                                              new_moduli.size(), input_rns_poly.dev(), result.dev_mut(), n, 1));
        result.rep_form = PolyRepForm::coeff;
        return result;
    } else if (new_moduli.size() == 1) {
        RnsPolynomial result(n, 1, new_moduli);
        b200::check(hehub_b200_rns_base_transform_to_single(b200::context(), input_rns_poly.modulus_vec().data(),
                                                            input_rns_poly.component_count(), new_moduli[0], input_rns_poly.dev(),
                                                            result.dev_mut(), n, 1));
        result.rep_form = PolyRepForm::coeff;
        return result;
    }
    throw "under development"; // :123
}

} // namespace hehub

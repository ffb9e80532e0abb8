// rns.h — hehub::RnsIntVec / hehub::RnsPolynomial with device-resident storage.
//
// Mirrors the container of the reference (src/fhe/common/rns.h:15-156, rns.cpp:9-171): same
// constructors, accessors, value semantics, operators and exceptions.  Storage differs: instead
// of one pooled host block per limb (allocator.h:105-223) a polynomial owns ONE contiguous
// pooled device slab laid out [limb][N] (hehub_b200_slab_alloc), plus a host mirror that is
// materialised only when a caller touches raw words through operator[] / data().  Device
// operators never move data over PCIe; `poly[k].data()` keeps working for host-side code
// (encoders, samplers, tests) at the cost of one synchronising copy.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "backend.h"

namespace hehub {

using u64 = uint64_t;
using u128 = unsigned __int128;
using i128 = __int128;

namespace detail {
/// One pooled device slab, returned to the pool when the last owner lets go.  Polynomials produced together by
/// one kernel call (the two halves of a ciphertext) share a block, so they are already contiguous when the next
/// call needs them as one [polys][L][N] operand.
struct SlabBlock {
    u64 *p = nullptr;
    explicit SlabBlock(size_t words) {
        if (words) b200::check(hehub_b200_slab_alloc(b200::context(), words, &p));
    }
    SlabBlock(const SlabBlock &) = delete;
    SlabBlock &operator=(const SlabBlock &) = delete;
    ~SlabBlock() {
        if (p) hehub_b200_slab_free(b200::context(), p);
    }
};

} // namespace detail

class RnsIntVec {
public:
    struct Params {
        size_t dimension = 0;
        size_t component_count;
        std::vector<u64> moduli;
    };

    /// View of one RNS component's N words in host memory (the reference's SmartArray<u64>).
    class ComponentData {
    public:
        u64 *data() { return owner_->host_words(index_, true); }
        const u64 *data() const { return owner_->host_words(index_, false); }
        u64 &operator[](size_t i) { return data()[i]; }
        const u64 &operator[](size_t i) const { return data()[i]; }
        u64 *begin() { return data(); }
        u64 *end() { return data() + owner_->dimension_; }
        const u64 *begin() const { return data(); }
        const u64 *end() const { return data() + owner_->dimension_; }
        size_t size() const { return owner_->dimension_; }
        /// copies the N words of another component (any polynomial), device to device
        ComponentData &operator=(const ComponentData &other) {
            if (other.owner_->dimension_ != owner_->dimension_) throw std::invalid_argument("Component sizes mismatch.");
            const size_t n = owner_->dimension_;
            b200::check(hehub_b200_slab_d2d(b200::context(), owner_->dev_mut() + index_ * n, other.owner_->dev() + other.index_ * n, n));
            return *this;
        }
        bool operator==(const ComponentData &other) const {
            return size() == other.size() && std::memcmp(data(), other.data(), size() * sizeof(u64)) == 0;
        }

    private:
        friend class RnsIntVec;
        ComponentData(RnsIntVec *owner, size_t index) : owner_(owner), index_(index) {}
        RnsIntVec *owner_;
        size_t index_;
    };

    enum class RepForm { coeff, value };

    RnsIntVec() {}

    RnsIntVec(const size_t dimension, const size_t components, const std::vector<u64> &moduli)
        : dimension_(dimension) {
        // rns.cpp:10-28: power-of-two dimension, at least `components` moduli (extra ones are ignored)
        if (dimension == 0 || (dimension & (dimension - 1)) != 0) throw std::invalid_argument("dimension should be a 2-power.");
        if (moduli.size() < components) throw std::invalid_argument("No matching number of moduli provided to create RnsIntVec.");
        moduli_.assign(moduli.begin(), moduli.begin() + components);
        while (((size_t)1 << log_dimension_) < dimension) log_dimension_++;
        allocate(components);
    }

    RnsIntVec(const Params &params) : RnsIntVec(params.dimension, params.component_count, params.moduli) {}

    RnsIntVec(const RnsIntVec &other) { copy_from(other); }
    RnsIntVec(RnsIntVec &&other) noexcept { steal(other); }
    RnsIntVec &operator=(const RnsIntVec &other) {
        if (this != &other) {
            release();
            copy_from(other);
        }
        return *this;
    }
    RnsIntVec &operator=(RnsIntVec &&other) noexcept {
        if (this != &other) {
            release();
            steal(other);
        }
        return *this;
    }
    ~RnsIntVec() { release(); }

    /// raw-word equality, like rns.h:44-48 (rep_form is not compared)
    const bool operator==(const RnsIntVec &other) const {
        if (dimension_ != other.dimension_ || moduli_ != other.moduli_) return false;
        const size_t words = dimension_ * moduli_.size();
        if (words == 0) return true;
        return std::memcmp(host_words(0, false), other.host_words(0, false), words * sizeof(u64)) == 0;
    }

    Params params() const { return Params{dimension_, moduli_.size(), moduli_}; }
    const size_t component_count() const { return moduli_.size(); }
    const size_t log_dimension() const { return log_dimension_; }
    const size_t dimension() const { return dimension_; }
    std::vector<ComponentData> &components() { return views_; }
    const std::vector<ComponentData> &components() const { return views_; }
    auto begin() { return views_.begin(); }
    const auto begin() const { return views_.cbegin(); }
    auto end() { return views_.end(); }
    const auto end() const { return views_.cend(); }
    auto last() { return views_.end() - 1; }
    const auto last() const { return views_.cend() - 1; }
    const u64 modulus_at(int i) const { return moduli_[i]; }
    const std::vector<u64> &modulus_vec() const { return moduli_; }
    ComponentData &operator[](int i) { return views_[i]; }
    const ComponentData &operator[](int i) const { return views_[i]; }

    /// rns.cpp:33-46: appends `adding` uninitialised components.  Fenced quirk: the
    /// reference appends ALL supplied moduli but only `adding` buffers; here only the first `adding`
    /// moduli are taken, so the container stays consistent.
    void add_components(const std::vector<u64> &new_moduli, size_t adding = 1) {
        if (new_moduli.size() < adding) throw std::invalid_argument("No matching number of moduli provided to add components.");
        const size_t old = moduli_.size(), words_old = old * dimension_;
        RnsIntVec grown;
        grown.dimension_ = dimension_;
        grown.log_dimension_ = log_dimension_;
        grown.moduli_ = moduli_;
        grown.moduli_.insert(grown.moduli_.end(), new_moduli.begin(), new_moduli.begin() + adding);
        grown.allocate(old + adding);
        if (words_old) b200::check(hehub_b200_slab_d2d(b200::context(), grown.dev_, dev(), words_old));
        *this = std::move(grown);
    }

    /// rns.cpp:48-56
    void remove_components(size_t removing = 1) {
        if (removing > moduli_.size()) throw std::invalid_argument("Trying to remove components more than existing.");
        dev(); // make the device copy current; the slab keeps its capacity
        host_valid_ = false;
        host_.clear();
        moduli_.resize(moduli_.size() - removing);
        rebuild_views();
    }

    // ---- device access for the operators of this library ---------------------------------
    /// device words [limb][N], current; the host mirror stays valid
    const u64 *dev() const {
        if (!dev_valid_) {
            b200::check(hehub_b200_slab_h2d(b200::context(), dev_, host_.data(), host_.size()));
            b200::synchronize();
            dev_valid_ = true;
        }
        return dev_;
    }
    /// device words for writing: the host mirror is dropped
    u64 *dev_mut() {
        dev();
        host_valid_ = false;
        return dev_;
    }

protected:
    size_t log_dimension_ = 0;
    size_t dimension_ = 0;
    std::vector<u64> moduli_;

public:
    /// A polynomial over `params` whose words are the `params.component_count * dimension` words at `at` inside
    /// `block` (filled by a kernel): no copy.  Polynomials adopted from one block are contiguous operands.
    static RnsIntVec adopt(std::shared_ptr<detail::SlabBlock> block, u64 *at, const Params &params) {
        RnsIntVec v;
        v.dimension_ = params.dimension;
        while (((size_t)1 << v.log_dimension_) < v.dimension_) v.log_dimension_++;
        v.moduli_.assign(params.moduli.begin(), params.moduli.begin() + params.component_count);
        v.block_ = std::move(block);
        v.dev_ = at;
        v.capacity_words_ = params.component_count * params.dimension;
        v.dev_valid_ = true;
        v.host_valid_ = false;
        v.rebuild_views();
        return v;
    }
    /// true when `next` starts where this polynomial's words end, in the same block, and both device copies are current
    bool device_adjacent(const RnsIntVec &next) const {
        return block_ && block_ == next.block_ && dev_valid_ && next.dev_valid_ && next.dev_ == dev_ + moduli_.size() * dimension_;
    }

private:
    std::shared_ptr<detail::SlabBlock> block_; // owns the slab dev_ points into (possibly shared with sibling polynomials)
    u64 *dev_ = nullptr;
    size_t capacity_words_ = 0;
    mutable std::vector<u64> host_;
    mutable bool dev_valid_ = true;
    mutable bool host_valid_ = false;
    std::vector<ComponentData> views_;

    void rebuild_views() {
        views_.clear();
        for (size_t k = 0; k < moduli_.size(); k++) views_.push_back(ComponentData(this, k));
    }
    void allocate(size_t components) {
        capacity_words_ = components * dimension_;
        block_ = std::make_shared<detail::SlabBlock>(capacity_words_);
        dev_ = block_->p;
        dev_valid_ = true;
        host_valid_ = false;
        rebuild_views();
    }
    void release() {
        block_.reset(); // the slab goes back to the per-size pool with its last owner
        dev_ = nullptr;
        capacity_words_ = 0;
        host_.clear();
        views_.clear();
    }
    void copy_from(const RnsIntVec &o) {
        dimension_ = o.dimension_;
        log_dimension_ = o.log_dimension_;
        moduli_ = o.moduli_;
        dev_ = nullptr;
        allocate(moduli_.size());
        const size_t words = moduli_.size() * dimension_;
        if (words) {
            if (o.dev_valid_) {
                b200::check(hehub_b200_slab_d2d(b200::context(), dev_, o.dev_, words));
            } else {
                host_ = o.host_;
                host_valid_ = true;
                dev_valid_ = false;
            }
        }
    }
    void steal(RnsIntVec &o) {
        dimension_ = o.dimension_;
        log_dimension_ = o.log_dimension_;
        moduli_ = std::move(o.moduli_);
        block_ = std::move(o.block_);
        dev_ = o.dev_;
        capacity_words_ = o.capacity_words_;
        host_ = std::move(o.host_);
        dev_valid_ = o.dev_valid_;
        host_valid_ = o.host_valid_;
        o.dev_ = nullptr;
        o.capacity_words_ = 0;
        o.dimension_ = o.log_dimension_ = 0;
        o.moduli_.clear();
        o.views_.clear();
        rebuild_views();
    }
    u64 *host_words(size_t component, bool for_write) const {
        const size_t words = moduli_.size() * dimension_;
        if (!host_valid_) {
            host_.resize(words);
            b200::check(hehub_b200_slab_d2h(b200::context(), host_.data(), dev_, words));
            b200::synchronize();
            host_valid_ = true;
        }
        if (for_write) dev_valid_ = false;
        return host_.data() + component * dimension_;
    }
};

class RnsPolynomial : public RnsIntVec {
public:
    using RnsIntVec::RnsIntVec;
    enum class RepForm { coeff, value };
    RnsPolynomial() {}
    RnsPolynomial(RnsIntVec &&rns_int_vec) : RnsIntVec(std::move(rns_int_vec)) {}
    RnsPolynomial(const RnsIntVec &rns_int_vec) : RnsIntVec(rns_int_vec) {}
    RepForm rep_form = RepForm::coeff;
};

using RnsPolyParams = RnsPolynomial::Params;
using PolyRepForm = RnsPolynomial::RepForm;

namespace detail {
// rns.cpp:58-72 / 88-103: `b` may carry MORE components than self (a rescaled ciphertext against a full-length key);
// the first self.component_count() moduli must agree and only those limbs take part.  The [limb][N] layout makes a
// prefix of b a valid operand.
inline void check_accumulate(const RnsIntVec &self, const RnsIntVec &b) {
    if (self.dimension() != b.dimension()) throw std::invalid_argument("Operands' poly len mismatch.");
    if (b.component_count() < self.component_count()) throw std::invalid_argument("Operand b contains less components than self.");
    for (size_t k = 0; k < self.component_count(); k++)
        if (self.modulus_at((int)k) != b.modulus_at((int)k)) throw std::invalid_argument("Operands' moduli mismatch.");
}
// fused ciphertext-by-ciphertext kernels take operands of one shape
inline void check_same_shape(const RnsIntVec &a, const RnsIntVec &b) {
    if (a.dimension() != b.dimension()) throw std::invalid_argument("Operands' poly len mismatch.");
    if (a.component_count() != b.component_count()) throw std::invalid_argument("Operands' component numbers mismatch.");
    if (a.modulus_vec() != b.modulus_vec()) throw std::invalid_argument("Operands' moduli mismatch.");
}
// rns.cpp:120-131: the product has min(a, b) components; the truncated modulus lists must agree
inline size_t check_product(const RnsIntVec &a, const RnsIntVec &b) {
    if (a.dimension() != b.dimension()) throw std::invalid_argument("Operands' poly len mismatch.");
    const size_t components = a.component_count() < b.component_count() ? a.component_count() : b.component_count();
    for (size_t k = 0; k < components; k++)
        if (a.modulus_at((int)k) != b.modulus_at((int)k)) throw std::invalid_argument("Operands' moduli mismatch.");
    return components;
}
} // namespace detail

// ---- RnsIntVec operators: rns.cpp:58-171 ---------------------------------------------------
inline const RnsIntVec &operator+=(RnsIntVec &self, const RnsIntVec &b) {
    detail::check_accumulate(self, b);
    if (self.component_count() == 0) return self;
    b200::check(hehub_b200_add_lazy(b200::context(), self.dimension(), self.modulus_vec().data(), self.component_count(), self.dev_mut(), b.dev(), 1));
    return self;
}
inline RnsIntVec operator+(const RnsIntVec &a, const RnsIntVec &b) {
    auto result(a);
    result += b;
    return result;
}
inline const RnsIntVec &operator-=(RnsIntVec &self, const RnsIntVec &b) {
    detail::check_accumulate(self, b);
    if (self.component_count() == 0) return self;
    b200::check(hehub_b200_sub_lazy(b200::context(), self.dimension(), self.modulus_vec().data(), self.component_count(), self.dev_mut(), b.dev(), 1));
    return self;
}
inline RnsIntVec operator-(const RnsIntVec &a, const RnsIntVec &b) {
    auto result(a);
    result -= b;
    return result;
}
inline RnsIntVec operator*(const RnsIntVec &a, const RnsIntVec &b) {
    const size_t components = detail::check_product(a, b);
    RnsIntVec result(a.dimension(), components, a.modulus_vec());
    if (components)
        b200::check(hehub_b200_mulmod_hybrid_lazy(b200::context(), a.dimension(), result.modulus_vec().data(), components, a.dev(), b.dev(), result.dev_mut(), 1));
    return result;
}
inline const RnsIntVec &operator*=(RnsIntVec &self, const RnsIntVec &b) {
    if (b.component_count() >= self.component_count()) { // in place: same limbs, same moduli prefix
        detail::check_product(self, b);
        if (self.component_count())
            b200::check(hehub_b200_mulmod_hybrid_lazy(b200::context(), self.dimension(), self.modulus_vec().data(), self.component_count(), self.dev(), b.dev(), self.dev_mut(), 1));
    } else {
        self = self * b; // the product is shorter than self (rns.cpp:120-140 through rns.h's operator*=)
    }
    return self;
}
inline const RnsIntVec &operator*=(RnsIntVec &self, const std::vector<u64> &rns_scalar) {
    if (rns_scalar.size() != self.component_count()) throw std::invalid_argument("Numbers of RNS components mismatch."); // rns.cpp:158
    b200::check(hehub_b200_mul_scalar_lazy(b200::context(), self.dimension(), self.modulus_vec().data(), self.component_count(), self.dev_mut(), rns_scalar.data(), 1));
    return self;
}
inline const RnsIntVec &operator*=(RnsIntVec &self, const u64 small_scalar) {
    return self *= std::vector<u64>(self.component_count(), small_scalar); // rns.cpp:142-153: scalar mod q_i per limb
}
inline RnsIntVec operator*(const RnsIntVec &v, const u64 small_scalar) {
    auto copy(v);
    copy *= small_scalar;
    return copy;
}
inline RnsIntVec operator*(const RnsIntVec &v, const std::vector<u64> &rns_scalar) {
    auto copy(v);
    copy *= rns_scalar;
    return copy;
}

// ---- RnsPolynomial operators: rns.h:207-281 (representation-form checks) ---------------------
inline const RnsPolynomial &operator+=(RnsPolynomial &self, const RnsPolynomial &b) {
    if (self.rep_form != b.rep_form) throw std::invalid_argument("Operands are in different representation form.");
    static_cast<RnsIntVec &>(self) += static_cast<const RnsIntVec &>(b);
    return self;
}
inline RnsPolynomial operator+(const RnsPolynomial &a, const RnsPolynomial &b) {
    auto result(a);
    result += b;
    return result;
}
inline const RnsPolynomial &operator-=(RnsPolynomial &self, const RnsPolynomial &b) {
    if (self.rep_form != b.rep_form) throw std::invalid_argument("Operands are in different representation form.");
    static_cast<RnsIntVec &>(self) -= static_cast<const RnsIntVec &>(b);
    return self;
}
inline RnsPolynomial operator-(const RnsPolynomial &a, const RnsPolynomial &b) {
    auto result(a);
    result -= b;
    return result;
}
inline RnsPolynomial operator*(const RnsPolynomial &a, const RnsPolynomial &b) {
    if (a.rep_form != PolyRepForm::value || b.rep_form != PolyRepForm::value)
        throw std::invalid_argument("Polynomial multiplication requires NTT form (value representation)."); // rns.h:242-247
    RnsPolynomial result(static_cast<const RnsIntVec &>(a) * static_cast<const RnsIntVec &>(b));
    result.rep_form = PolyRepForm::value;
    return result;
}
inline const RnsPolynomial &operator*=(RnsPolynomial &self, const RnsPolynomial &b) {
    self = self * b;
    return self;
}
inline const RnsPolynomial &operator*=(RnsPolynomial &self, const u64 small_scalar) {
    static_cast<RnsIntVec &>(self) *= small_scalar;
    return self;
}
inline const RnsPolynomial &operator*=(RnsPolynomial &self, const std::vector<u64> &rns_scalar) {
    static_cast<RnsIntVec &>(self) *= rns_scalar;
    return self;
}
inline RnsPolynomial operator*(const RnsPolynomial &poly, const std::vector<u64> &rns_scalar) {
    auto copy(poly);
    copy *= rns_scalar;
    return copy;
}
inline RnsPolynomial operator*(const RnsPolynomial &poly, const u64 small_scalar) {
    auto copy(poly);
    copy *= small_scalar;
    return copy;
}

} // namespace hehub

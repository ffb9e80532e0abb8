// serialize.h — raw-slab dump / load of RnsPolynomial, ciphertexts and key-switch keys (SURVEY §8(f)
// rank 4).  The reference has no wire or disk format; this is the `[header][moduli][limb slabs]`
// layout documented in hehub_b200/slabio.py (same bytes), so fixtures travel between machines and
// between this mirror and the Python harness.
#pragma once
#include <array>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rgsw.h"
#include "rlwe.h"
#include "rns.h"

namespace hehub {
namespace b200 {

enum SlabKind : uint32_t { kPolynomial = 1, kCiphertext = 2, kKeySwitchKey = 3 };

namespace detail_io {
struct Header {
    char magic[8];
    uint32_t version, kind;
    uint64_t n, limbs, polys;
    uint32_t rep_form, reserved;
};
static_assert(sizeof(Header) == 48, "packed little-endian header");

inline void write_all(const std::string &path, uint32_t kind, const std::vector<const RnsPolynomial *> &polys) {
    if (polys.empty()) throw std::invalid_argument("nothing to write");
    const RnsPolynomial &p0 = *polys[0];
    Header h{};
    std::memcpy(h.magic, "HEHB200\0", 8);
    h.version = 1;
    h.kind = kind;
    h.n = p0.dimension();
    h.limbs = p0.component_count();
    h.polys = polys.size();
    h.rep_form = p0.rep_form == PolyRepForm::value ? 1u : 0u;
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && std::fwrite(p0.modulus_vec().data(), 8, h.limbs, f) == h.limbs;
    for (const RnsPolynomial *p : polys) {
        if (p->dimension() != h.n || p->modulus_vec() != p0.modulus_vec() || p->rep_form != p0.rep_form) {
            std::fclose(f);
            throw std::invalid_argument("polynomials of one file must share shape, moduli and representation");
        }
        for (size_t k = 0; k < h.limbs && ok; k++) ok = std::fwrite((*p)[k].data(), 8, h.n, f) == h.n;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("short write to " + path);
}

inline std::vector<RnsPolynomial> read_all(const std::string &path, uint32_t kind) {
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    Header h{};
    std::vector<RnsPolynomial> out;
    try {
        if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "HEHB200\0", 8) != 0 || h.version != 1)
            throw std::invalid_argument("not a hehub_b200 slab file: " + path);
        if (h.kind != kind) throw std::invalid_argument("slab file holds a different kind of object");
        if (h.limbs == 0 || h.limbs > 1024 || h.n == 0 || (h.n & (h.n - 1)) || h.n > (1u << 16) || h.polys == 0 || h.polys > (1u << 20))
            throw std::invalid_argument("implausible slab header");
        std::vector<u64> moduli(h.limbs);
        if (std::fread(moduli.data(), 8, h.limbs, f) != h.limbs) throw std::invalid_argument("truncated slab file");
        for (uint64_t i = 0; i < h.polys; i++) {
            RnsPolynomial p(h.n, h.limbs, moduli);
            for (size_t k = 0; k < h.limbs; k++)
                if (std::fread(p[k].data(), 8, h.n, f) != h.n) throw std::invalid_argument("truncated slab file");
            p.rep_form = h.rep_form ? PolyRepForm::value : PolyRepForm::coeff;
            out.push_back(std::move(p));
        }
        if (std::fgetc(f) != EOF) throw std::invalid_argument("trailing bytes in slab file");
    } catch (...) {
        std::fclose(f);
        throw;
    }
    std::fclose(f);
    return out;
}
} // namespace detail_io

inline void save(const std::string &path, const RnsPolynomial &poly) { detail_io::write_all(path, kPolynomial, {&poly}); }
template <size_t K>
inline void save(const std::string &path, const std::array<RnsPolynomial, K> &ct) {
    std::vector<const RnsPolynomial *> v;
    for (const auto &p : ct) v.push_back(&p);
    detail_io::write_all(path, kCiphertext, v);
}
inline void save(const std::string &path, const RlweKsk &key) {
    std::vector<const RnsPolynomial *> v;
    for (const auto &row : key)
        for (const auto &p : row) v.push_back(&p);
    detail_io::write_all(path, kKeySwitchKey, v);
}

inline RnsPolynomial load_polynomial(const std::string &path) { return std::move(detail_io::read_all(path, kPolynomial).at(0)); }
inline RlweCt load_ciphertext(const std::string &path) {
    auto v = detail_io::read_all(path, kCiphertext);
    if (v.size() != 2) throw std::invalid_argument("not a degree-1 ciphertext");
    return RlweCt{std::move(v[0]), std::move(v[1])};
}
inline RlweKsk load_key_switch_key(const std::string &path) {
    auto v = detail_io::read_all(path, kKeySwitchKey);
    if (v.size() % 2) throw std::invalid_argument("odd number of polynomials in a key file");
    RgswCt rows;
    for (size_t r = 0; r + 1 < v.size(); r += 2) rows.push_back(RlweCt{std::move(v[r]), std::move(v[r + 1])});
    return RlweKsk(std::move(rows));
}

} // namespace b200
} // namespace hehub

// bitset.cuh — move sets as bit masks: 64 bits for card games (52 cards), 256 bits
// for boards up to 16x16.  Replaces the std::vector<Move> the reference passes
// around (game.h:27, node.h:58-91).
#pragma once

#include "devdefs.cuh"

namespace ismcts {

struct Mask64 {
    uint64_t w;
    ISMCTS_DEV static Mask64 none() { return Mask64{0}; }
    ISMCTS_DEV Mask64 with(uint32_t b) const { return Mask64{w | ((uint64_t)1 << b)}; }
    ISMCTS_DEV bool any() const { return w != 0; }
    ISMCTS_DEV uint32_t count() const { return popc64(w); }
    ISMCTS_DEV uint32_t kth(uint32_t k) const { return kth_bit64(w, k); }
    ISMCTS_DEV uint32_t pop_lowest() { const uint32_t b = ctz64(w); w &= w - 1; return b; }
    ISMCTS_DEV bool test(uint32_t b) const { return (w >> b) & 1u; }
    ISMCTS_DEV void set(uint32_t b) { w |= (uint64_t)1 << b; }
    ISMCTS_DEV Mask64 minus(Mask64 o) const { return Mask64{w & ~o.w}; }
    // shared trees: the mask lives in a node that other lanes update concurrently; setting a bit PUBLISHES the child
    // (a release: what the lane wrote to the child before is visible to whoever sees the bit)
    ISMCTS_DEV static Mask64 load_shared(const Mask64* p) { return Mask64{load_u64_cg(&p->w)}; }
    ISMCTS_DEV static void set_shared(Mask64* p, uint32_t b) { atomic_or_u64_release(&p->w, (uint64_t)1 << b); }
    ISMCTS_DEV static bool test_shared_acquire(const Mask64* p, uint32_t b) { return (load_u64_acquire(&p->w) >> b) & 1u; }
};

struct Mask256 {
    uint64_t w[4];
    ISMCTS_DEV static Mask256 none() { return Mask256{{0, 0, 0, 0}}; }
    ISMCTS_DEV Mask256 with(uint32_t b) const { Mask256 r = *this; r.set(b); return r; }
    ISMCTS_DEV bool any() const { return (w[0] | w[1] | w[2] | w[3]) != 0; }
    ISMCTS_DEV uint32_t count() const { return popc64(w[0]) + popc64(w[1]) + popc64(w[2]) + popc64(w[3]); }
    ISMCTS_DEV uint32_t kth(uint32_t k) const
    {
        uint32_t base = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint32_t c = popc64(w[i]);
            if (k < c) return base + kth_bit64(w[i], k);
            k -= c; base += 64;
        }
        return base + kth_bit64(w[3], k);
    }
    ISMCTS_DEV uint32_t pop_lowest()
    {
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (w[i]) { const uint32_t b = ctz64(w[i]); w[i] &= w[i] - 1; return 64u * i + b; }
        const uint32_t b = ctz64(w[3]); w[3] &= w[3] - 1; return 192u + b;
    }
    // word selection is written as unrolled selects so that the mask stays in registers
    ISMCTS_DEV bool test(uint32_t b) const
    {
        const uint32_t q = b >> 6;
        const uint64_t x = q == 0 ? w[0] : q == 1 ? w[1] : q == 2 ? w[2] : w[3];
        return (x >> (b & 63u)) & 1u;
    }
    ISMCTS_DEV void set(uint32_t b)
    {
        const uint64_t bit = (uint64_t)1 << (b & 63u);
        const uint32_t q = b >> 6;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) w[i] |= (q == i) ? bit : 0;
    }
    ISMCTS_DEV static Mask256 load_shared(const Mask256* p)
    {
        return Mask256{{load_u64_cg(&p->w[0]), load_u64_cg(&p->w[1]), load_u64_cg(&p->w[2]), load_u64_cg(&p->w[3])}};
    }
    ISMCTS_DEV static void set_shared(Mask256* p, uint32_t b) { atomic_or_u64_release(&p->w[b >> 6], (uint64_t)1 << (b & 63u)); }
    ISMCTS_DEV static bool test_shared_acquire(const Mask256* p, uint32_t b) { return (load_u64_acquire(&p->w[b >> 6]) >> (b & 63u)) & 1u; }
    ISMCTS_DEV Mask256 minus(const Mask256& o) const
    {
        return Mask256{{w[0] & ~o.w[0], w[1] & ~o.w[1], w[2] & ~o.w[2], w[3] & ~o.w[3]}};
    }
};

} // namespace ismcts

// philox.cuh — counter-based Philox4x32-10 stream of one search iteration.
//
// Replaces the reference's thread_local std::mt19937 (utility.h:62-67 and
// knockoutwhist.cpp:28-32).  Word d of an iteration is
//   Philox4x32-10(counter = (d/4, iteration, global tree id, position id),
//                 key = seed)[d % 4],
// so any random choice can be recomputed from its coordinates alone and the
// result of a search does not depend on how trees map to threads or GPUs.
#pragma once

#include "devdefs.cuh"

namespace ismcts {

struct Philox {
    uint32_t k0, k1;         // key = seed
    uint32_t c1, c2, c3;     // iteration, tree, position
    uint32_t d;              // next word index
    uint32_t r0, r1, r2, r3; // words of the current block not yet handed out (r0 first)

    ISMCTS_DEV void init(uint64_t seed, uint32_t pos, uint32_t tree)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c1 = 0; c2 = tree; c3 = pos; d = 0;
        r0 = r1 = r2 = r3 = 0;
    }

    ISMCTS_DEV void begin_iteration(uint32_t iteration) { c1 = iteration; d = 0; }

    // Positions the stream at word `nd` of the current iteration (host helpers resume
    // a stream in the middle of a block; the search only starts at boundaries).
    ISMCTS_DEV void seek(uint32_t nd)
    {
        d = nd & ~3u;
        if (nd & 3u) {
            refill();
            for (uint32_t i = 0; i < (nd & 3u); ++i) { r0 = r1; r1 = r2; r2 = r3; }
        }
        d = nd;
    }

    // Skips to the next block boundary (rollouts start on one so that all lanes
    // of a warp refill on the same step).
    ISMCTS_DEV void align() { d = (d + 3u) & ~3u; }

    ISMCTS_DEV void refill()
    {
        uint32_t x0 = d >> 2, x1 = c1, x2 = c2, x3 = c3;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int round = 0; round < 10; ++round) {
            const uint32_t hi0 = mulhi32(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = mulhi32(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            x0 = hi1 ^ x1 ^ a;
            x1 = lo1;
            x2 = hi0 ^ x3 ^ b;
            x3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        r0 = x0; r1 = x1; r2 = x2; r3 = x3;
    }

    // Next 32-bit word of the stream.  Words are always consumed in order from a
    // block boundary (d starts at 0 and align() only moves to boundaries).
    ISMCTS_DEV uint32_t word()
    {
        if ((d & 3u) == 0u) refill();
        const uint32_t w = r0;
        r0 = r1; r1 = r2; r2 = r3;
        ++d;
        return w;
    }

    // Uniform integer in [0, n) — always consumes exactly one word.
    ISMCTS_DEV uint32_t below(uint32_t n) { return mulhi32(word(), n); }
};

} // namespace ismcts

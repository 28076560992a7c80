// mnk_game.cuh — the m,n,k-game as a device-side POD state machine.
//
// Device mirror of ISMCTS::Game<int> for the reference's MnkGame
// (test/common/mnkgame.{h,cpp}): n rows by m columns,
// move = cell index, a win is a run of EXACTLY k through the last stone
// (mnkgame.cpp:109-121).  Perfect information, no chance events.  The two stone
// sets (256-bit masks) live in shared memory, one column per thread.
#pragma once

#include "devdefs.cuh"
#include "bitset.cuh"
#include "../../include/ismcts_b200.h"

namespace ismcts {

struct MnkGame {
    static constexpr uint32_t kGameId = ISMCTS_GAME_MNK;
    static constexpr uint32_t kMoveStride = 256;
    static constexpr uint32_t kHandWords = 8; // stones[2][4]
    static constexpr uint32_t kMaxPlayers = 2;
    static constexpr uint32_t kMaxLegal = ISMCTS_MNK_MAX_CELLS;
    static constexpr uint32_t kMaxDepth = 256;
    using Mask = Mask256;
    using Root = ismcts_mnk_state;
    using Prepared = ismcts_mnk_state;   // perfect information: an iteration starts from the root itself

    uint64_t* stones; // stones[(p * 4 + word) * stride]
    uint32_t stride;
    uint32_t m, n, k, player;
    int32_t res2;     // 2 * m_result from player 0's view, -1 while running

    // `scratch`: the block's scratch, rows of `scratch_stride` 8-byte words; this game owns column `column`
    ISMCTS_DEV void bind(uint64_t* scratch, uint32_t scratch_stride, uint32_t column) { stones = scratch + column; stride = scratch_stride; }
    ISMCTS_DEV static void init_block(uint32_t, uint32_t) {}   // no per-block tables
    // the same columns again (a copy of the game whose pointers the compiler can follow)
    ISMCTS_DEV void rebind(uint64_t* scratch, uint32_t scratch_stride, uint32_t column) { stones = scratch + column; stride = scratch_stride; }

    ISMCTS_DEV void load(const Root* r)
    {
        for (uint32_t i = 0; i < 8; ++i) stones[i * stride] = r->stones[i >> 2][i & 3];
        m = r->m; n = r->n; k = r->k; player = r->player; res2 = r->result2;
    }

    ISMCTS_DEV uint32_t current_player() const { return player; }
    ISMCTS_DEV static uint32_t root_player(const Prepared* p) { return p->player; }
    ISMCTS_DEV static uint32_t root_players(const Prepared*) { return 3u; } // players(), mnkgame.cpp:76-79
    ISMCTS_DEV bool chance_pending() const { return false; }    // perfect information, no chance events
    // cloneAndRandomise is a plain copy, mnkgame.cpp:34-37
    ISMCTS_DEV static void prepare(const Root* r, Prepared* out, uint64_t*) { *out = *r; }
    ISMCTS_DEV void start(const Prepared* p) { load(p); }
    ISMCTS_DEV Mask chance_set() const { return Mask::none(); }
    ISMCTS_DEV void chance_step(uint32_t) {}
    ISMCTS_DEV uint32_t chance_count() const { return 0; }
    ISMCTS_DEV void chance_draw(uint32_t) {}
    ISMCTS_DEV bool terminal() const { return res2 != -1; }     // validMoves() is empty, mnkgame.cpp:61-64
    ISMCTS_DEV void play(uint32_t cell) { step(cell); }

    ISMCTS_DEV bool occupied_by(uint32_t p, uint32_t cell) const
    {
        return (stones[(p * 4 + (cell >> 6)) * stride] >> (cell & 63u)) & 1u;
    }

    // validMoves, mnkgame.cpp:61-64
    ISMCTS_DEV Mask legal() const
    {
        Mask out;
        const uint32_t cells = m * n;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t lo = 64u * i;
            const uint64_t board = cells >= lo + 64u ? ~(uint64_t)0 : cells > lo ? (((uint64_t)1 << (cells - lo)) - 1u) : 0;
            out.w[i] = res2 == -1 ? board & ~(stones[i * stride] | stones[(4 + i) * stride]) : 0;
        }
        return out;
    }

    // hasWinningSequence, mnkgame.cpp:109-130: run length through `cell` along (dr, dc)
    ISMCTS_DEV bool exact_run(uint32_t cell, int dr, int dc) const
    {
        uint32_t count = 1;
        const int r0 = (int)(cell / m), c0 = (int)(cell % m);
        for (int sign = -1; sign <= 1; sign += 2) {
            int r = r0 + sign * dr, c = c0 + sign * dc;
            while (r >= 0 && c >= 0 && r < (int)n && c < (int)m && occupied_by(player, (uint32_t)(r * (int)m + c))) {
                ++count; r += sign * dr; c += sign * dc;
            }
        }
        return count == k;
    }

    // doMove, mnkgame.cpp:39-54
    ISMCTS_DEV void step(uint32_t cell)
    {
        uint64_t* word = &stones[(player * 4 + (cell >> 6)) * stride];
        *word |= (uint64_t)1 << (cell & 63u);
        if (exact_run(cell, 0, 1) || exact_run(cell, 1, 0) || exact_run(cell, 1, 1) || exact_run(cell, -1, 1)) {
            res2 = 2 * (1 - (int32_t)player);
            return;
        }
        uint32_t filled = 0;
        for (uint32_t i = 0; i < 8; ++i) filled += popc64(stones[i * stride]);
        if (filled < m * n) player = 1u - player; else res2 = 1;
    }

    // getResult * 2, mnkgame.cpp:56-59
    ISMCTS_DEV uint32_t result2(uint32_t who) const { return (uint32_t)(who == 0 ? res2 : 2 - res2); }
    ISMCTS_DEV uint32_t outcome() const { return (uint32_t)res2 & 0xFFu; }
};

} // namespace ismcts

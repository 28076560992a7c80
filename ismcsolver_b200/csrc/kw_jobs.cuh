// kw_jobs.cuh — the headline search (Knockout Whist, SO-ISMCTS, UCB1, one tree per job) with the game
// states in shared memory and several jobs per lane.
//
// lane_search.cuh keeps one tree per lane with its state in registers; the lanes of a warp are in
// different phases by construction, so every pass of its loop executes the union of the code paths its
// lanes need (chance step + move step + rare paths), each with about half of the lanes active
// (profiles/r01_bc_summary.md).  Here a lane owns K jobs (trees) whose states live in shared memory, and a
// pass performs ONE kind of micro-step for the whole warp — the kind most lanes can serve — on one job
// per lane: a lane whose first job waits for the other kind steps its second job.  A pass therefore runs
// one short code path with most lanes active (tests/emu emu_multijob_sim measured 20 of 32 lanes for
// K = 2, 24 for K = 4, against 13 + 13 per double-length pass before).  The price is that a step loads and
// stores the fields it touches instead of finding them in registers.
//
// Semantics are unchanged: every job performs exactly the sequential sequence of operations of
// SOSolver::search (sosolver.h:44-71) with the Philox streams of philox.cuh, so results are bit-identical
// with lane_search.cuh and the oracle whatever the schedule.
#pragma once

#include "devdefs.cuh"
#include "philox.cuh"
#include "pool.cuh"
#include "kw_game.cuh"
#include "lane_search.cuh"

namespace ismcts {

// What a job waits for.
enum JobWant : uint32_t { kWantChance = 0, kWantMove = 1, kWantTree = 2, kWantNothing = 3 };

// K jobs per lane, NP players' hands stored per job.
template <uint32_t K, uint32_t NP>
struct KwJobs {
    using Mask = KwGame::Mask;
    using Slot = SlotUcb<Mask>;
    static constexpr uint32_t kRing = 8;
    // 32-bit words of shared memory per job: A B alive4 tricks need pool(2) R done node + ring + hands
    static constexpr uint32_t kWordsPerJob = 10 + kRing + 2 * NP;

    // ---- packed fields
    // A: player[0:3) trick_len[3:6) lead_suit[6:8) win_player[8:11) trump[11:13) request_trump[13]
    //    win_key[14:20) tricks_left[20:23) dealer[23:26) best[26:29)
    // B: mode[0:2) deal_player[2:5) deal_left[5:8) want[8:10) rollout[10] depth[16:24)
    // R: count[0:4) head[4:7) next_block[8:32)

    const SearchParams& P;
    uint32_t J;          // jobs of this CTA = K * threads
    uint32_t T;          // threads of this CTA
    uint32_t first_job;  // global id (= table index = position * trees + tree) of job 0 of this CTA
    uint32_t total_jobs; // n_positions * trees
    uint32_t* path_base; // [total_jobs][KwGame::kMaxDepth] in global memory
    uint64_t deadline;
    uint32_t *A, *B, *alive4, *tricks, *need, *pool_lo, *pool_hi, *R, *done, *node, *ring;
    uint64_t* hands;     // [NP][J]

    ISMCTS_DEV explicit KwJobs(const SearchParams& p) : P(p) {}

    ISMCTS_DEV void bind(uint32_t* smem, uint32_t threads, uint32_t cta_first_job)
    {
        T = threads; J = K * threads; first_job = cta_first_job; total_jobs = P.n_positions * P.trees;
        hands = reinterpret_cast<uint64_t*>(smem);            // 8-byte aligned part first
        uint32_t* w = smem + 2 * NP * J;
        A = w; B = w + J; alive4 = w + 2 * J; tricks = w + 3 * J; need = w + 4 * J; pool_lo = w + 5 * J;
        pool_hi = w + 6 * J; R = w + 7 * J; done = w + 8 * J; node = w + 9 * J; ring = w + 10 * J;
    }

    // ---- game state <-> shared memory

    ISMCTS_DEV void load_game(uint32_t j, KwGame& g) const
    {
        g.bind(hands + j, J);
        const uint32_t a = A[j], b = B[j];
        g.player = a & 7u; g.trick_len = (a >> 3) & 7u; g.lead_suit = (a >> 6) & 3u; g.win_player = (a >> 8) & 7u;
        g.trump = (a >> 11) & 3u; g.request_trump = (a >> 13) & 1u; g.win_key = (a >> 14) & 63u;
        g.tricks_left = (a >> 20) & 7u; g.dealer = (a >> 23) & 7u; g.best = (a >> 26) & 7u;
        g.mode = b & 3u; g.deal_player = (b >> 2) & 7u; g.deal_left = (b >> 5) & 7u;
        g.alive4 = alive4[j]; g.tricks = tricks[j]; g.need = need[j];
        g.pool = (uint64_t)pool_lo[j] | ((uint64_t)pool_hi[j] << 32);
    }

    ISMCTS_DEV static uint32_t pack_a(const KwGame& g)
    {
        return g.player | (g.trick_len << 3) | (g.lead_suit << 6) | (g.win_player << 8) | (g.trump << 11) |
               (g.request_trump << 13) | (g.win_key << 14) | (g.tricks_left << 20) | (g.dealer << 23) | (g.best << 26);
    }

    // What the job waits for next (rollout = the expansion of this iteration is done).
    ISMCTS_DEV static uint32_t want_of(const KwGame& g, bool rollout)
    {
        if (g.chance_pending()) return kWantChance;
        if (g.terminal() || !rollout) return kWantTree;
        return kWantMove;
    }

    ISMCTS_DEV void store_game(uint32_t j, const KwGame& g, bool rollout, uint32_t depth)
    {
        A[j] = pack_a(g);
        B[j] = g.mode | (g.deal_player << 2) | (g.deal_left << 5) | (want_of(g, rollout) << 8) | ((rollout ? 1u : 0u) << 10) |
               (depth << 16);
        alive4[j] = g.alive4; tricks[j] = g.tricks; need[j] = g.need;
        pool_lo[j] = (uint32_t)g.pool; pool_hi[j] = (uint32_t)(g.pool >> 32);
    }

    // ---- random words: a ring of 8 per job, refilled by refill_pass

    ISMCTS_DEV uint32_t words_waiting(uint32_t j) const { return R[j] & 15u; }

    ISMCTS_DEV uint32_t take_word(uint32_t j)
    {
        const uint32_t r = R[j];
        const uint32_t head = (r >> 4) & 7u;
        const uint32_t w = ring[head * J + j];
        R[j] = (r & ~0x7Fu) | ((r & 15u) - 1u) | (((head + 1u) & 7u) << 4);
        return w;
    }

    ISMCTS_DEV void philox_block(uint32_t j, uint32_t block, uint32_t out[4]) const
    {
        const uint32_t gid = first_job + j;
        philox4x32_10(block, done[j], P.tree_base + gid % P.trees, P.pos_base + gid / P.trees, (uint32_t)P.seed,
                      (uint32_t)(P.seed >> 32), out);
    }

    ISMCTS_DEV void refill(uint32_t j)
    {
        const uint32_t r = R[j];
        uint32_t w[4];
        philox_block(j, r >> 8, w);
        const uint32_t at = ((r >> 4) & 7u) + (r & 15u);
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) ring[((at + i) & 7u) * J + j] = w[i];
        R[j] = r + 4u + (1u << 8);
    }

    // ---- tree operations (lane_search.cuh: descend / finish_iteration, on the job's own tree)

    ISMCTS_DEV TreePool<Slot> tree_of(uint32_t j) const
    {
        TreePool<Slot> t;
        t.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)(first_job + j) << P.log2cap), P.log2cap, 1);
        return t;
    }
    ISMCTS_DEV uint32_t* path_of(uint32_t j) const { return path_base + (size_t)(first_job + j) * KwGame::kMaxDepth; }

    ISMCTS_DEV uint32_t quota_of(uint32_t j) const
    {
        const uint32_t gid = first_job + j;
        return (uint32_t)tree_iterations(P.iterations, P.trees_total, P.tree_base + gid % P.trees);
    }

    // The start of iteration done[j]: cloneAndRandomise (sosolver.h:46) as a pending deal.
    ISMCTS_DEV void start_iteration(uint32_t j)
    {
        const uint32_t gid = first_job + j;
        const KwPrepared p = reinterpret_cast<const KwPrepared*>(P.prepared)[gid / P.trees];
        R[j] = 0;
        refill(j);
        KwGame g;
        g.bind(hands + j, J);
        for (uint32_t i = 0; i < NP; ++i) g.set_hand(i, 0);
        g.set_hand(p.observer, p.hand_obs);
        g.player = p.player; g.dealer = p.dealer; g.trump = p.trump; g.tricks_left = p.tricks_left;
        g.request_trump = p.request_trump; g.alive4 = p.alive4; g.tricks = p.tricks; g.best = p.best;
        g.trick_len = p.trick_len; g.lead_suit = p.lead_suit; g.win_key = p.win_key; g.win_player = p.win_player;
        g.mode = KwGame::kPlay; g.pool = 0; g.deal_player = 0; g.deal_left = 0; g.need = 0;
        g.begin_deal(p.unseen, p.sizes);
        node[j] = 0;
        store_game(j, g, false, 0);
    }

    ISMCTS_DEV bool budget_left(uint32_t j) const
    {
        if (done[j] + 2u > P.max_nodes) return false;
        return P.iterations > 0 ? done[j] < quota_of(j) : global_timer_ns() < deadline;
    }

    // The game is over: backPropagate (solverbase.h:45-51, node.h:74-80, ucb1.h:53-56), then the next iteration.
    ISMCTS_DEV void finish_iteration(uint32_t j, const KwGame& g, uint32_t depth)
    {
        TreePool<Slot> tree = tree_of(j);
        const uint32_t* path = path_of(j);
        for (uint32_t i = 0; i < depth; ++i) {
            const uint32_t e = path[i];
            red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[e & 0xFFFFFFu]), 1u | ((uint64_t)g.result2(e >> 24) << 32));
        }
        done[j] += 1;
        if (budget_left(j)) start_iteration(j);
        else B[j] = kWantNothing << 8;
    }

    // select + expand (sosolver.h:53-71), see Lane::descend.
    ISMCTS_DEV void descend(uint32_t j, KwGame& g, uint32_t depth)
    {
        TreePool<Slot> tree = tree_of(j);
        tree.count = done[j] + 1;                                  // one node per finished iteration, and the root
        uint32_t* path = path_of(j);
        uint32_t cur = node[j];
        Mask kids = load_slot(&tree.slots[cur]).child_mask;
        bool rollout = false;
        for (;;) {
            const Mask legal = g.legal();
            if (!legal.any()) break;
            const uint32_t who = g.current_player();
            const Mask untried = legal.minus(kids);
            uint32_t move;
            if (untried.any()) {
                if (words_waiting(j) == 0) refill(j);
                move = untried.kth(mulhi32(take_word(j), untried.count()));
                tree.slots[cur].child_mask = kids.with(move);
                cur = tree.insert(cur, move, who, tree.count);
                path[depth++] = cur | (who << 24);
                g.step(move);
                rollout = true;
                break;
            }
            uint32_t best = 0, best_created = 0xFFFFFFFFu, stored = who;
            double best_u = -1.0;
            Mask best_kids = Mask::none();
            move = 0;
            for (Mask rest = legal; rest.any();) {
                const uint32_t m = rest.pop_lowest();
                Slot c;
                const uint32_t ci = tree.find(cur, m, c);
                c.avail += 1;
                tree.slots[ci].avail = c.avail;
                const double u = ismcts_ucb_value(c.score2, c.visits, P.ln_table[c.avail], P.exploration);
                if (u > best_u || (u == best_u && c.created < best_created)) {
                    best_u = u; best = ci; best_created = c.created; move = m; best_kids = c.child_mask; stored = c.player;
                }
            }
            cur = best; kids = best_kids;
            path[depth++] = cur | (stored << 24);
            g.step(move);
            if (g.chance_pending()) break;
        }
        node[j] = cur;
        store_game(j, g, rollout, depth);
    }

    // ---- one pass, in two halves (the warp votes sit between them)

    // Bit x set = this lane has a job that waits for kind x (JobWant); bit 3: a live job exists.
    ISMCTS_DEV uint32_t classify(uint32_t tid) const
    {
        uint32_t bits = 0;
#pragma unroll
        for (uint32_t q = 0; q < K; ++q) {
            const uint32_t want = (B[q * T + tid] >> 8) & 3u;
            bits |= want == kWantNothing ? 0u : (1u << want) | 8u;
        }
        return bits;
    }

    // Steps the first job of this lane that waits for `kind`.
    ISMCTS_DEV void act(uint32_t tid, uint32_t kind)
    {
        uint32_t j = 0xFFFFFFFFu;
#pragma unroll
        for (uint32_t q = 0; q < K; ++q)
            if (j == 0xFFFFFFFFu && ((B[q * T + tid] >> 8) & 3u) == kind) j = q * T + tid;
        if (j == 0xFFFFFFFFu) return;
        KwGame g;
        load_game(j, g);
        const uint32_t b = B[j];
        const bool rollout = (b >> 10) & 1u;
        const uint32_t depth = (b >> 16) & 0xFFu;
        if (kind == kWantTree) {
            if (g.terminal()) finish_iteration(j, g, depth);
            else descend(j, g, depth);
            return;
        }
        if (words_waiting(j) == 0) return;                         // dry: wait for the next refill pass
        const uint32_t w = take_word(j);
        if (kind == kWantChance) {
            // a determinisation (cloneAndRandomise) or re-deal card, or a tie-break
            const Mask set = g.chance_set();
            g.chance_step(set.kth(mulhi32(w, set.count())));
        } else {
            // simulate, solverbase.h:36-43 with RandomElement, utility.h:69-83
            const Mask legal = g.legal();
            g.play(legal.kth(mulhi32(w, legal.count())));
        }
        store_game(j, g, rollout, depth);
    }

    // Every lane generates one block for the job that is lowest on words (if below four).
    ISMCTS_DEV void refill_pass(uint32_t tid)
    {
        uint32_t j = 0xFFFFFFFFu, least = 4;
#pragma unroll
        for (uint32_t q = 0; q < K; ++q) {
            const uint32_t jq = q * T + tid;
            const uint32_t n = words_waiting(jq);
            if (((B[jq] >> 8) & 3u) != kWantNothing && n < least) { least = n; j = jq; }
        }
        if (j != 0xFFFFFFFFu) refill(j);
    }

    ISMCTS_DEV void init_lane(uint32_t tid)
    {
#pragma unroll
        for (uint32_t q = 0; q < K; ++q) {
            const uint32_t j = q * T + tid;
            done[j] = 0; node[j] = 0; R[j] = 0; A[j] = 0;
            B[j] = kWantNothing << 8;
            if (first_job + j < total_jobs) {
                reinterpret_cast<Slot*>(P.pool)[(size_t)(first_job + j) << P.log2cap] = root_slot();
                if (budget_left(j)) start_iteration(j);
            }
        }
    }

    ISMCTS_DEV static Slot root_slot()
    {
        Slot r;
        r.set_fresh();
        r.key = kRootKey; r.created = 0; r.player = 0; r.child_mask = Mask::none();
        return r;
    }

    ISMCTS_DEV void finish_lane(uint32_t tid) const
    {
#pragma unroll
        for (uint32_t q = 0; q < K; ++q) {
            const uint32_t j = q * T + tid;
            if (first_job + j < total_jobs) P.iters_done[first_job + j] = done[j];
        }
    }

    // The kind of step the next pass performs, from the numbers of lanes that can serve each kind.
    ISMCTS_DEV static uint32_t choose_kind(uint32_t n_chance, uint32_t n_move, uint32_t n_tree, uint32_t tree_thresh,
                                           uint32_t tree_period, uint32_t pass)
    {
        if (n_tree != 0 && (n_tree >= tree_thresh || (pass & (tree_period - 1u)) == tree_period - 1u ||
                            n_chance + n_move == 0)) return kWantTree;
        return n_chance >= n_move ? kWantChance : kWantMove;
    }

    // The loop of one warp (GPU): all lanes of the warp call this together.
    ISMCTS_DEV void run(uint32_t tid)
    {
        for (uint32_t pass = 0;; ++pass) {
            const uint32_t bits = classify(tid);
            if (!warp_any((bits & 8u) != 0)) break;
            const uint32_t n_chance = popc32(warp_ballot((bits & 1u) != 0)), n_move = popc32(warp_ballot((bits & 2u) != 0)),
                           n_tree = popc32(warp_ballot((bits & 4u) != 0));
            const uint32_t kind = choose_kind(n_chance, n_move, n_tree, P.tree_thresh, P.tree_period, pass);
            act(tid, kind);
            if ((pass & 3u) == 3u) refill_pass(tid);
        }
    }
};

} // namespace ismcts

/*
 * ref_harness.cpp — drives the UNMODIFIED reference (sfranzen/ismcsolver) from
 * the command line.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile against the reference sources where they lie
 * (-I/root/reference/include -I/root/reference/test plus the reference's own
 * knockoutwhist.cpp and mnkgame.cpp) into oracle/_ref/ref_harness.  It is a
 * separate executable on purpose: the reference and the new engine both
 * define namespace ISMCTS and must never be linked together.
 *
 * Uses: (1) golden statistics for the parity tests (oracle/gen_golden.py ->
 * tests/golden/), (2) the timed CPU baseline of bench.py (--impl reference and
 * cpu_baseline), both through the reference's own public API:
 * SOSolver/MOSolver::operator() (sosolver.h:30, mosolver.h:36).
 *
 * Positions are passed as the hex of the POD records of include/ismcts_b200.h.
 */
#include <ismcts/sosolver.h>
#include <ismcts/mosolver.h>
#include "common/knockoutwhist.h"
#include "common/mnkgame.h"
#include "common/goofspiel.h"
#include "common/phantommnkgame.h"

#include "../include/ismcts_b200.h"

#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

using namespace ISMCTS;

namespace
{

Card cardOf(int id) { return Card{Card::Rank(id % 13), Card::Suit(id / 13)}; }

// A Knockout Whist game put into an arbitrary position (the reference keeps
// its state in protected members, knockoutwhist.h:32-48).
struct PosedWhist : public KnockoutWhist
{
    explicit PosedWhist(ismcts_kw_state const &s) : KnockoutWhist{s.num_players}
    {
        m_numPlayers = s.num_players;
        m_players.clear();
        m_playerCards.assign(m_numPlayers, {});
        m_tricksTaken.assign(m_numPlayers, 0);
        for (unsigned p = 0; p < m_numPlayers; ++p) {
            if (s.alive_mask >> p & 1)
                m_players.push_back(p);
            m_tricksTaken[p] = s.tricks_taken[p];
            for (int c = 0; c < 52; ++c)
                if (s.hands[p] >> c & 1)
                    m_playerCards[p].push_back(cardOf(c));
        }
        m_unknownCards.clear();
        for (int c = 0; c < 52; ++c)
            if (s.unknown >> c & 1)
                m_unknownCards.push_back(cardOf(c));
        m_currentTrick.clear();
        for (unsigned i = 0; i < s.trick_len; ++i)
            m_currentTrick.emplace_back(s.trick_player[i], cardOf(s.trick_card[i]));
        m_tricksLeft = s.tricks_left;
        m_player = s.player;
        m_dealer = s.dealer;
        m_trumpSuit = Card::Suit(s.trump);
        m_requestTrump = s.request_trump != 0;
    }

    // Export the current position (used to print the reference's own opening deal).
    ismcts_kw_state pod() const
    {
        ismcts_kw_state s;
        std::memset(&s, 0, sizeof s);
        s.num_players = m_numPlayers;
        for (auto p : m_players)
            s.alive_mask |= 1u << p;
        for (unsigned p = 0; p < m_numPlayers; ++p) {
            s.tricks_taken[p] = m_tricksTaken[p];
            for (auto const &c : m_playerCards[p])
                s.hands[p] |= std::uint64_t(1) << int(c);
        }
        for (auto const &c : m_unknownCards)
            s.unknown |= std::uint64_t(1) << int(c);
        s.trick_len = m_currentTrick.size();
        for (unsigned i = 0; i < m_currentTrick.size(); ++i) {
            s.trick_player[i] = m_currentTrick[i].first;
            s.trick_card[i] = int(m_currentTrick[i].second);
        }
        s.tricks_left = m_tricksLeft;
        s.player = m_player;
        s.dealer = m_dealer;
        s.trump = m_trumpSuit;
        s.request_trump = m_requestTrump;
        return s;
    }

    unsigned survivor() const { return m_players.front(); }
    unsigned tricksLeft() const { return m_tricksLeft; }
    explicit PosedWhist(unsigned players) : KnockoutWhist{players} {}
};

struct PosedMnk : public MnkGame
{
    explicit PosedMnk(ismcts_mnk_state const &s) : MnkGame{s.m, s.n, s.k}
    {
        m_moves.clear();
        for (int c = 0; c < s.m * s.n; ++c) {
            int owner = -1;
            for (int p = 0; p < 2; ++p)
                if (s.stones[p][c >> 6] >> (c & 63) & 1)
                    owner = p;
            m_board[c / s.m][c % s.m] = owner;
            if (owner < 0)
                m_moves.push_back(c);
        }
        m_player = s.player;
        m_result = s.result2 < 0 ? -1 : s.result2 / 2.0;
    }
};

struct PosedPhantom : public PhantomMnkGame
{
    explicit PosedPhantom(ismcts_phantom_state const &s) : PhantomMnkGame{s.m, s.n, s.k}
    {
        m_moves.clear();
        for (auto &a : m_available)
            a.clear();
        for (int c = 0; c < s.m * s.n; ++c) {
            int owner = -1;
            for (int p = 0; p < 2; ++p) {
                if (s.stones[p][c >> 6] >> (c & 63) & 1)
                    owner = p;
                if (s.avail[p][c >> 6] >> (c & 63) & 1)
                    m_available[p].push_back(c);
            }
            m_board[c / s.m][c % s.m] = owner;
            if (owner < 0)
                m_moves.push_back(c);
        }
        m_player = s.player;
        m_result = s.result2 < 0 ? -1 : s.result2 / 2.0;
    }
};

template<class T>
bool fromHex(std::string const &hex, T &out)
{
    if (hex.size() != 2 * sizeof(T))
        return false;
    auto *bytes = reinterpret_cast<unsigned char *>(&out);
    for (std::size_t i = 0; i < sizeof(T); ++i)
        bytes[i] = static_cast<unsigned char>(std::stoul(hex.substr(2 * i, 2), nullptr, 16));
    return true;
}

template<class T>
std::string toHex(T const &in)
{
    auto const *bytes = reinterpret_cast<unsigned char const *>(&in);
    std::string s;
    char buf[3];
    for (std::size_t i = 0; i < sizeof(T); ++i) {
        std::snprintf(buf, sizeof buf, "%02x", bytes[i]);
        s += buf;
    }
    return s;
}

int moveId(Card const &c) { return int(c); }
int moveId(int m) { return m; }

// "[M:KH by 0, V/S/A: 12/3.0/45]" (ucb1.h:41-47) -> score and availability;
// "[M:.. by p, V/S/P: v/s/p]" (exp3.h:33-39) -> score and probability.
void parseNode(std::string const &text, double &score, double &third)
{
    auto const colon = text.rfind(": ");
    unsigned v = 0;
    std::sscanf(text.c_str() + colon + 2, "%u/%lf/%lf", &v, &score, &third);
}

template<class Move, class TreeList>
void printRootStats(TreeList const &trees, std::ostream &out)
{
    // one JSON object: {"move": [visits, score, avail_or_prob_sum], ...} summed over trees
    std::map<int, std::array<double, 3>> acc;
    for (auto const &tree : trees) {
        for (auto const &node : tree->children()) {
            double score = 0, third = 0;
            parseNode(std::string(*node), score, third);
            auto &a = acc[moveId(node->move())];
            a[0] += node->visits();
            a[1] += score;
            a[2] += third;
        }
    }
    out << "{";
    bool first = true;
    for (auto const &kv : acc) {
        out << (first ? "" : ", ") << "\"" << kv.first << "\": [" << kv.second[0] << ", " << kv.second[1] << ", "
            << kv.second[2] << "]";
        first = false;
    }
    out << "}";
}

template<class Solver, class Game, class Move>
void runStats(Game const &game, std::size_t iterations, unsigned threads, unsigned reps, bool mo)
{
    (void)mo;
    std::cout << "[";
    for (unsigned r = 0; r < reps; ++r) {
        Solver solver{iterations, threads};
        auto const t0 = std::chrono::steady_clock::now();
        Move const best = solver(game);
        double const secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << (r ? ",\n " : "") << "{\"best\": " << moveId(best) << ", \"seconds\": " << secs << ", \"root\": ";
        std::ostringstream oss;
        oss.precision(17);
        printRootStats<Move>(solver.currentTrees(), oss);
        std::cout << oss.str() << "}";
    }
    std::cout << "]\n";
}

// Sequential solvers take one argument; give them the same two-argument shape.
template<class Move>
struct SeqSO : SOSolver<Move, Sequential>
{
    SeqSO(std::size_t n, unsigned) : SOSolver<Move, Sequential>{n} {}
};
int usage()
{
    std::fprintf(stderr,
        "ref_harness kw-opening <players>\n"
        "ref_harness kw-stats <hex kw_state> <so|mo|mo-exp3> <seq|root|tree> <iterations> <threads> <reps>\n"
        "ref_harness mnk-stats <hex mnk_state> <so|mo> <seq|root|tree> <iterations> <threads> <reps>\n"
        "ref_harness kw-playouts <hex kw_state> <count>\n"
        "ref_harness kw-time <players> <so|mo> <seq|root|tree> <iterations> <threads> <reps>\n"
        "ref_harness kw-vs-random <players> <seq|root> <iterations> <threads> <games>\n"
        "ref_harness kw-moves <count|ms> <budget> <threads>      (hex kw_state per line on stdin -> move per line)\n"
        "ref_harness phantom-stats <hex phantom_state> <so|mo> <seq|root|tree> <iterations> <threads> <reps>\n"
        "ref_harness goof-stats <prize rank 0-12> <player 0 bid rank or -1> <so|mo> <iterations> <reps>\n"
        "ref_harness cores\n");
    return 2;
}

// MOSolver::currentTrees returns one map per worker; reduce to the acting player's roots.
template<class Solver, class Game, class Move>
void runStatsMO(Game const &game, std::size_t iterations, unsigned threads, unsigned reps)
{
    std::cout << "[";
    for (unsigned r = 0; r < reps; ++r) {
        Solver solver{iterations, threads};
        auto const t0 = std::chrono::steady_clock::now();
        Move const best = solver(game);
        double const secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::vector<std::shared_ptr<Node<Move>>> roots;
        for (auto const &map : solver.currentTrees())
            roots.push_back(map.at(game.currentPlayer()));
        std::cout << (r ? ",\n " : "") << "{\"best\": " << moveId(best) << ", \"seconds\": " << secs << ", \"root\": ";
        std::ostringstream oss;
        oss.precision(17);
        printRootStats<Move>(roots, oss);
        std::cout << oss.str() << "}";
    }
    std::cout << "]\n";
}

template<class Move>
struct SeqMO2 : MOSolver<Move, Sequential>
{
    SeqMO2(std::size_t n, unsigned) : MOSolver<Move, Sequential>{n} {}
};
template<class Move>
struct SeqMOExp3 : MOSolver<Move, Sequential, EXP3>
{
    SeqMOExp3(std::size_t n, unsigned) : MOSolver<Move, Sequential, EXP3>{n} {}
};

template<class Game, class Move>
int dispatchStats(Game const &game, std::string const &solver, std::string const &policy, std::size_t it,
                  unsigned threads, unsigned reps)
{
    if (solver == "so") {
        if (policy == "seq") runStats<SeqSO<Move>, Game, Move>(game, it, threads, reps, false);
        else if (policy == "root") runStats<SOSolver<Move, RootParallel>, Game, Move>(game, it, threads, reps, false);
        else if (policy == "tree") runStats<SOSolver<Move, TreeParallel>, Game, Move>(game, it, threads, reps, false);
        else return usage();
    } else if (solver == "mo") {
        if (policy == "seq") runStatsMO<SeqMO2<Move>, Game, Move>(game, it, threads, reps);
        else if (policy == "root") runStatsMO<MOSolver<Move, RootParallel>, Game, Move>(game, it, threads, reps);
        else if (policy == "tree") runStatsMO<MOSolver<Move, TreeParallel>, Game, Move>(game, it, threads, reps);
        else return usage();
    } else if (solver == "mo-exp3") {
        if (policy == "seq") runStatsMO<SeqMOExp3<Move>, Game, Move>(game, it, threads, reps);
        else if (policy == "root") runStatsMO<MOSolver<Move, RootParallel, EXP3>, Game, Move>(game, it, threads, reps);
        else return usage();
    } else if (solver == "so-exp3") {
        if (policy == "root") runStats<SOSolver<Move, RootParallel, EXP3>, Game, Move>(game, it, threads, reps, false);
        else return usage();
    } else {
        return usage();
    }
    return 0;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc < 2)
        return usage();
    std::string const cmd = argv[1];

    if (cmd == "cores") {
        std::cout << std::thread::hardware_concurrency() << "\n";
        return 0;
    }
    if (cmd == "kw-opening" && argc == 3) {
        // The reference's own opening deal on a fresh main thread (default-seeded
        // mt19937, knockoutwhist.cpp:28-32) as a POD record.
        PosedWhist game{unsigned(std::atoi(argv[2]))};
        std::cout << toHex(game.pod()) << "\n";
        return 0;
    }
    if (cmd == "kw-stats" && argc == 8) {
        ismcts_kw_state s;
        if (!fromHex(argv[2], s))
            return usage();
        PosedWhist game{s};
        return dispatchStats<PosedWhist, Card>(game, argv[3], argv[4], std::strtoull(argv[5], nullptr, 10),
                                              unsigned(std::atoi(argv[6])), unsigned(std::atoi(argv[7])));
    }
    if (cmd == "mnk-stats" && argc == 8) {
        ismcts_mnk_state s;
        if (!fromHex(argv[2], s))
            return usage();
        PosedMnk game{s};
        return dispatchStats<PosedMnk, int>(game, argv[3], argv[4], std::strtoull(argv[5], nullptr, 10),
                                           unsigned(std::atoi(argv[6])), unsigned(std::atoi(argv[7])));
    }
    if (cmd == "kw-playouts" && argc == 4) {
        // Determinise + uniform random playout, the reference's own game code
        // (cloneAndRandomise + validMoves/doMove): winner and length histograms.
        ismcts_kw_state s;
        if (!fromHex(argv[2], s))
            return usage();
        PosedWhist const root{s};
        unsigned long const count = std::strtoul(argv[3], nullptr, 10);
        std::vector<unsigned long> wins(8, 0), rounds(9, 0);
        unsigned long plies = 0, maxPlies = 0;
        for (unsigned long i = 0; i < count; ++i) {
            auto clone = root.cloneAndRandomise(root.currentPlayer());
            unsigned long n = 0, trumpCalls = 0;
            for (;;) {
                auto const moves = clone->validMoves();
                if (moves.empty())
                    break;
                // a trump request offers (Two, suit) x 4 (knockoutwhist.cpp:18-24): count the rounds
                if (moves.size() == 4 && moves[0].rank == Card::Two && moves[1].rank == Card::Two &&
                    moves[0].suit == Card::Clubs && moves[1].suit == Card::Diamonds && moves[3].suit == Card::Spades &&
                    moves[2].rank == Card::Two && moves[3].rank == Card::Two)
                    ++trumpCalls;
                clone->doMove(randomElement(moves));
                ++n;
            }
            plies += n;
            maxPlies = std::max(maxPlies, n);
            for (unsigned p = 0; p < s.num_players; ++p)
                if (clone->getResult(p) == 1)
                    ++wins[p];
            ++rounds[std::min<unsigned long>(trumpCalls, 8)];
        }
        std::cout << "{\"count\": " << count << ", \"mean_plies\": " << double(plies) / count
                  << ", \"max_plies\": " << maxPlies << ", \"wins\": [";
        for (unsigned p = 0; p < s.num_players; ++p)
            std::cout << (p ? ", " : "") << wins[p];
        std::cout << "], \"trump_requests_hist\": [";
        for (unsigned i = 0; i < 9; ++i)
            std::cout << (i ? ", " : "") << rounds[i];
        std::cout << "]}\n";
        return 0;
    }
    if (cmd == "kw-time" && argc == 8) {
        // Timed CPU baseline: steady_clock around operator() only.
        unsigned const players = unsigned(std::atoi(argv[2]));
        std::string const solver = argv[3], policy = argv[4];
        std::size_t const it = std::strtoull(argv[5], nullptr, 10);
        unsigned const threads = unsigned(std::atoi(argv[6]));
        unsigned const reps = unsigned(std::atoi(argv[7]));
        std::cout << "[";
        for (unsigned r = 0; r < reps; ++r) {
            PosedWhist game{players}; // a fresh deal from the reference's own game RNG per repetition
            double secs = 0;
            auto timeIt = [&](auto &s) {
                auto const t0 = std::chrono::steady_clock::now();
                Card const best = s(game);
                secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                return int(best);
            };
            int best = -1;
            if (solver == "so" && policy == "seq") { SOSolver<Card, Sequential> s{it}; best = timeIt(s); }
            else if (solver == "so" && policy == "root") { SOSolver<Card, RootParallel> s{it, threads}; best = timeIt(s); }
            else if (solver == "so" && policy == "tree") { SOSolver<Card, TreeParallel> s{it, threads}; best = timeIt(s); }
            else if (solver == "mo" && policy == "seq") { MOSolver<Card, Sequential> s{it}; best = timeIt(s); }
            else if (solver == "mo" && policy == "root") { MOSolver<Card, RootParallel> s{it, threads}; best = timeIt(s); }
            else if (solver == "mo-exp3" && policy == "root") { MOSolver<Card, RootParallel, EXP3> s{it, threads}; best = timeIt(s); }
            else return usage();
            std::cout << (r ? ", " : "") << "{\"seconds\": " << secs << ", \"iterations\": " << it
                      << ", \"best\": " << best << "}";
            std::cout.flush();
        }
        std::cout << "]\n";
        return 0;
    }
    if (cmd == "phantom-stats" && argc == 8) {
        ismcts_phantom_state s;
        if (!fromHex(argv[2], s))
            return usage();
        PosedPhantom game{s};
        return dispatchStats<PosedPhantom, int>(game, argv[3], argv[4], std::strtoull(argv[5], nullptr, 10),
                                               unsigned(std::atoi(argv[6])), unsigned(std::atoi(argv[7])));
    }
    if (cmd == "goof-stats" && argc == 7) {
        // Goofspiel keeps its state private: take fresh games until the first prize is the wanted one,
        // then (optionally) let player 0 bid.  Every move is simultaneous, so the solvers select with
        // EXP3 (solverbase.h:58-64).  Root statistics of the player to move, as for kw-stats.
        int const prize = std::atoi(argv[2]), bid = std::atoi(argv[3]);
        std::string const solver = argv[4];
        std::size_t const it = std::strtoull(argv[5], nullptr, 10);
        unsigned const reps = unsigned(std::atoi(argv[6]));
        std::cout << "[";
        for (unsigned r = 0; r < reps; ++r) {
            std::unique_ptr<Goofspiel> game;
            for (;;) {
                game = std::make_unique<Goofspiel>();
                if (int(game->validMoves().front().rank) == prize)
                    break;
            }
            game->doMove(game->validMoves().front());
            if (bid >= 0)
                game->doMove(Card{Card::Rank(bid), Card::Spades});
            std::ostringstream oss;
            oss.precision(17);
            int best = -1;
            if (solver == "so") {
                SOSolver<Card, Sequential> s {it};
                best = int(s(*game));
                printRootStats<Card>(s.currentTrees(), oss);
            } else {
                MOSolver<Card, Sequential> s {it};
                best = int(s(*game));
                std::vector<std::shared_ptr<Node<Card>>> roots;
                for (auto const &map : s.currentTrees())
                    roots.push_back(map.at(game->currentPlayer()));
                printRootStats<Card>(roots, oss);
            }
            std::cout << (r ? ",\n " : "") << "{\"best\": " << best << ", \"seconds\": 0, \"root\": " << oss.str() << "}";
        }
        std::cout << "]\n";
        return 0;
    }
    if (cmd == "kw-moves" && argc == 5) {
        // The reference as a player: one position (hex record) per line on stdin, one decision per line on
        // stdout — SOSolver<Card, RootParallel> with a budget of <budget> iterations ("count") or milliseconds
        // ("ms") per decision on <threads> threads, stock operator() (sosolver.h:30-36).
        std::string const mode = argv[2];
        double const budget = std::atof(argv[3]);
        unsigned const threads = unsigned(std::atoi(argv[4]));
        std::string line;
        while (std::getline(std::cin, line)) {
            ismcts_kw_state s;
            if (!fromHex(line, s))
                return usage();
            PosedWhist game{s};
            auto const t0 = std::chrono::steady_clock::now();
            int best = -1;
            if (mode == "ms") {
                SOSolver<Card, RootParallel> solver{std::chrono::duration<double>(budget * 1e-3), threads};
                best = int(solver(game));
            } else {
                SOSolver<Card, RootParallel> solver{std::size_t(budget), threads};
                best = int(solver(game));
            }
            double const secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::cout << best << " " << secs << std::endl;
        }
        return 0;
    }
    if (cmd == "kw-vs-random" && argc == 7) {
        // test/benchmark.cpp:73-89,144-151: player 0 is the solver, the others move uniformly at random
        unsigned const players = unsigned(std::atoi(argv[2]));
        std::string const policy = argv[3];
        std::size_t const it = std::strtoull(argv[4], nullptr, 10);
        unsigned const threads = unsigned(std::atoi(argv[5]));
        unsigned const games = unsigned(std::atoi(argv[6]));
        std::vector<unsigned> wins(players, 0);
        unsigned long decisions = 0;
        double seconds = 0;
        SOSolver<Card, Sequential> seq {it};
        SOSolver<Card, RootParallel> root {it, threads};
        for (unsigned g = 0; g < games; ++g) {
            KnockoutWhist game {players};
            for (;;) {
                auto const moves = game.validMoves();
                if (moves.empty())
                    break;
                if (game.currentPlayer() == 0) {
                    auto const t0 = std::chrono::steady_clock::now();
                    Card const move = policy == "root" ? root(game) : seq(game);
                    seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    ++decisions;
                    game.doMove(move);
                } else {
                    game.doMove(randomElement(moves));
                }
            }
            for (unsigned p = 0; p < players; ++p)
                if (game.getResult(p) == 1)
                    ++wins[p];
        }
        std::cout << "{\"games\": " << games << ", \"iterations\": " << it << ", \"decisions\": " << decisions
                  << ", \"seconds\": " << seconds << ", \"wins\": [";
        for (unsigned p = 0; p < players; ++p)
            std::cout << (p ? ", " : "") << wins[p];
        std::cout << "]}\n";
        return 0;
    }
    return usage();
}

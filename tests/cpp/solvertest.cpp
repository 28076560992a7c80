// solvertest.cpp — API conformance of the CUDA execution policies.
//
// Restates the assertions of the reference's test/solvertest.cpp (:64-138) — constructors and
// getters, setters, "operator() returns a valid move" by iteration count and by time, and the
// P1DrawOrLose position that every solver must answer with move 2 from 16 iterations — for
// {SOSolver, MOSolver} x {CudaRootParallel, CudaTreeParallel} (x EXP3 where the reference's tests use it), plus the game checks of
// test/gametest.cpp that concern the host game classes (:92-106,120-130,184-206).
//
//   solvertest            all checks; needs a CUDA device
//   solvertest --host     only the checks that run without a device, and that a search without a
//                         device fails loudly (no CPU fallback)
#include <ismcts/sosolver.h>
#include <ismcts/mosolver.h>
#include <ismcts/games/knockoutwhist.h>
#include <ismcts/games/mnkgame.h>
#include <ismcts/games/goofspiel.h>
#include <ismcts/games/phantommnkgame.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>

using namespace ISMCTS;
using namespace std::chrono_literals;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                                          \
    do {                                                                                     \
        ++checks;                                                                            \
        if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define CHECK_THROWS_AS(expr, Exception)                                                     \
    do {                                                                                     \
        ++checks;                                                                            \
        bool caught = false;                                                                 \
        try { expr; } catch (Exception const &) { caught = true; } catch (...) {}            \
        if (!caught) { ++failures; std::printf("FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #Exception); } \
    } while (0)

namespace
{

std::size_t const numGames = 10;
auto const iterationTime = 5ms;

// solvertest.cpp:45-60: board [[-1,1,-1],[1,0,0],[0,0,1]], player 1 to move; 2 draws, 0 loses
struct P1DrawOrLose : public MnkGame
{
    P1DrawOrLose()
    {
        for (int move : {4, 1, 5, 3, 6, 8, 7})
            doMove(move);
    }
};

// a game without a device form
struct OpaqueGame : public Game<int>
{
    Clone cloneAndRandomise(Player) const override { return std::make_unique<OpaqueGame>(*this); }
    Player currentPlayer() const override { return 0; }
    std::vector<int> validMoves() const override { return {0, 1}; }
    void doMove(int const) override {}
    double getResult(Player) const override { return 0; }
};

template<class Solver>
void testConstruction()
{
    // solvertest.cpp:64-78
    {
        Solver solver {numGames};
        CHECK(solver.iterationCount() == numGames);
        CHECK(solver.iterationTime() == std::chrono::duration<double>::zero());
    }
    {
        Solver solver {iterationTime};
        CHECK(solver.iterationTime() == iterationTime);
        CHECK(solver.iterationCount() == 0);
    }
    {
        Solver solver;
        CHECK(solver.iterationCount() == 1000);
        CHECK(solver.numThreads() >= 1);
    }
    // solvertest.cpp:80-99
    Solver solver {numGames};
    solver.setIterationTime(iterationTime);
    CHECK(solver.iterationTime() == iterationTime);
    CHECK(solver.iterationCount() == 0);
    solver.setIterationCount(numGames);
    CHECK(solver.iterationTime() == std::chrono::duration<double>::zero());
    CHECK(solver.iterationCount() == numGames);
}

template<class Game, class Move>
bool isValid(Game const &game, Move const &move)
{
    auto const moves = game.validMoves();
    return std::find(moves.begin(), moves.end(), move) != moves.end();
}

template<class Solver>
void testKnockoutWhist()
{
    // solvertest.cpp:101-119
    KnockoutWhist game {2};
    {
        Solver solver {numGames};
        Card const move = solver(game);
        CHECK(isValid(game, move));
    }
    {
        Solver solver {iterationTime};
        Card const move = solver(game);
        CHECK(isValid(game, move));
        CHECK(solver.lastStatistics().iterations() > 0);
    }
    // a full game played by the solver against itself terminates within the turn bound
    Solver solver {2000};
    KnockoutWhist four {4, 12345};
    unsigned turns = 0;
    while (!four.validMoves().empty() && turns < 6 + 4 * 28) {
        Card const move = solver(four);
        CHECK(isValid(four, move));
        four.doMove(move);
        ++turns;
    }
    CHECK(four.validMoves().empty());
}

template<class Solver>
void testGoofspiel()
{
    // solvertest.cpp:121-128: simultaneous tree policy; advance to a regular player first
    Goofspiel game;
    game.doMove(game.validMoves().front());
    {
        Solver solver {numGames};
        Card const move = solver(game);
        CHECK(isValid(game, move));
    }
    {
        Solver solver {iterationTime};
        Card const move = solver(game);
        CHECK(isValid(game, move));
    }
    // a whole game: the environment's moves are searched as well (one legal move each)
    Solver solver {500};
    Goofspiel full {42};
    unsigned turns = 0;
    while (!full.validMoves().empty() && turns < 60) {
        Card const move = solver(full);
        CHECK(isValid(full, move));
        full.doMove(move);
        ++turns;
    }
    CHECK(turns == 52);
    CHECK(full.getResult(0) + full.getResult(1) == 1.0);
}

template<class Solver>
void testP1DrawOrLose()
{
    // solvertest.cpp:131-138
    P1DrawOrLose game;
    CHECK(game.currentPlayer() == 1);
    CHECK((game.validMoves() == std::vector<int>{0, 2}));
    for (int rep = 0; rep < 5; ++rep) {
        Solver solver {16};
        CHECK(solver(game) == 2);
    }
}

template<class Solver>
void testTrees()
{
    P1DrawOrLose game;
    Solver solver {500};
    solver.setSeed(7);
    int const move = solver(game);
    CHECK(move == 2);
    auto const &stats = solver.lastStatistics();
    CHECK(stats.iterations() == 500);
    CHECK(stats.visits[0] + stats.visits[2] == 500);
}

void testHostOnly()
{
    // gametest.cpp:92-98
    KnockoutWhist whist {2};
    CHECK(whist.currentPlayer() == 0);
    CHECK((whist.players() == std::vector<unsigned>{0, 1}));
    CHECK(whist.validMoves().size() == 7);
    // gametest.cpp:120-130
    {
        KnockoutWhist game {2, 99};
        game.doMove(game.validMoves().front());
        CHECK(game.currentPlayer() == 1);
    }
    {
        KnockoutWhist game {2, 99};
        auto const hand = game.validMoves();
        Card missing {Card::Two, Card::Clubs};
        for (int id = 0; id < 52; ++id) {
            missing = Card::fromId(id);
            if (std::find(hand.begin(), hand.end(), missing) == hand.end())
                break;
        }
        CHECK_THROWS_AS(game.doMove(missing), std::out_of_range);
    }
    // cloneAndRandomise keeps what the observer sees (knockoutwhist.cpp:56-76)
    {
        KnockoutWhist game {4, 5};
        auto clone = game.cloneAndRandomise(0);
        auto const *c = dynamic_cast<KnockoutWhist const *>(clone.get());
        CHECK(c != nullptr);
        CHECK(c->state().hands[0] == game.state().hands[0]);
        CHECK(__builtin_popcountll(c->state().hands[1]) == 7);
        CHECK((c->state().hands[0] & c->state().hands[1]) == 0);
    }
    // gametest.cpp:100-106,184-206
    MnkGame mnk;
    CHECK(mnk.currentPlayer() == 0);
    CHECK(mnk.validMoves().size() == 9);
    CHECK_THROWS_AS(mnk.doMove(9), std::out_of_range);
    int const wins[8][5] = {{0,3,1,4,2},{3,0,4,1,5},{6,0,7,1,8},{0,2,3,4,6},{1,0,4,2,7},{2,0,5,1,8},{0,1,4,2,8},{2,0,4,1,6}};
    for (auto const &sequence : wins) {
        MnkGame game;
        for (int move : sequence)
            game.doMove(move);
        CHECK(game.validMoves().empty());
        CHECK(game.getResult(0) == 1);
    }
    // gametest.cpp:108-114,209-257
    {
        Goofspiel goof;
        CHECK(goof.currentPlayer() == 2);
        CHECK((goof.players() == std::vector<unsigned>{0, 1, 2}));
        CHECK(goof.validMoves().size() == 1);
        goof.doMove(goof.validMoves().front());
        CHECK(goof.currentPlayer() == 0);
        CHECK(goof.validMoves().size() == 13);
        Goofspiel equal = goof;
        for (int i = 0; i < 3; ++i)
            equal.doMove(equal.validMoves().front());
        CHECK(equal.getResult(0) == equal.getResult(1));
        Goofspiel unequal = goof;
        for (int i : {0, 1})
            unequal.doMove(unequal.validMoves()[i]);
        unequal.doMove(unequal.validMoves().front());
        CHECK(unequal.getResult(0) == 0);
        CHECK(unequal.getResult(1) != 0);
        unsigned turns = 0;
        while (!goof.validMoves().empty() && turns <= 13) {
            if (goof.currentPlayer() == 0)
                ++turns;
            goof.doMove(goof.validMoves().front());
        }
        CHECK(turns == 13);
        CHECK(goof.validMoves().empty());
        CHECK(goof.currentMoveSimultaneous());
    }
    // gametest.cpp:100-106,184-206 for the phantom game, and :264-279 (cloneAndRandomise keeps known positions)
    {
        PhantomMnkGame phantom;
        CHECK(phantom.currentPlayer() == 0);
        CHECK(phantom.validMoves().size() == 9);
        CHECK_THROWS_AS(phantom.doMove(9), std::out_of_range);
        for (auto const &sequence : wins) {
            PhantomMnkGame game;
            for (int move : sequence)
                game.doMove(move);
            CHECK(game.validMoves().empty());
            CHECK(game.getResult(0) == 1);
        }
        PhantomMnkGame game {3, 3, 3, 11};
        for (int move : {4, 4, 0, 6, 2})
            game.doMove(move);
        CHECK(game.currentPlayer() == 0);
        for (int rep = 0; rep < 20; ++rep) {
            auto clone = game.cloneAndRandomise(0);
            auto const &c = dynamic_cast<PhantomMnkGame const &>(*clone).state();
            CHECK(c.avail[0][0] == game.state().avail[0][0]);
            CHECK(__builtin_popcountll(c.avail[1][0]) == __builtin_popcountll(game.state().avail[1][0]));
            CHECK(__builtin_popcountll(c.stones[1][0]) == 2);
            CHECK(c.stones[0][0] == ((1u << 4) | (1u << 6)));
        }
    }
    // node.h: the textual form of nodes (ucb1.h:41-47, docs/node.md:13-22)
    UCBNode<int> root;
    auto child = std::make_unique<UCBNode<int>>(6, 0);
    child->setVisits(988);
    child->setStatistics(988.0, 998);
    root.addChild(std::move(child));
    CHECK(root.treeToString() == "[M:0 by 0, V/S/A: 0/0.0/1]\n| [M:6 by 0, V/S/A: 988/988.0/998]\n");
    CHECK(root.height() == 1);
    CHECK(root.children().front()->depth() == 1);
    CHECK((root.untriedMoves({5, 6, 7}) == std::vector<int>{5, 7}));
    // an opaque host game has no device form: a clear error, not a CPU fallback
    OpaqueGame opaque;
    SOSolver<int, CudaRootParallel> solver {10};
    CHECK_THROWS_AS(solver(opaque), std::invalid_argument);
}

} // namespace

int main(int argc, char **argv)
{
    bool const hostOnly = argc > 1 && std::strcmp(argv[1], "--host") == 0;
    //   solvertest --devices N   additionally: setDevices({0..N-1}) against one device
    testConstruction<SOSolver<Card, CudaRootParallel>>();
    testConstruction<SOSolver<Card, CudaTreeParallel>>();
    testConstruction<MOSolver<Card, CudaRootParallel>>();
    testConstruction<MOSolver<int, CudaRootParallel, EXP3>>();
    testConstruction<MOSolver<Card, CudaTreeParallel>>();
    testHostOnly();
    if (hostOnly) {
        // without a device every search throws
        KnockoutWhist game {2};
        SOSolver<Card, CudaRootParallel> solver {10};
        int devices = 0;
        ismcts_ctx *probe = nullptr;
        if (ismcts_create(0, &probe) == ISMCTS_OK) { devices = 1; ismcts_destroy(probe); }
        if (!devices)
            CHECK_THROWS_AS(solver(game), CudaError);
    } else {
        testKnockoutWhist<SOSolver<Card, CudaRootParallel>>();
        testKnockoutWhist<SOSolver<Card, CudaTreeParallel>>();
        testKnockoutWhist<MOSolver<Card, CudaRootParallel>>();
        testKnockoutWhist<MOSolver<Card, CudaRootParallel, EXP3>>();
        testKnockoutWhist<MOSolver<Card, CudaTreeParallel>>();
        testKnockoutWhist<SOSolver<Card, CudaTreeParallel, EXP3>>();
        testKnockoutWhist<MOSolver<Card, CudaTreeParallel, EXP3>>();
        testGoofspiel<SOSolver<Card, CudaRootParallel>>();
        testGoofspiel<MOSolver<Card, CudaRootParallel>>();
        testGoofspiel<SOSolver<Card, CudaTreeParallel>>();
        testGoofspiel<MOSolver<Card, CudaTreeParallel>>();
        {
            // the phantom game through both solvers: a valid move by count and by time
            PhantomMnkGame game {3, 3, 3, 5};
            for (int move : {4, 4, 0, 6, 2})
                game.doMove(move);
            SOSolver<int, CudaRootParallel> so {2000};
            CHECK(isValid(game, so(game)));
            MOSolver<int, CudaRootParallel> mo {iterationTime};
            CHECK(isValid(game, mo(game)));
            // player 0 holds 4 and 6: cell 2 completes the diagonal unless player 1 is there
            SOSolver<int, CudaRootParallel> deep {20000, 8};
            int const move = deep(game);
            CHECK(isValid(game, move));
        }
        testP1DrawOrLose<SOSolver<int, CudaRootParallel>>();
        testP1DrawOrLose<SOSolver<int, CudaTreeParallel>>();
        testP1DrawOrLose<MOSolver<int, CudaRootParallel>>();
        testP1DrawOrLose<MOSolver<int, CudaTreeParallel>>();
        testTrees<SOSolver<int, CudaRootParallel>>();
        testTrees<SOSolver<int, CudaTreeParallel>>();
        testTrees<MOSolver<int, CudaRootParallel>>();
        testTrees<MOSolver<int, CudaTreeParallel>>();
        {
            // currentTrees(): the device pool as host nodes (sosolver.h:38-41)
            P1DrawOrLose game;
            SOSolver<int, CudaRootParallel> solver {400, 2};
            solver(game);
            auto const trees = solver.currentTrees();
            CHECK(trees.size() == 2);
            unsigned visits = 0;
            for (auto const &tree : trees)
                for (auto const &node : tree->children())
                    visits += node->visits();
            CHECK(visits == 400);
            CHECK(trees.front()->treeToString().find("[M:2 by 1, V/S/A: ") != std::string::npos);
            MOSolver<int, CudaRootParallel> mo {300, 1};
            mo(game);
            auto const maps = mo.currentTrees();
            CHECK(maps.size() == 1 && maps.front().size() == 2);
        }
    }
    if (!hostOnly) {
        // batched operator(): every game gets the full budget, one engine call (sosolver.h:30-36 per game)
        std::vector<std::unique_ptr<KnockoutWhist>> games;
        std::vector<Game<Card> const *> roots;
        std::vector<POMGame<Card> const *> pomRoots;
        for (unsigned g = 0; g < 12; ++g) {
            games.push_back(std::make_unique<KnockoutWhist>(2 + g % 6, 100 + g));
            for (unsigned ply = 0; ply < 3 * g && !games.back()->validMoves().empty(); ++ply)
                games.back()->doMove(games.back()->validMoves().front());
            roots.push_back(games.back().get());
            pomRoots.push_back(games.back().get());
        }
        SOSolver<Card, CudaRootParallel> batch {20000};
        CHECK(batch.lanesPerTree() == 0);     // chosen per search
        batch.setLanesPerTree(1);     // one lane per tree: a seeded search is reproducible bit for bit
        batch.setSeed(7);
        auto const moves = batch(roots);
        CHECK(moves.size() == roots.size());
        CHECK(batch.lastBatchStatistics().size() == roots.size());
        for (std::size_t g = 0; g < roots.size(); ++g) {
            CHECK(isValid(*roots[g], moves[g]));
            CHECK(batch.lastBatchStatistics()[g].iterations() == 20000);
        }
        auto const first = batch.lastBatchStatistics().front();
        SOSolver<Card, CudaRootParallel> single {20000};
        single.setLanesPerTree(1);
        single.setSeed(7);
        CHECK(single(*roots.front()) == moves.front());
        CHECK(single.lastStatistics().visits == first.visits);
        CHECK(single.lastStatistics().score2 == first.score2);
        MOSolver<Card, CudaRootParallel> moBatch {5000};
        auto const moMoves = moBatch(pomRoots);
        for (std::size_t g = 0; g < roots.size(); ++g)
            CHECK(isValid(*roots[g], moMoves[g]));
        SOSolver<Card, CudaTreeParallel> treeBatch {4000, 64};
        auto const treeMoves = treeBatch(roots);
        for (std::size_t g = 0; g < roots.size(); ++g)
            CHECK(isValid(*roots[g], treeMoves[g]));
        CHECK(batch(std::vector<Game<Card> const *>{}).empty());
        // the default geometry: one tree per host thread, a block of lanes on each (virtual loss inside a tree)
        SOSolver<Card, CudaRootParallel> blocks {200000};
        CHECK(blocks.numThreads() == std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 200000 / 1024));
        auto const blockMoves = blocks(roots);
        for (std::size_t g = 0; g < roots.size(); ++g) {
            CHECK(isValid(*roots[g], blockMoves[g]));
            CHECK(blocks.lastBatchStatistics()[g].iterations() == 200000);
        }
        // a budget that takes more trees than the GPU holds lanes (persistent grid) through the C++ layer
        SOSolver<Card, CudaRootParallel> wide {400000, 20000};
        wide.setSeed(3);
        auto const wideMoves = wide(roots);
        for (std::size_t g = 0; g < roots.size(); ++g) {
            CHECK(isValid(*roots[g], wideMoves[g]));
            CHECK(wide.lastBatchStatistics()[g].iterations() == 400000);
        }
    }
    if (argc > 2 && std::strcmp(argv[1], "--devices") == 0) {
        // setDevices: the trees of a search sharded over several GPUs, root arrays summed on the host —
        // identical to the search on one GPU at a fixed seed (tree ids are global)
        int const n = std::atoi(argv[2]);
        std::vector<int> devices;
        for (int d = 0; d < n; ++d)
            devices.push_back(d);
        KnockoutWhist game {4, 42};
        SOSolver<Card, CudaRootParallel> one {100000, 64}, many {100000, 64};
        one.setLanesPerTree(1);
        many.setLanesPerTree(1);
        one.setSeed(11);
        many.setSeed(11);
        many.setDevices(devices);
        Card const a = one(game), b = many(game);
        CHECK(a == b);
        CHECK(one.lastStatistics().visits == many.lastStatistics().visits);
        CHECK(one.lastStatistics().score2 == many.lastStatistics().score2);
        CHECK(one.lastStatistics().available == many.lastStatistics().available);
        CHECK(many.lastStatistics().iterations() == 100000);
        MOSolver<Card, CudaRootParallel, EXP3> moOne {30000, 32}, moMany {30000, 32};
        moOne.setLanesPerTree(1);
        moMany.setLanesPerTree(1);
        moOne.setSeed(5);
        moMany.setSeed(5);
        moMany.setDevices(devices);
        CHECK(moOne(game) == moMany(game));
        CHECK(moOne.lastStatistics().visits == moMany.lastStatistics().visits);
        std::printf("devices: %d\n", n);
    }
    std::printf("%d checks, %d failures\n", checks, failures);
    return failures ? 1 : 0;
}

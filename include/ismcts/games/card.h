/*
 * card.h — playing card, the Move type of Knockout Whist.
 * Same public shape as the reference's Card (test/common/card.h:12-60): rank Two..Ace,
 * suit Clubs..Spades, int(card) = rank + 13 * suit (which is also the kernels' move id).
 */
#ifndef ISMCTS_GAMES_CARD_H
#define ISMCTS_GAMES_CARD_H

#include "../game.h"

#include <ostream>
#include <string>

struct Card
{
    enum Rank : int { Two = 0, Three, Four, Five, Six, Seven, Eight, Nine, Ten, Jack, Queen, King, Ace };
    enum Suit : int { Clubs = 0, Diamonds, Hearts, Spades };

    Rank rank {Two};
    Suit suit {Clubs};

    Card() = default;
    constexpr Card(Rank r, Suit s) : rank{r}, suit{s} {}

    static Card fromId(int id) { return Card{static_cast<Rank>(id % 13), static_cast<Suit>(id / 13)}; }

    explicit operator int() const { return rank + 13 * suit; }

    operator std::string() const
    {
        static const char ranks[] = "23456789TJQKA";
        static const char suits[] = "CDHS";
        return std::string{ranks[rank]} + suits[suit];
    }

    bool operator==(Card const &o) const { return rank == o.rank && suit == o.suit; }
    bool operator!=(Card const &o) const { return !(*this == o); }
    bool operator<(Card const &o) const { return int(*this) < int(o); }
};

inline std::ostream &operator<<(std::ostream &out, Card const &card) { return out << std::string(card); }

namespace ISMCTS
{
template<>
struct MoveCodec<Card>
{
    static std::uint32_t toId(Card const &c) { return static_cast<std::uint32_t>(int(c)); }
    static Card fromId(std::uint32_t id) { return Card::fromId(static_cast<int>(id)); }
};
} // ISMCTS

#endif // ISMCTS_GAMES_CARD_H

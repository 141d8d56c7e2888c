"""`Env::render` of the four tabular envs and `Agent::example`, host side.

Pure functions of an env's state, written from the reference's text (paths under its `src/`); the strings are what
the reference prints.  Nothing here touches the device: `Agent.example` (api.py) drives the step-level C-ABI calls and
hands the states it downloads to these functions.
"""

TAXI_MAP = ("+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |", "+---------+")   # env/taxi.rs:21-29
CLIFF_MAP = "____________\n____________\n____________\n@!!!!!!!!!!G"                                                    # env/cliff_walking.rs:20


def _skip_newlines(text, pos):
    """The reference's cursor fix-up (e.g. env/taxi.rs:164-168): walk the newline offsets in order and push the cursor
    one to the right for each one at or before it — the cursor moves WHILE the offsets are walked."""
    for i, ch in enumerate(text):
        if ch == "\n" and pos >= i:
            pos += 1
    return pos


def _put(text, pos, ch):
    return text[:pos] + ch + text[pos + 1:]


def render_taxi(curr_obs):
    """env/taxi.rs:161-172: the map with a 'T' on the taxi's cell (passenger and destination are not drawn)."""
    row, col = curr_obs // 100, (curr_obs // 20) % 5                  # decode, taxi.rs:44-55
    text = "\n".join(TAXI_MAP)
    pos = _skip_newlines(text, 11 * (row + 1) + (2 * col + 1))          # from_2d_to_1d(11, row + 1, 2 * col + 1), utils.rs:45-47
    return _put(text, pos, "T")


def render_frozen_lake(map_rows, player_pos):
    """env/frozen_lake.rs:136-149: every 'S' becomes 'F', then '@' on the player's cell."""
    text = "\n".join(map_rows).replace("S", "F")
    return _put(text, _skip_newlines(text, player_pos), "@")


def render_cliff_walking(player_pos):
    """env/cliff_walking.rs:91-102: the start marker (byte 39) becomes '_', then '@' on the player's cell."""
    text = _put(CLIFF_MAP, 39, "_")
    return _put(text, _skip_newlines(text, player_pos), "@")


def render_blackjack(ready, dealer_cards, player_cards):
    """env/blackjack.rs:165-184: while the hand is live only the dealer's first card shows; every card is followed by a
    blank, and the dealer line ends with one more before the newline."""
    if ready:
        out = "Dealer: %d \nPlayer: " % dealer_cards[0]
    else:
        out = "Dealer: %s \nPlayer: " % "".join("%d " % c for c in dealer_cards)
    return out + "".join("%d " % c for c in player_cards)


def cards_from_words(words):
    """rand 0.8.5 `Uniform<u8>(1..11)` over a run of 32-bit stream words (env/blackjack.rs:54-56,76): widening multiply
    by 10, the 6 values whose low half exceeds 0xfffffff9 are rejected and the next word is tried."""
    cards = []
    for w in words:
        m = int(w) * 10
        if (m & 0xffffffff) <= 0xfffffff9:
            cards.append(1 + (m >> 32))
    return cards


def rust_debug_f64(x):
    """`{:?}` of an f64 for the values an episode can produce (finite, |x| < 1e16): shortest round-trip digits, always
    with a fractional part — Python's repr agrees on that range."""
    return repr(float(x))


def example_lines(label, transitions):
    """agent.rs:143-163 as a list of printed lines.  `transitions` yields, per step, (render_before, action, reward,
    terminated, render_after_if_terminated)."""
    lines, total, steps = [], 0.0, 0
    for before, action, reward, terminated, after in transitions:
        steps += 1
        lines.append(before)
        lines.append('"%s"' % label(action))                       # println!("{:?}", &str) keeps the quotes
        lines.append("step reward %s" % rust_debug_f64(reward))
        total += float(reward)
        if terminated:
            lines.append(after)
            lines.append("episode reward %s" % rust_debug_f64(total))
            lines.append("terminated with %d steps" % steps)
            break
    return lines

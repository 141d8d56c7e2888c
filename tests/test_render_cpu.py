"""`Env::render`, `Agent::example`'s transcript format and the bins' charts — host code, no
GPU.  Expected strings are written out by hand from the reference's text (env/*.rs render(), agent.rs:143-163)."""
import importlib
import math
import struct
import zlib

R = importlib.import_module("rl-rust_b200.render")
CH = importlib.import_module("rl-rust_b200.charts")


def test_taxi_render_puts_T_on_the_taxi_cell():
    # state = ((row*5 + col)*5 + pass)*4 + dest (taxi.rs:33-42); only row/col are drawn (taxi.rs:161-172)
    assert R.render_taxi(0) == "+---------+\n|T: | : :G|\n| : | : : |\n| : : : : |\n| | : | : |\n|Y| : |B: |\n+---------+"
    s = ((4 * 5 + 3) * 5 + 2) * 4 + 1
    assert R.render_taxi(s) == "+---------+\n|R: | : :G|\n| : | : : |\n| : : : : |\n| | : | : |\n|Y| : |T: |\n+---------+"
    s = ((2 * 5 + 2) * 5 + 4) * 4 + 3
    assert R.render_taxi(s).split("\n")[3] == "| : :T: : |"
    for st in range(500):                       # exactly one cell changes, always in the taxi's row, never a wall or border
        out, base = R.render_taxi(st), "\n".join(R.TAXI_MAP)
        diff = [i for i, (a, b) in enumerate(zip(out, base)) if a != b]
        row, col = st // 100, (st // 20) % 5
        assert len(out) == len(base) and diff == [12 * (row + 1) + 2 * col + 1] and out[diff[0]] == "T"


def test_frozen_lake_render():
    m4 = ("SFFF", "FHFH", "FFFH", "HFFG")
    assert R.render_frozen_lake(m4, 0) == "@FFF\nFHFH\nFFFH\nHFFG"          # 'S' is redrawn as 'F' first (frozen_lake.rs:138-140)
    assert R.render_frozen_lake(m4, 5) == "FFFF\nF@FH\nFFFH\nHFFG"
    assert R.render_frozen_lake(m4, 15) == "FFFF\nFHFH\nFFFH\nHFF@"
    m8 = ("SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG")
    assert R.render_frozen_lake(m8, 63).split("\n")[7] == "FFFHFFF@"
    assert R.render_frozen_lake(m8, 19).split("\n")[2] == "FFF@FFFF"


def test_cliff_walking_render():
    assert R.render_cliff_walking(36) == "____________\n____________\n____________\n@!!!!!!!!!!G"
    assert R.render_cliff_walking(47) == "____________\n____________\n____________\n_!!!!!!!!!!@"
    assert R.render_cliff_walking(0) == "@___________\n____________\n____________\n_!!!!!!!!!!G"
    assert R.render_cliff_walking(40) == "____________\n____________\n____________\n_!!!@!!!!!!G"


def test_blackjack_render_and_cards():
    assert R.render_blackjack(True, [5, 10], [1, 10, 3]) == "Dealer: 5 \nPlayer: 1 10 3 "          # blackjack.rs:167-168
    assert R.render_blackjack(False, [5, 10, 4], [1, 10]) == "Dealer: 5 10 4  \nPlayer: 1 10 "     # :170-175: every card + ' ', then " \n"
    # Uniform<u8>(1..11) through u32 (rand 0.8.5): card = 1 + hi32(w * 10); the 6 words with lo32 > 0xfffffff9 are rejected
    # 0x19999999 * 10 = 0xFFFFFFFA: hi 0 but the low half is in the rejection zone -> skipped, the next word is tried
    assert R.cards_from_words([0, 0x19999999, 0x1999999A, 0xFFFFFFFF, 0x80000000]) == [1, 2, 10, 6]
    cand = [((j << 32) // 10) + d for j in range(1, 10) for d in (-2, -1, 0, 1, 2)]
    rejected = [w for w in cand if R.cards_from_words([w]) == []]
    assert len(rejected) == 6 and all(((w * 10) & 0xffffffff) > 0xfffffff9 for w in rejected)   # rand 0.8.5: ints_to_reject = 6
    assert R.cards_from_words([0xE6666667]) == [10] and R.cards_from_words([0xE6666666]) == [] and R.cards_from_words([0xE6666665]) == [9]


def test_example_transcript_format():
    steps = [("view0", 1, -1.0, False, None), ("view1", 4, -10.0, False, None), ("view2", 5, 20.0, True, "last")]
    lines = R.example_lines(lambda a: ("DOWN", "UP", "RIGHT", "LEFT", "PICKUP", "DROPOFF")[a], iter(steps))
    assert lines == ["view0", '"UP"', "step reward -1.0", "view1", '"PICKUP"', "step reward -10.0", "view2", '"DROPOFF"',
                     "step reward 20.0", "last", "episode reward 9.0", "terminated with 3 steps"]
    assert R.rust_debug_f64(0) == "0.0" and R.rust_debug_f64(-100) == "-100.0" and R.rust_debug_f64(0.5) == "0.5"


def _png_size_and_pixels(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, depth, ctype = struct.unpack(">IIBB", raw[16:26])
    return w, h, depth, ctype


def test_value_range_follows_the_reference_guard():
    assert CH.value_range([[1.0, 2.0, 3.0], [0.5, 9.0]]) == (3, 0.5, 9.0)
    assert CH.value_range([[2.0, 2.0]]) == (2, -1.0, 1.0)                               # flat -> [-1, 1] (utils.rs:127-130)
    assert CH.value_range([[math.nan, math.nan], [1.0, 2.0]]) == (2, -1.0, 1.0)         # first series all NaN: the range stays NaN -> guard
    assert CH.value_range([[1.0, math.nan, 3.0], [2.0]]) == (3, 1.0, 3.0)               # f64::min / max skip NaNs inside a series


def test_charts_are_written(tmp_path):
    from PIL import Image
    res = {"legends": ["ε-Greedy One-Step Sarsa", "UCB One-Step Sarsa"]}
    for key, _ in CH.TITLES:
        res[key] = [[float(i % 7) for i in range(50)], [3.0 + math.sin(i / 5.0) for i in range(40)]]
    res["train_errors"][1][5] = math.nan
    paths = CH.plot_experiment(res, str(tmp_path), verbose=False)
    assert [p.split("/")[-1] for p in paths] == ["Train Rewards.png", "Train Episodes Length.png", "Training Error.png",
                                                 "Test Rewards.png", "Test Episodes Length.png"]
    for p in paths:
        assert _png_size_and_pixels(p)[:2] == (600, 400)                                # BitMapBackend::new(.., (600, 400))
        import numpy as np
        px = set(map(tuple, np.asarray(Image.open(p).convert("RGB")).reshape(-1, 3).tolist()))
        assert (0, 0, 255) in px and (0, 255, 0) in px and (255, 255, 255) in px       # BLUE and GREEN series on WHITE

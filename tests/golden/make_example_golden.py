"""Golden `Agent::example` transcripts (tests/golden/example_v1.json): one untrained episode per env on stream 0 of seed
0xE8A3, built from the CPU oracle driven through the reference's call order (agent.rs:143-163) and the render functions of
rl-rust_b200/render.py.  Regenerate with `python tests/golden/make_example_golden.py` (CPU only)."""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 0xE8A3
DECAY = 1.0 / 15.0
MAP_8X8 = ("SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG")
LABELS = {"taxi": ("DOWN", "UP", "RIGHT", "LEFT", "PICKUP", "DROPOFF"), "frozen_lake": ("LEFT", "DOWN", "RIGHT", "UP"),
          "cliff_walking": ("LEFT", "DOWN", "RIGHT", "UP"), "blackjack": ("HIT", "STICK")}


def transcripts():
    from oracle import oracle_py as O
    from test_gpu_example import oracle_transcript
    R = importlib.import_module("rl-rust_b200.render")
    abi = importlib.import_module("rl-rust_b200._abi")

    def hands(n0, n1):
        return R.cards_from_words(abi.rng_words(SEED, 0, n0, n1 - n0))
    cases = {"taxi": (O.ENV_TAXI, {}, lambda pos, ready: R.render_taxi(pos), None),
             "frozen_lake": (O.ENV_FROZEN_LAKE, dict(map_id=1, slippery=True), lambda pos, ready: R.render_frozen_lake(MAP_8X8, pos), None),
             "cliff_walking": (O.ENV_CLIFF_WALKING, {}, lambda pos, ready: R.render_cliff_walking(pos), None),
             "blackjack": (O.ENV_BLACKJACK, {}, lambda pos, ready, dealer, player: R.render_blackjack(ready, dealer, player), hands)}
    out = {}
    for name, (kind, extra, view, hd) in cases.items():
        cfg = O.make_config(kind, target=O.TARGET_QLEARNING, eps_decay=DECAY, seed=SEED, **extra)
        out[name] = oracle_transcript(cfg, kind, view, lambda a, name=name: LABELS[name][a], hd)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "example_v1.json"), "w") as fh:
        json.dump(transcripts(), fh, indent=0)
    print("wrote example_v1.json")

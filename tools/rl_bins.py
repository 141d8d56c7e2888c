#!/usr/bin/env python
"""The reference's bins over the B200 engine:  python tools/rl_bins.py taxi -n 1000 --n_agents 4096 --out taxi.json
(flags and defaults of src/bin/taxi.rs:22-68; see rl-rust_b200/driver.py)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if __name__ == "__main__":
    importlib.import_module("rl-rust_b200.driver").main()

"""Taxi's start-state draw (reference src/env/taxi.rs:135-142): the device's direct index form (rlb_taxi_start.h)
equals `categorical_sample`'s scan at every threshold and in between; compiled from the product's own host sources."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_direct_start_index_equals_the_scan(tmp_path):
    exe = str(tmp_path / "taxi_start_main")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "taxi_start_main.cpp"),
                           os.path.join(ROOT, "rl-rust_b200", "csrc", "rlb_host.cpp"), "-o", exe])
    out = subprocess.check_output([exe], text=True)
    assert out.strip().endswith("OK") and "tail 17" in out and "skew_direct 0" in out, out

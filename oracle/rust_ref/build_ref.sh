#!/bin/bash
# One command to pin the oracle to the REAL reference the day a Rust toolchain is at hand (none in this image):
#   oracle/rust_ref/build_ref.sh [/path/to/RL-Rust]     (default /root/reference)
# Copies the crate to a scratch directory (the reference tree is read-only and is never modified), adds src/rng.rs (the
# injected Philox stream), applies reference_rng.patch (the 9 thread_rng() call sites + `pub mod rng`), adds the
# parity_dump bin, builds --release (README.md:29: the slippery FrozenLake needs wrapping arithmetic) and leaves the
# binary in oracle/_ref/parity_dump, where tests/test_rust_ref.py picks it up.  Needs network or a vendored ~/.cargo
# for the crate's dependencies (rand 0.8.5, fxhash 0.2.1, ...).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
command -v cargo >/dev/null || { echo "cargo not found: no Rust toolchain on this machine" >&2; exit 3; }
W="$(mktemp -d)"
cp -r "$REF"/. "$W"/
cp "$HERE/rng.rs" "$W/src/rng.rs"
cp "$HERE/parity_dump.rs" "$W/src/bin/parity_dump.rs"
(cd "$W" && patch -p1 < "$HERE/reference_rng.patch")
(cd "$W" && cargo test --release --lib rng::tests && cargo build --release --bin parity_dump)
mkdir -p "$HERE/../_ref"
cp "$W/target/release/parity_dump" "$HERE/../_ref/parity_dump"
echo "built $HERE/../_ref/parity_dump"

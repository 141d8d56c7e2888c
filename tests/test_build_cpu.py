"""Build hygiene of the bench kernels, from the ptxas logs the Makefile keeps (rl-rust_b200/build/*.ptxas.log): a source
change that silently pushes a register array into local memory (a loop the unroller gives up on) shows up here as
kilobytes of spill code — it once cost the Taxi kernel 3.7x before any GPU saw it."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (mangled-name fragment, max registers, max spill-store bytes, max stack-frame bytes)
KERNELS = {
    "C4 k_run<TAXI,f32,Basic,eps,one-step,HBM>": ("k_runILi3EfLi0ELi0ELb0ELi1ELb0E", 64, 400, 200),
    "C3 k_run<CLIFF,f32,Double,UCB,one-step,HBM>": ("k_runILi2EfLi1ELi1ELb0ELi1ELb0E", 96, 300, 120),   # 5 CTAs/SM
    "C1 k_run<BLACKJACK,f32,Basic,eps,one-step,HBM>": ("k_runILi0EfLi0ELi0ELb0ELi1ELb0E", 64, 400, 200),
    "C2 k_run<FROZEN_LAKE,f32,Basic,eps,traces,hybrid>": ("k_runILi1EfLi0ELi0ELb1ELi3ELb0E", 255, 0, 0),
    # C5's slowest cells: Taxi trace agents on the lazy store (eps-greedy: no launch bound; UCB: 4 CTAs/SM = 128 registers)
    "C5 k_run<TAXI,f32,Basic,eps,traces,HBM-lazy>": ("k_runILi3EfLi0ELi0ELb1ELi4ELb0E", 168, 0, 0),
    "C5 k_run<TAXI,f32,Double,UCB,traces,HBM-lazy>": ("k_runILi3EfLi1ELi1ELb1ELi4ELb0E", 128, 300, 128),
}


@pytest.mark.parametrize("name", sorted(KERNELS))
def test_bench_kernels_stay_in_registers(name):
    frag, max_regs, max_spill, max_stack = KERNELS[name]
    logs = glob.glob(os.path.join(ROOT, "rl-rust_b200", "build", "*.ptxas.log"))
    if not logs:
        import __graft_entry__
        __graft_entry__.build()
        logs = glob.glob(os.path.join(ROOT, "rl-rust_b200", "build", "*.ptxas.log"))
    found = None
    for f in logs:
        txt = open(f).read()
        i = txt.find("Function properties for _ZN3rlb5" + frag)
        if i >= 0:
            found = txt[i:i + 600]
            break
    assert found, "no ptxas record for %s" % name
    stack = int(re.search(r"(\d+) bytes stack frame", found).group(1))
    spill = int(re.search(r"(\d+) bytes spill stores", found).group(1))
    regs = int(re.search(r"Used (\d+) registers", found).group(1))
    assert regs <= max_regs and spill <= max_spill and stack <= max_stack, (name, regs, spill, stack)

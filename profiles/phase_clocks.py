"""Cycle timeline of ONE thread's step at the bench shape (65,536 envs x 8 OU pairs, L2-cold slab rotation), from a
-DMDG_PHASE_CLOCKS build of the library:
    MDG_LIB_VARIANT=clk MDG_EXTRA_NVCC_FLAGS="-DMDG_PHASE_CLOCKS" python -m madigan_b200.build     (here)
    MDG_LIB_VARIANT=clk python profiles/phase_clocks.py [hold]                                     (on the GPU box)
Slots: 0 entry, 1 prologue loads issued, 2 normals generated, per pair p: 3+4p state/units in registers, 4+4p / 5+4p
after the two transactions, 6+4p after the pair's tick + stores; 39 loop done, 40 tail done."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from madigan_b200 import _lib

dev = torch.device("cuda", 0)
hold = len(sys.argv) > 1 and sys.argv[1] == "hold"
n = 65536
slabs = int(os.environ.get("PCLK_SLABS", "8"))      # 1: the slab's pages stay in the TLBs (and its state in L2)
flush = int(os.environ.get("PCLK_FLUSH", "0"))      # 1: write 256 MB between launches (L2-cold, TLB mostly warm)
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
acts = bench.synth_actions(4, n, 1, device=dev)
for i in range(40 * slabs):
    envs[i % slabs].step(acts[i % 4], auto_reset=True)
torch.cuda.synchronize()
L = _lib.lib()
L.mdg_debug_phase_clocks.argtypes = [C.c_void_p]
rows = []
for i in range(3 * slabs):
    env = envs[i % slabs]
    if flush:
        scratch.fill_(i & 255)
    env.step(acts[i % 4] * 0. if hold else acts[i % 4])
    torch.cuda.synchronize()
    buf = np.zeros((64, 64), np.int64)
    assert L.mdg_debug_phase_clocks(buf.ctypes.data_as(C.c_void_p)) == 0
    rows.append(buf - buf[:, :1])
    env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
r = np.median(np.concatenate(rows[slabs:]), axis=0)  # median over blocks and launches, cycles since entry
print("mode", "hold" if hold else "multi", "slabs", slabs, "flush", flush)
if r[2] < r[1]:  # final kernel: the normals are generated before the prologue's loaded values are first used
    print(f"tick load + ring set-up + normals {r[2]:.0f}  rest of the prologue {r[1] - r[2]:.0f}  loop {r[39] - r[1]:.0f}  tail {r[40] - r[39]:.0f}  total {r[40]:.0f} cycles")
    r[2] = r[1]
else:
    print(f"prologue {r[1]:.0f}  normals {r[2] - r[1]:.0f}  loop {r[39] - r[2]:.0f}  tail {r[40] - r[39]:.0f}  total {r[40]:.0f} cycles")
if r[41] > 0:
    print(f"  pair 0: data landed {r[41] - r[2]:.0f} after the normals, barrier {r[42] - r[41]:.0f};  pair 3: mbarrier wait {r[43] - r[14]:.0f}, barrier {r[44] - r[43]:.0f}")
for p in range(8):
    b = 3 + 4 * p
    prev = r[2] if p == 0 else r[b - 1]
    print(f"  pair {p}: wait/load {r[b] - prev:6.0f}  tx0 {r[b + 1] - r[b]:6.0f}  tx1 {r[b + 2] - r[b + 1]:6.0f}  tick+stores {r[b + 3] - r[b + 2]:6.0f}")

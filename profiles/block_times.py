"""Start / end time (%globaltimer) and SM of every block of ONE step-kernel launch at the bench shape, from a
-DMDG_PHASE_CLOCKS build (see profiles/phase_clocks.py): launch ramp, per-block duration by blocks-per-SM, drain."""
import collections
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from madigan_b200 import _lib

dev = torch.device("cuda", 0)
n, slabs = 65536, 8
envs = [bench.make_env(dev, s * n, n_envs=n) for s in range(slabs)]
acts = bench.synth_actions(4, n, 1, device=dev)
for i in range(40 * slabs):
    envs[i % slabs].step(acts[i % 4], auto_reset=True)
torch.cuda.synchronize()
L = _lib.lib()
L.mdg_debug_block_times.argtypes = [C.c_void_p]
for i in range(2 * slabs):
    env = envs[i % slabs]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(acts[i % 4]); b.record()
    torch.cuda.synchronize()
    buf = np.zeros((1024, 4), np.uint64)
    assert L.mdg_debug_block_times(buf.ctypes.data_as(C.c_void_p)) == 0
    env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
    if i < slabs:
        continue
    nblk = int((buf[:, 0] > 0).sum())
    t = buf[:nblk].astype(np.int64)
    t0 = t[:, 0].min()
    start, end, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
    per_sm = collections.Counter(sm.tolist())
    nb = np.array([per_sm[s] for s in sm])
    dur = end - start
    print(f"launch {i}: events {a.elapsed_time(b) * 1e3:.1f} us | first start 0, last start {start.max():.1f} us, first end {end.min():.1f}, "
          f"last end {end.max():.1f} us | SMs used {len(per_sm)}, blocks/SM {sorted(collections.Counter(per_sm.values()).items())}")
    bsz = n // nblk
    done = env.t["done"].view(torch.uint8).cpu().numpy().reshape(nblk, bsz).sum(1)
    risk = (env.t["risk"].cpu().numpy() != 0).sum(0).reshape(nblk, bsz).sum(1)
    tu = (env.t["trans_units"].cpu().numpy() != 0).sum(0).reshape(nblk, bsz).sum(1)
    pct = np.percentile(dur, [50, 90, 99, 100])
    print(f"    block duration percentiles 50/90/99/100: {pct.round(1)} us; corr(duration, #done) {np.corrcoef(dur, done)[0, 1]:.2f}, "
          f"corr(duration, #rejected orders) {np.corrcoef(dur, risk)[0, 1]:.2f}, corr(duration, #executed orders) {np.corrcoef(dur, tu)[0, 1]:.2f}")
    top = np.argsort(-dur)[:8]
    print("    slowest blocks:", [(int(b_), round(float(dur[b_]), 1), int(sm[b_]), int(done[b_]), int(risk[b_]), int(tu[b_])) for b_ in top], "(block, us, sm, done, rejected, executed)")
    for k in sorted(set(nb.tolist())):
        m = nb == k
        print(f"    blocks on SMs with {k} blocks: n={m.sum():3d}  duration median {np.median(dur[m]):.1f} us  max {dur[m].max():.1f}  end median {np.median(end[m]):.1f}")

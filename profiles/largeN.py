"""Step-kernel throughput as a function of envs per launch (one Env, device-resident actions, no resets in the
event pairs): shows the latency-bound regime at 65,536 envs (14 warps/SM) vs the throughput-bound one."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
for n in (16_384, 65_536, 262_144, 1_048_576):
    env = bench.make_env(dev, 0, n_envs=n)
    acts = bench.synth_actions(2, n, 1, device=dev)
    for i in range(5):
        env.step(acts[i % 2], auto_reset=True)
    evs = []
    for i in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[i % 2]); b.record()
        env._reset_launch(env.t["done"], bench.WINDOW, True, None, None)
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in evs)[len(evs) // 2]
    gbs = bench.bytes_per_env_step() * n / (ms * 1e-3) / 1e9
    print(f"N={n:8d}  step kernel {ms*1e3:8.1f} us  {n/(ms*1e-3):.3e} env-steps/s  {gbs:7.1f} GB/s algorithmic  {gbs/6549.4*100:5.1f}% of measured HBM peak")
    del env, acts
    torch.cuda.empty_cache()

"""Device time of the window materialiser at the bench shape (65,536 envs x 16 features x 64 rows)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
env = bench.make_env(dev, 0)
acts = bench.synth_actions(4, bench.ENVS_PER_GPU, 1, device=dev)
for i in range(80): env.step(acts[i % 4], auto_reset=True)
out32 = torch.empty((env.N, env.k, env.nA), dtype=torch.float32, device=dev)
out64 = torch.empty((env.N, env.k, env.nA), dtype=torch.float64, device=dev)
def timed(f, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f(); torch.cuda.synchronize(); a.record()
    for i in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
alg = env.N * env.k * env.nA * (8 + 4)
for norm in (None, "lookback", "lookback_log", "log", "standard_normal", "log_standard_normal", "expanding"):
    t32 = timed(lambda: env.window(norm, out=out32)); t64 = timed(lambda: env.window(norm, out=out64))
    print(f"{str(norm):20s} fp32 out {t32:8.1f} us ({alg / t32 / 1e3:7.1f} GB/s algorithmic)   fp64 out {t64:8.1f} us")
print("portfolio_window %.1f us, time_window %.1f us" % (timed(lambda: env.portfolio_window()), timed(lambda: env.time_window())))

"""Step throughput of the other BASELINE configurations (C2, C3, C4) on one B200: device time per step (step kernel
+ masked auto-reset with history fill, one stream, device-resident units), env-steps/s and algorithmic GB/s."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madigan_b200.environments import Env
dev = torch.device("cuda", 0)

def bytes_per_env_step(nA, G, R, ra, sh=1):
    return 8 * (14 * nA + 2 * G + 2 * R + 7 + ra + sh) + nA + 2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import COMPOSITE16  # Synth4 + OU4 + OUPair + SimpleTrend2 + TrendOU2 + TrendyOU2

CONFIGS = {
    "C2 4,096 x OU(1), log-return": dict(ds=("OU", {"mean": [10.], "theta": [.08], "phi": [.04]}), N=4096, nA=1, G=0,
        reward=dict(reward_shaper_config={"reward_shaper": None}, nstep_return=1, reduce_rewards=True), R=0,
        margins=(1., .25), costs=(0., 0.), unit=5000.),
    "C3 65,536 x OUPair, cost+slippage, DSR n=1": dict(ds=("OUPair", {"theta": .015, "phi": .01, "noise": .03}), N=65536,
        nA=2, G=1, reward=dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=1,
        reduce_rewards=True), R=2, margins=(1., .25), costs=(.02, .001), unit=5000.),
    "C3 65,536 x OUPair, cost+slippage, DDR n=20": dict(ds=("OUPair", {"theta": .015, "phi": .01, "noise": .03}), N=65536,
        nA=2, G=1, reward=dict(reward_shaper_config={"reward_shaper": "DDR", "adaptation_rate": .001}, nstep_return=20,
        reduce_rewards=True), R=2, margins=(1., .25), costs=(.02, .001), unit=5000.),
    "C4 65,536 x composite16, margin .1, PPC n=5": dict(ds=("Composite", COMPOSITE16), N=65536, nA=16, G=4 + 1 + 4 + 6 + 8,
        reward=dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1.] + [0.] * 16,
        "cosine_temp": .025}, nstep_return=5, reduce_rewards=False), R=0, margins=(.1, .25), costs=(.001, 0.), unit=5000.),
}
for name, c in CONFIGS.items():
    env = Env(c["ds"][0], 1e6, {"data_source_config": c["ds"][1]}, n_envs=c["N"], window=64, seed=5, device=dev,
              reward=c["reward"])
    env.setRequiredMargin(c["margins"][0]); env.setMaintenanceMargin(c["margins"][1])
    env.setTransactionCost(c["costs"][0], 0.); env.setSlippage(c["costs"][1], 0.)
    env.reset(fill_history=True)
    g = torch.Generator().manual_seed(1)
    acts = [(torch.randint(-1, 2, (c["N"], c["nA"]), generator=g).double() * c["unit"]).to(dev) for _ in range(4)]
    for i in range(60): env.step(acts[i % 4], auto_reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    n = 300
    for i in range(n): env.step(acts[i % 4], auto_reset=True)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    ra = 1 if c["reward"]["reduce_rewards"] else c["nA"]
    B = bytes_per_env_step(c["nA"], c["G"], c["R"], ra)
    w = 0.
    if name.startswith("C4"):
        out = torch.empty((c["N"], 64, 16), dtype=torch.float32, device=dev)
        a.record()
        for i in range(20): env.window("standard_normal", out=out)
        b.record(); torch.cuda.synchronize(); w = a.elapsed_time(b) / 20 * 1e3
    print(f"{name:48s} {us:8.1f} us/step  {c['N'] / us * 1e6:.3e} env-steps/s  {B} B/env-step -> {B * c['N'] / us / 1e3:7.1f} GB/s"
          + (f"   + standard_normal fp32 window {w:.0f} us" if w else ""))
    del env, acts
    torch.cuda.empty_cache()

import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_gpu_parity as G
N = int(sys.argv[1])
reward = dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=1, discount=.99, reduce_rewards=True)
env, orc, P = G.make_pair("pairs8", N, 8, reward, (1., .25), (.02, 0., .001, 0.))
rng = np.random.default_rng(1)
orc.reset(fill_ticks=1, clear_nstep=False)
G.sync_state_from_oracle(env, orc)
nz, uz = G.noise(rng, P, N, ticks=8)
env.reset(fill_history=True, normals=nz, uniforms=uz)
orc.reset(fill_ticks=8, normals=nz, uniforms=uz)
for t in range(12):
    units = (rng.integers(-1, 2, size=(N, 16)) * 5000.).astype(np.float64)
    nz, uz = G.noise(rng, P, N)
    env.step(torch.from_numpy(units), normals=nz, uniforms=uz)
    orc.step(units, normals=nz, uniforms=uz)
    G.compare_step(env, orc, True, t, shaped=True)
    d = orc.done.astype(bool)
    print(t, "done rate", d.mean(), float(env.t["done"].float().mean()))
    if d.any():
        nz, uz = G.noise(rng, P, N, ticks=8)
        env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
        orc.reset(mask=orc.done.copy(), fill_ticks=8, normals=nz, uniforms=uz)
print("OK", N)

"""Host-side cost of enqueuing one step (no device wait): must stay below the device time per step."""
import sys, time, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device('cuda', 0)
env = bench.make_env(dev, 0)
acts = bench.synth_actions(4, bench.ENVS_PER_GPU, 1, device=dev)
for i in range(50): env.step(acts[i % 4], auto_reset=True)
torch.cuda.synchronize()
def t(f, n=150):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): f(i)
    dt = (time.perf_counter() - t0) / n * 1e6; torch.cuda.synchronize(); return dt
print("step only                       host us: %.1f" % t(lambda i: env.step(acts[i % 4])))
print("reset only (_reset_launch)      host us: %.1f" % t(lambda i: env._reset_launch(env.t["done"], 64, True, None, None)))
print("step(auto_reset=True), one call host us: %.1f" % t(lambda i: env.step(acts[i % 4], auto_reset=True)))
import ctypes as C
from madigan_b200.environments.env import _reset_workspace
lib = env._lib
L = env._launch(1, 0)
env._IO.units = acts[0].data_ptr()
ws = _reset_workspace(lib, env.P, env.N, 64, env.device)
print("raw mdg_step ctypes             host us: %.1f" % t(lambda i: lib.mdg_step(env._pP, env._pR, env._pS, env._pIO, env._pL)))
print("raw mdg_reset_ws ctypes         host us: %.1f" % t(lambda i: lib.mdg_reset_ws(env._pP, env._pS, env._pIO, env._pL, env._done_u8.data_ptr(), 64, 1, ws.data_ptr(), ws.numel())))
print("raw mdg_step_autoreset ctypes   host us: %.1f" % t(lambda i: lib.mdg_step_autoreset(env._pP, env._pR, env._pS, env._pIO, env._pL, 64, 1, ws.data_ptr(), ws.numel())))
x = torch.zeros(8, device=dev)
print("torch tiny kernel (x.add_(1))   host us: %.1f" % t(lambda i: x.add_(1)))

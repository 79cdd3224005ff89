"""Per-step CUDA-event timings of the env-step kernel over a long run + NVML clock/power samples.
Explains how the per-launch time evolves (episode maturity, power/clock behaviour).  Tuning tool, not a bench."""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pynvml
from ml4ca_b200.env import RevoltFinal, StandInHull

n = int(os.environ.get("N", 1 << 24))
steps = int(os.environ.get("STEPS", 400))
dev = torch.device("cuda", 0)
env = RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=dev, seed=2, auto_reset=True)
env.reset(fraction=0.8)
SHIFT = 4 * 1031
room = SHIFT * (steps // 2 + 2)
flat = [torch.rand(7 * n + room, device=dev) * 2 - 1 for _ in range(2)]
if os.environ.get("ZERO_ACT"):
    for p in flat: p.zero_()
def actions(i):
    off = SHIFT * (i // 2)
    return flat[i % 2][off:off + 7 * n].view(7, n)
out = (torch.empty(9, n, device=dev), torch.empty(n, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = threading.Event()
def poll():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.005)
th = threading.Thread(target=poll, daemon=True); th.start()
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
t0 = time.perf_counter()
evs[0].record()
dones = []
for i in range(steps):
    env.step_into(actions(i), *out)
    evs[i + 1].record()
    if i % 50 == 0: dones.append(out[2].ne(0).float().mean())
torch.cuda.synchronize()
t1 = time.perf_counter()
stop.set(); th.join()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
print("wall %.1f ms for %d steps" % ((t1 - t0) * 1e3, steps))
print("ms/step at", {i: round(ms[i], 3) for i in list(range(0, 12)) + list(range(12, steps, max(1, steps // 40)))})
print("done fraction", [round(float(d), 4) for d in dones])
ins = [s for s in samples if t0 <= s[0] <= t1]
print("nvml", [(round((s[0] - t0) * 1e3), s[1], round(s[2]), hex(s[3])) for s in ins[:: max(1, len(ins) // 30)]])

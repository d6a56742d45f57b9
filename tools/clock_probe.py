"""SM clock / power while one kernel family runs in a loop: python tools/clock_probe.py k1|fused|k3|ub"""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml
from spectrobot_b200 import engine, synthetic as S
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.05)
mode = sys.argv[1] if len(sys.argv) > 1 else "k1"
w0, w1, n_lev = 2850.0, 3450.0, 12
g = S.spectral_grid(w0, w1)
lines = S.line_table(30000, w0, w1, n_levels=n_lev)
ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
out = torch.empty((8, n_lev, 3, len(g)), dtype=torch.float64, device="cuda")
pts = [[0.05 * (3 + j), 150.0 + 2.0 * j] for j in range(8)]
if mode == "k1":
    fn = lambda: ls.gcoeff_cells(pts, out=out, check_status=False)
    unit = 8 * ls.n_active * 13010
else:
    fn = lambda: engine.fp64_peak(40000)
    unit = 0
fn(); torch.cuda.synchronize()
t = threading.Thread(target=sampler, daemon=True); t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
t_end = time.time() + 4.0
while time.time() < t_end:
    e0.record(); 
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
stop = True; t.join()
clk = np.array([s[0] for s in samples]); pw = np.array([s[1] for s in samples])
print(mode, "ms/call first %.3f median %.3f last %.3f" % (res[0], np.median(res), res[-1]),
      ("-> %.3e evals/s" % (unit / (np.median(res) * 1e-3))) if unit else "")
print("clock MHz min %d median %d max %d; power W median %.0f max %.0f; reasons %s" %
      (clk.min(), np.median(clk), clk.max(), np.median(pw), pw.max(), sorted(set(hex(s[2]) for s in samples))))

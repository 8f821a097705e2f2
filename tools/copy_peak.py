"""Burst vs sustained HBM copy bandwidth on this GPU (context for the roofline denominator: the
driver's MEASURED_PEAKS.json hbm_gbs is a best-of-10 BURST of the same torch copy)."""
import json, subprocess, time, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
a.normal_()
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = ev(), ev(); e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out = {"burst_gbs": 2 * n * 2 / best / 1e6}
for reps in (100, 400, 1200):
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    time.sleep(0.05); smi.terminate(); lines = smi.communicate()[0].strip().splitlines()
    ms = e0.elapsed_time(e1)
    out[f"sustained_{reps}_gbs"] = 2 * n * 2 * reps / ms / 1e6
    out[f"sustained_{reps}_s"] = ms / 1e3
    out[f"smi_{reps}"] = lines[-3:]
print(json.dumps(out))

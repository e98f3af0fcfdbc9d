#!/usr/bin/env python
"""Measured int8 tensor peak of this B200: a plain dense int8 x int8 -> int32
GEMM by the vendor library (torch._int_mm -> cuBLASLt, tcgen05.mma.kind::i8 on sm_100), timed exactly as
MEASURED_PEAKS.json times its bf16 GEMM: 8192^3, best of 10 (burst) and back to back for 4 s (sustained),
CUDA events, nvidia-smi clocks sampled under load.  Writes profiles/int8_peak.json, the denominator of every
tensor-bound roofline bench.py prints.
usage: python tools/i8_peak.py [--out profiles/int8_peak.json]"""
import argparse
import json
import statistics
import subprocess
import threading
import time
from pathlib import Path

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=str(Path(__file__).resolve().parent.parent / "profiles" / "int8_peak.json"))
ap.add_argument("--n", type=int, default=8192)
args = ap.parse_args()
dev = torch.device("cuda", 0)
n = args.n
a = torch.randint(-5, 6, (n, n), dtype=torch.int8, device=dev)
b = torch.randint(-5, 6, (n, n), dtype=torch.int8, device=dev).t()      # column-major B, as cuBLASLt int8 wants
ops = 2.0 * n * n * n
for _ in range(5):
    torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100",
                         "-i", "0"], stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [rows.append(l.split(",")) for l in proc.stdout], daemon=True)
th.start()
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0
t0 = time.time()
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        torch._int_mm(a, b)
    reps += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained_ms = e0.elapsed_time(e1) / reps
proc.terminate()
clk = [float(r[0]) for r in rows if len(r) >= 2 and r[0].strip().replace(".", "").isdigit()]
pw = [float(r[1]) for r in rows if len(r) >= 2]
busy = [c for c in clk if c >= 0.5 * max(clk)] if clk else []
out = {"tops": ops / (best / 1e3) / 1e12, "tops_sustained": ops / (sustained_ms / 1e3) / 1e12,
       "sm_mhz_median_under_load": statistics.median(busy) if busy else None,
       "power_w_max": max(pw) if pw else None, "gpu_name": torch.cuda.get_device_name(0),
       "how": f"torch._int_mm (cuBLASLt int8 -> int32) {n}^3: best of 10 (tops) and back to back for 4 s "
              f"(tops_sustained), CUDA events; tools/i8_peak.py",
       "torch": torch.__version__}
Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))

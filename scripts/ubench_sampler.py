"""Does sampling the clocks disturb a short timed region?  One GPU, PageRank at RMAT-22 in steps of 10 iterations (~2 ms per
step, one host synchronisation per step — the host cannot run ahead, like a multi-GPU step of a few ms): step-time mean /
p99 / max without a sampler, with `nvidia-smi -lms 100` running (bench.py's sampler), with an in-process NVML thread
(clock + event reasons only) every 20 ms, and the latency of the individual NVML queries.  No torch.
usage: python scripts/ubench_sampler.py  -> JSON lines (gpurun_out/ubench_sampler.txt)"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "ubench_sampler.txt")
os.makedirs(os.path.dirname(out_path), exist_ok=True)


def emit(obj):
    line = json.dumps(obj)
    print(line, flush=True)
    with open(out_path, "a") as f:
        f.write(line + "\n")


scale = 22
g = L.LuxGraph.from_rmat(scale, 1 << scale, 16 << scale, scale)
g.init()
g.iterate(50)


def steps(seconds):
    ts, t_end = [], time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        t0 = time.perf_counter()
        g.iterate(10)
        ts.append(1e3 * (time.perf_counter() - t0))
    a = np.array(ts)
    return dict(steps=len(a), mean_ms=float(a.mean()), p50_ms=float(np.percentile(a, 50)), p99_ms=float(np.percentile(a, 99)),
                max_ms=float(a.max()), over_2x_median=int((a > 2 * np.median(a)).sum()))


emit(dict(case="no sampler", **steps(2.5)))

Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
t0 = time.perf_counter()
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
emit(dict(case="nvidia-smi start-up (first 1.5 s after Popen)", **steps(1.5)))
emit(dict(case="nvidia-smi -lms 100 running", **steps(2.5)))
proc.terminate()
proc.wait(timeout=5)
emit(dict(case="no sampler again", **steps(1.5)))

import pynvml  # noqa: E402
t0 = time.perf_counter()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
emit(dict(case="nvmlInit + handle", ms=1e3 * (time.perf_counter() - t0)))


def lat(fn, n=20):
    v = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        v.append(1e3 * (time.perf_counter() - t))
    return dict(mean_ms=float(np.mean(v)), max_ms=float(np.max(v)))


reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
emit(dict(case="NVML query latency", clock_sm=lat(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
          max_clock_sm=lat(lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)),
          reasons=lat(lambda: reasons(h)), power=lat(lambda: pynvml.nvmlDeviceGetPowerUsage(h))))

stop = False
samples = []


def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), reasons(h)))
        time.sleep(0.02)


th = threading.Thread(target=poll, daemon=True)
th.start()
emit(dict(case="NVML thread: clock + reasons every 20 ms", **steps(2.5), samples=len(samples)))
stop = True
th.join()
emit(dict(case="no sampler, end", **steps(1.0)))
g.close()

"""Host RNG: time of legacy_randn / legacy_randn_f32 per size with the threaded path on and off (where do the threads
start to pay on this host?)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dbn_b200
from dbn_b200 import _hostrng as hr
print("cpus", os.cpu_count(), "threads", hr._THREADS)
for n in (65536, 131072, 250000, 392000, 524288, 1000000, 4000000):
    row = []
    for par_min in (1 << 40, 1):
        hr._PAR_MIN = par_min
        ts = []
        for _ in range(7):
            np.random.seed(1)
            t = time.perf_counter(); a = hr.legacy_randn(n); ts.append((time.perf_counter() - t) * 1e3)
        out = np.empty(n, dtype=np.float32)
        t2 = []
        for _ in range(7):
            np.random.seed(1)
            t = time.perf_counter(); hr.legacy_randn_f32(out, 3.0); t2.append((time.perf_counter() - t) * 1e3)
        row.append((min(ts), np.median(ts), min(t2), np.median(t2)))
    np.random.seed(1); t = time.perf_counter(); b = np.random.randn(n); tn = (time.perf_counter() - t) * 1e3
    print("n=%8d  sequential f64 %.2f (med %.2f) f32 %.2f (med %.2f) | threaded f64 %.2f (med %.2f) f32 %.2f (med %.2f) | numpy %.2f ms"
          % ((n,) + row[0] + row[1] + (tn,)), flush=True)

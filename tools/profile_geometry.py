"""One launch each of K5 (kabsch_batched_kernel, 1 M problems of 11 points) and K6 (discrepancy_all_vs_all_kernel,
N = 20 000, n = 7) for ncu, with their CUDA-event times and the measured DFMA peak.
   ncu --set full --clock-control none --import-source on -k regex:"kabsch|discrepancy_all" -c 2 -o out python tools/profile_geometry.py"""
import ctypes as C
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fr3d_python_b200 import _lib, synthetic
from fr3d_python_b200.engine import Engine
from fr3d_python_b200.geometry import discrepancy as disc

eng = Engine.get()
dev = eng.device
with gzip.open(os.path.join(ROOT, "tests", "golden", "geometry.json.gz"), "rt") as f:
    av = json.load(f)["all_vs_all"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
Cn, Rn = synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 5000)
n = Cn.shape[1]
dc = torch.from_numpy(Cn).to(dev)
dr = torch.from_numpy(Rn.reshape(N, n, 9)).to(dev)
D = torch.empty((N, N), dtype=torch.float64, device=dev)
P, k = 1000000, 11
rng = np.random.default_rng(1)
s1 = torch.from_numpy(rng.normal(size=(P * k, 3))).to(dev)
s2 = torch.from_numpy(rng.normal(size=(P * k, 3))).to(dev)
off = torch.arange(0, (P + 1) * k, k, dtype=torch.int32, device=dev)
U = torch.empty((P, 9), dtype=torch.float64, device=dev)
m1 = torch.empty((P, 3), dtype=torch.float64, device=dev)
rm = torch.empty(P, dtype=torch.float64, device=dev)
new1 = torch.empty((P * k, 3), dtype=torch.float64, device=dev)


def k5():
    _lib.check(eng.lib.fr3d_kabsch_batched(s1.data_ptr(), s2.data_ptr(), None, off.data_ptr(), P, U.data_ptr(), m1.data_ptr(),
                                           None, rm.data_ptr(), None, new1.data_ptr(), eng.stream_ptr()), "kabsch")


def k6():
    _lib.check(eng.lib.fr3d_discrepancy_all_vs_all(dc.data_ptr(), dr.data_ptr(), None, N, n, 1.0, 0, N, D.data_ptr(),
                                                   eng.stream_ptr()), "all_vs_all")


out = {}
for name, fn, reps in (("kabsch_batched_kernel", k5, 5), ("discrepancy_all_vs_all_kernel", k6, 3)):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / reps
out["fp64_peak_tflops_measured"] = disc.fp64_peak_tflops()
out["k5_problems"] = P
out["k5_points"] = k
# K5 model: 2*9k (cov) + ~600 (rotation) + 2*9k (fitted) + 8k (sse) flops per problem
out["k5_model_tflops"] = P * (18 * k + 600 + 18 * k + 8 * k) / (out["kabsch_batched_kernel_ms"] * 1e-3) / 1e12
out["k5_bytes_gbs"] = P * (2 * 24 * k + 24 * k + 72 + 24 + 8) / (out["kabsch_batched_kernel_ms"] * 1e-3) / 1e9
out["k6_pairs"] = N * (N - 1) // 2
out["k6_model_tflops"] = out["k6_pairs"] * (154 * n + 600) / (out["discrepancy_all_vs_all_kernel_ms"] * 1e-3) / 1e12
print(json.dumps(out))

#!/usr/bin/env python
"""Streaming-family roofline at a given code size / schedule / batch (device-resident, CUDA-event kernel time).
Usage: python scripts/stream_frac.py n schedule batch [p] [max_iter] [kernel]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_b200 import BpDecoder, _capi, codes  # noqa: E402

n = int(sys.argv[1])
schedule = sys.argv[2]
B = int(sys.argv[3])
p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
max_iter = int(sys.argv[5]) if len(sys.argv) > 5 else (100 if schedule == "serial" else 50)
kernel = sys.argv[6] if len(sys.argv) > 6 else "stream"
H = codes.regular_ldpc(n, 3, 6, seed=1)
m, E = H.shape[0], int(H.nnz)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(99)
Hd = torch.tensor(H.toarray(), dtype=torch.float16, device=dev)
d_syn = torch.empty((B, m), dtype=torch.uint8, device=dev)
chunk = max(1, (1 << 27) // n)
for lo in range(0, B, chunk):
    hi = min(B, lo + chunk)
    e = (torch.rand((hi - lo, n), device=dev, generator=gen) < p).to(torch.float16)
    d_syn[lo:hi] = (e @ Hd.T).to(torch.int32).remainder_(2).to(torch.uint8)
del Hd
d = BpDecoder(H, error_rate=p, max_iter=max_iter, bp_method="ms", schedule=schedule, ms_scaling_factor=0.625,
              input_vector_type="syndrome", kernel=kernel)
h = d._ensure_handle()
L = _capi.lib()
d_dec = torch.empty((B, n), dtype=torch.uint8, device=dev)
d_conv = torch.empty(B, dtype=torch.uint8, device=dev)
d_its = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(dev)


def step():
    rc = L.bpb_decode_batch_device(h, 0, C.c_void_p(d_syn.data_ptr()), B, C.c_void_p(d_dec.data_ptr()),
                                   C.c_void_p(d_conv.data_ptr()), C.c_void_p(d_its.data_ptr()), None,
                                   C.c_void_p(st.cuda_stream))
    _capi.check(h, rc)


step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
step()
e1.record(st)
torch.cuda.synchronize()
info = d.info()
per_iter = (6 * E * 8 if schedule == "serial" else 4 * E * 8) + n + m
its = int(d_its.sum(dtype=torch.int64).item())
kit = info["stream_iterations"] if info["kernel_family"] == 1 else its
alg = kit * per_iter + B * (m + n + 5)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
print(json.dumps({"n": n, "schedule": schedule, "batch": B, "p": p, "max_iter": max_iter, "family": info["kernel_family"],
                  "step_ms": e0.elapsed_time(e1), "kernel_ms": info["last_kernel_ms"], "decodes_per_s": B / (e0.elapsed_time(e1) * 1e-3),
                  "mean_it": its / B, "conv": float(d_conv.float().mean().item()), "handed_off": info["stream_handed_off"],
                  "alg_GBps": alg / info["last_kernel_ms"] / 1e6, "frac_of_measured_hbm": alg / info["last_kernel_ms"] / 1e6 / peak,
                  "grid": info["grid"], "block": info["block"]}))

#!/usr/bin/env python
"""bench.py -- syndrome decodes/sec of the batched BP decoder on BASELINE.json's headline configuration.

Workload (BASELINE.json configs[1], the one the metric is quoted on): (3,6)-regular LDPC, n=1000, m=500,
min-sum, parallel schedule, max_iter=50, ms_scaling_factor=0.625, one batch of 2^20 synthetic BSC(p=0.05)
syndromes per step.  A "step" = one pass of the hot path over that batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--kernel auto|stream|smem]

* ours      : `value` = decodes/s with the syndromes already resident in HBM (CUDA events on the launching
              stream, max over ranks); `e2e` = the same through the host-buffer C-ABI call (pinned host memory,
              H2D + kernels + D2H inside the timed region); `roofline` for the message-update kernel against the
              measured HBM peak (MEASURED_PEAKS.json); `cpu_baseline` = the reference's own C++ timed on this box.
* reference : the unmodified reference C++ (oracle/_ref) on all host threads, on a bounded sample of the same
              workload per step.

N > 1 is launched by torchrun (one rank per GPU); the batch shards by rank with no collective on the data path
("weak" scaling: every rank decodes its own 2^20 syndromes), only the timing is reduced (max over ranks).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CODE, DV, DC, CODE_SEED = 1000, 3, 6, 1
P_ERR, MAX_ITER, MS_SCALING = 0.05, 50, 0.625
METRIC = "syndrome decodes/sec (50 BP iters) on n=1000 (3,6)-LDPC"


def workload_name(batch):
    return (f"(3,6)-regular LDPC n={N_CODE} (seed {CODE_SEED}), min_sum parallel max_iter={MAX_ITER} "
            f"ms_scaling={MS_SCALING}, batch={batch} BSC p={P_ERR} syndromes")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Number of samples received so far (to cut the window of interest out of the stream)."""
        return len(self.rows)

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)  # let the last samples of the window arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        rows = self.rows[first:] if len(self.rows) > first else self.rows[-3:]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_traffic():
    """DRAM bytes per launch of the message-update kernels from the committed ncu captures (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def host_cores():
    """CPU threads this process may actually use: min(affinity, cgroup cpu.max quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return max(1, n)


def build_code():
    from ldpc_b200 import codes
    return codes.regular_ldpc(N_CODE, DV, DC, seed=CODE_SEED)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation (unmodified C++ in oracle/_ref, else the C port) on host cores."""
    if rank != 0:
        return
    import oracle
    from ldpc_b200 import codes
    H = build_code()
    if oracle.have_ref():
        impl, kind = oracle.RefOracle(), "reference"
        cores = host_cores()
    else:
        if not oracle.have_port():
            oracle.build()
        impl, kind, cores = oracle.PortOracle(), "port", 1
    # bounded sample per step: ~2-4 s of CPU work at ~4-5k decodes/s/core
    sample = int(min(args.batch, max(2048, 4096 * cores)))
    syn = codes.bsc_syndromes(H, P_ERR, sample, seed=7)
    kw = dict(max_iter=MAX_ITER, bp_method="ms", schedule="parallel", ms_scaling_factor=MS_SCALING, want_llr=False)

    def step():
        t0 = time.perf_counter()
        if kind == "reference":
            impl.decode_batch(H, syn, P_ERR, threads=cores, **kw)
        else:
            impl.decode_batch(H, syn, P_ERR, **kw)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "decodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "sample_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "decodes/s", "cores": cores, "kind": kind,
                             "sample": f"{sample} syndromes of the workload per step, {cores} threads, one decoder "
                                       f"object per thread (reference C++ BpDecoder::decode per syndrome)"},
            "e2e": {"value": value, "unit": "decodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from ldpc_b200 import BpDecoder, _capi

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (ldpc_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # keep stdout to the ONE JSON line: native libraries (NCCL's version banner, ...) write to fd 1 directly, so park
    # the real stdout and point fd 1 at stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H = build_code()
    m, n = H.shape
    E = int(H.nnz)
    B = int(args.batch)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # synthetic BSC syndromes generated on the device (seeded per rank): e ~ Bernoulli(p), s = H e mod 2
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    Hd = torch.tensor(H.toarray(), dtype=torch.float16, device=dev)
    d_syn = torch.empty((B, m), dtype=torch.uint8, device=dev)
    chunk = 1 << 16
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        e = (torch.rand((hi - lo, n), device=dev, generator=gen) < P_ERR).to(torch.float16)
        d_syn[lo:hi] = (e @ Hd.T).to(torch.int32).remainder_(2).to(torch.uint8)
    del Hd

    dec = BpDecoder(H, error_rate=P_ERR, max_iter=MAX_ITER, bp_method="ms", ms_scaling_factor=MS_SCALING,
                    schedule="parallel", input_vector_type="syndrome", device=local_rank, kernel=args.kernel)
    h = dec._ensure_handle()
    L = _capi.lib()
    d_dec = torch.empty((B, n), dtype=torch.uint8, device=dev)
    d_conv = torch.empty(B, dtype=torch.uint8, device=dev)
    d_its = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def device_step():
        rc = L.bpb_decode_batch_device(h, _capi.INPUT_SYNDROME, C.c_void_p(d_syn.data_ptr()), B,
                                       C.c_void_p(d_dec.data_ptr()), C.c_void_p(d_conv.data_ptr()),
                                       C.c_void_p(d_its.data_ptr()), None, C.c_void_p(stream.cuda_stream))
        _capi.check(h, rc)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------------------
    # (the nvidia-smi sampler was started before the data generation: it needs ~1 s to deliver its first sample;
    #  the reported clocks are the samples taken from the first warm-up step to the end of the timed steps)
    warm = max(args.warmup, 3)
    first_sample = sampler.mark()
    for _ in range(warm):
        device_step()
    barrier()
    launches0 = dec.info()["launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kernel_ms = []
    ev[0].record(stream)
    for k in range(args.steps):
        device_step()
        ev[k + 1].record(stream)
    barrier()
    clocks = sampler.stop(first_sample) if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms.append(dec.info()["last_kernel_ms"])
    launches = dec.info()["launches"] - launches0
    total_ms = max_over_ranks(total_ms)
    value = B * world * args.steps / (total_ms * 1e-3)

    # statistics of the decoded batch (same every step: same inputs)
    its_sum = int(d_its.sum(dtype=torch.int64).item())
    conv_frac = float(d_conv.to(torch.float32).mean().item())

    # ---- roofline of the message-update kernel -----------------------------------------------------------
    # algorithmic bytes: per iteration 4*E*w (check pass reads b2c + writes c2b, bit pass reads
    # c2b + writes b2c; w = 8, binary64) + n + m; once per decode m + n + 4 + 1.
    w = 8
    info = dec.info()
    alg_bytes = its_sum * (4 * E * w + n + m) + B * (m + n + 5)
    if info["kernel_family"] == 1 and info["stream_iterations"] > 0:
        # streaming family: count the iterations the timed kernel itself executed (its ramp-down hands the last
        # stragglers to the second-stage kernel, whose time is not in kernel_ms)
        alg_bytes = info["stream_iterations"] * (4 * E * w + n + m) + (B - info["stream_handed_off"]) * (m + n + 5)
    kms = float(np.mean(kernel_ms))
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": {1: "bp_stream_kernel", 2: "bp_smem_kernel"}.get(info["kernel_family"], "?"),
                "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                "note": ("on-chip family: messages stay in shared memory, so the algorithmic HBM bytes are never moved "
                         "(frac > 1 by design; see roofline.traffic for the real DRAM bytes and DESIGN.md section 5.1)")
                if info["kernel_family"] == 2 else "streaming family: messages resident in HBM",
                "mean_iterations": its_sum / B, "converged_fraction": conv_frac,
                "handed_to_second_stage": int(info["stream_handed_off"]) if info["kernel_family"] == 1 else 0}
    traffic = load_traffic()
    if roofline["kernel"] in traffic and traffic[roofline["kernel"]].get("batch") == B:
        roofline["traffic"] = traffic[roofline["kernel"]]["dram_bytes_per_launch"]  # ncu --set full, per launch

    # ---- the streaming (HBM-resident) family on the same batch, for the record ----------------------------------
    stream_family = None
    if info["kernel_family"] == 2 and args.kernel == "auto" and not args.no_stream_family:
        sdec = BpDecoder(H, error_rate=P_ERR, max_iter=MAX_ITER, bp_method="ms", ms_scaling_factor=MS_SCALING,
                         schedule="parallel", input_vector_type="syndrome", device=local_rank, kernel="stream")
        sh = sdec._ensure_handle()

        def stream_step():
            rc = L.bpb_decode_batch_device(sh, _capi.INPUT_SYNDROME, C.c_void_p(d_syn.data_ptr()), B,
                                           C.c_void_p(d_dec.data_ptr()), C.c_void_p(d_conv.data_ptr()),
                                           C.c_void_p(d_its.data_ptr()), None, C.c_void_p(stream.cuda_stream))
            _capi.check(sh, rc)

        for _ in range(2):
            stream_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(2):
            stream_step()
        e1.record(stream)
        barrier()
        s_ms = max_over_ranks(e0.elapsed_time(e1)) / 2
        sinfo = sdec.info()
        s_bytes = sinfo["stream_iterations"] * (4 * E * w + n + m) + (B - sinfo["stream_handed_off"]) * (m + n + 5)
        s_ach = s_bytes / (sinfo["last_kernel_ms"] * 1e-3) / 1e9
        stream_family = {"value": B * world / (s_ms * 1e-3), "unit": "decodes/s", "ms_per_step": s_ms,
                         "roofline": {"bound": "hbm", "achieved": s_ach, "peak": peak, "unit": "GB/s",
                                      "frac": s_ach / peak, "kernel": "bp_stream_kernel",
                                      "kernel_ms": sinfo["last_kernel_ms"], "algorithmic_bytes_per_launch": s_bytes,
                                      "handed_to_second_stage": int(sinfo["stream_handed_off"]),
                                      "traffic": (traffic.get("bp_stream_kernel", {}).get("dram_bytes_per_launch")
                                                  if traffic.get("bp_stream_kernel", {}).get("batch") == B else None)},
                         "note": "same batch decoded by the HBM-streaming kernel family (kernel='stream'): messages "
                                 "laid out batch-minor in HBM, one lane per syndrome"}
        del sdec

    # ---- end to end through the host-buffer API (pinned host memory, copies inside the timed region) --------
    e2e_steps = max(1, min(args.steps, 3))
    pin_in = _capi.PinnedArray((B, m), np.uint8)
    pin_dec = _capi.PinnedArray((B, n), np.uint8)
    pin_conv = _capi.PinnedArray((B,), np.uint8)
    pin_its = _capi.PinnedArray((B,), np.int32)
    pin_in.array[...] = d_syn.cpu().numpy()

    def host_step():
        rc = L.bpb_decode_batch(h, _capi.INPUT_SYNDROME, _capi.host_ptr(pin_in.array), B,
                                _capi.host_ptr(pin_dec.array), _capi.host_ptr(pin_conv.array),
                                _capi.host_ptr(pin_its.array), None)
        _capi.check(h, rc)

    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = B * world * e2e_steps / e2e_s
    same = bool(np.array_equal(pin_dec.array[:4096], d_dec[:4096].cpu().numpy()))

    # ---- reference C++ on this box's host cores, bounded sample (rank 0, N = 1 only) ------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(H, pin_in.array)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "decodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(B), "batch_per_gpu": B, "parallelism": f"batch-shard x{world}",
                           "l2": "no flush needed: per-step working set (messages + I/O, several GB) >> 126 MB L2",
                           "kernel_family": roofline["kernel"], "grid": info["grid"], "block": info["block"]},
                "clocks": clocks, "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": "decodes/s", "h2d_bytes_per_step": B * m * world,
                        "d2h_bytes_per_step": B * (n + 5) * world, "steps": e2e_steps,
                        "matches_device_run": same},
                "gpu_launches": int(launches), "cpu_baseline": cpu_baseline}
        if stream_family is not None:
            line["stream_family"] = stream_family
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(H, syn_host):
    """Reference C++ (oracle/_ref) or the C port, on a bounded sample of the SAME syndromes (~10-20 s)."""
    import oracle
    cores = host_cores()
    kw = dict(max_iter=MAX_ITER, bp_method="ms", schedule="parallel", ms_scaling_factor=MS_SCALING, want_llr=False)
    if oracle.have_ref():
        ref = oracle.RefOracle()
        sample = int(min(syn_host.shape[0], 8192 * cores))
        syn = np.ascontiguousarray(syn_host[:sample])
        ref.decode_batch(H, syn[: 64 * cores], P_ERR, threads=cores, **kw)
        dt = ref.decode_batch(H, syn, P_ERR, threads=cores, return_seconds=True, **kw)[-1]
        return {"value": sample / dt, "unit": "decodes/s", "cores": cores, "kind": "reference",
                "sample": f"first {sample} syndromes of the GPU batch, {cores} threads (cgroup cpu quota) x one reference "
                          f"BpDecoder each, decode loop only"}
    if not oracle.have_port():
        oracle.build()
    port = oracle.PortOracle()
    sample = int(min(syn_host.shape[0], 32768))
    syn = np.ascontiguousarray(syn_host[:sample])
    t0 = time.perf_counter()
    port.decode_batch(H, syn, P_ERR, **kw)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "decodes/s", "cores": 1, "kind": "port",
            "sample": f"first {sample} syndromes of the GPU batch, single thread"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--kernel", default="auto", choices=["auto", "stream", "smem"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-family", action="store_true", help="skip the extra streaming-family measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

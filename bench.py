#!/usr/bin/env python
"""bench.py -- headline measurement of the RLE evaluation hot path on B200.

Workload (BASELINE.json configs[1], "C2"): float32 run-end encoded arrays of 1e9 elements
and 1e7 runs each (4 operands), evaluated through the fused 6-op element-wise DAG
    z = max(a*b + c, d) - abs(a) * 0.5f
with equal-run compaction, on one GPU.  A "step" is one such evaluation with the operands
already resident in HBM (`value`), or starting from pinned HOST buffers and ending with the
result runs back on the host (`e2e`).  Metric: distinct input runs merged per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): every rank evaluates its own coordinate-range shard of
the same size (weak scaling) and the shards are stitched with the fuse_stores boundary rule
(all-gather of first value / last value / run count; include/pyskip/detail/eval.hpp:581-588).

--impl reference times the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/build_ref.sh) on the host CPUs, on a bounded 1/10 sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELEMS = 1_000_000_000
N_RUNS = 10_000_000
N_OPERANDS = 4
SAMPLE_DIV = 10          # cpu baseline sample: 1e8 elements / 1e6 runs per operand
METRIC = "rle_eval_input_runs_per_s"
UNIT = "Gruns/s"


# ---- synthetic inputs ------------------------------------------------
def gen_operand(seed, n_elems, n_runs):
    """Run-end arrays: n_runs-1 distinct cut points ~U{1..n-1}; values (1+U{0..63})*0.25f with
    v_i != v_{i-1} (canonical), exactly representable so compaction really fires."""
    rng = np.random.default_rng(seed)
    cuts = np.unique(rng.integers(1, n_elems, size=int(n_runs * 1.01) + 16, dtype=np.int64))
    while len(cuts) < n_runs - 1:
        extra = rng.integers(1, n_elems, size=n_runs // 50 + 16, dtype=np.int64)
        cuts = np.unique(np.concatenate([cuts, extra]))
    if len(cuts) > n_runs - 1:
        drop = rng.choice(len(cuts), size=len(cuts) - (n_runs - 1), replace=False)
        cuts = np.delete(cuts, drop)
    ends = np.concatenate([cuts, [n_elems]]).astype(np.int32)
    steps = rng.integers(1, 64, size=n_runs, dtype=np.int64)
    vals = ((1 + np.cumsum(steps) % 64) * 0.25).astype(np.float32)
    return ends, vals


def dag_program(A):
    """Postfix program of max(a*b + c, d) - abs(a)*0.5f over sources a,b,c,d = 0..3."""
    F = A.OP_F
    half = A.value_bits(0.5, A.F32)
    return [(A.OP_SRC, 0), (A.OP_SRC, 1), (F["mul"], 0), (A.OP_SRC, 2), (F["add"], 0), (A.OP_SRC, 3),
            (F["max"], 0), (A.OP_SRC, 0), (F["abs"], 0), (A.OP_CONST, half), (F["mul"], 0), (F["sub"], 0)]


# ---- clocks (B200_PROFILING.md recipe) ---------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of k_merge_eval from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("merge_eval_dram_bytes_per_launch")
    except Exception:
        return None


# ---- the reference arm / cpu baseline ----------------------------------------------------
def reference_step_setup():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import _pyskip_cpp_ext as R           # the unmodified reference, compiled
    n, r = N_ELEMS // SAMPLE_DIV, N_RUNS // SAMPLE_DIV
    arrs = []
    for j in range(N_OPERANDS):
        ends, vals = gen_operand(2000 + j, n, r)
        lens = np.diff(np.concatenate([[0], ends.astype(np.int64)]))
        arrs.append(R.from_numpy(np.repeat(vals, lens)))
    a, b, c, d = arrs

    def step():
        z = ((a * b + c).max(d) - a.abs() * 0.5).eval()
        ends, vals = z.runs()
        return len(ends)

    return step, N_OPERANDS * r, f"1/{SAMPLE_DIV} of the workload: {n:.0e} elements, {r:.0e} runs per operand"


def time_reference(steps, warmup):
    step, runs_in, sample = reference_step_setup()
    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return runs_in / dt / 1e9, dt * 1e3, sample


def run_reference(args, rank):
    if rank != 0:
        return
    try:
        val, ms, sample = time_reference(args.steps, args.warmup)
    except Exception as ex:  # oracle/_ref missing
        print(json.dumps({"impl": "reference", "unavailable": f"{type(ex).__name__}: {ex}"}))
        return
    cores = os.cpu_count()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config():
    return {"workload": "C2: 4 x float32 RLE arrays, 1e9 elements / 1e7 runs each; "
                        "z = max(a*b + c, d) - abs(a)*0.5f + equal-run compaction",
            "runs_counted": "distinct input runs (4e7 per step per GPU)",
            "l2": "inputs 320 MB per step > 126 MB L2 (no flush needed)"}


# ---- our arm -----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    from pyskip_b200 import _abi as A
    lib = A.lib()
    assert not lib.is_simulator()
    torch.cuda.set_device(local_rank)
    lib.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_elems = int(N_ELEMS * args.scale)
    n_runs = int(N_RUNS * args.scale)
    # pinned host buffers (the e2e leg copies from / to these every step)
    host = []
    for j in range(N_OPERANDS):
        ends, vals = gen_operand(2000 + j + 100 * rank, n_elems, n_runs)
        pe = torch.empty(n_runs, dtype=torch.int32, pin_memory=True)
        pv = torch.empty(n_runs, dtype=torch.float32, pin_memory=True)
        pe.numpy()[:] = ends
        pv.numpy()[:] = vals
        host.append((pe, pv))
    out_e = torch.empty(N_OPERANDS * n_runs, dtype=torch.int32, pin_memory=True)
    out_v = torch.empty(N_OPERANDS * n_runs, dtype=torch.float32, pin_memory=True)
    prog = dag_program(A)

    def upload():
        return [lib.store_from_runs(pe.numpy(), pv.numpy()) for pe, pv in host]

    STITCH_BATCH = 32
    pending = [0]      # triples written on the device and not exchanged yet
    last_total = [None]

    def stitch(res, step):
        """fuse_stores boundary rule across ranks (eval.hpp:581-588): every rank contributes
        (first value, last value, run count) of its shard; a rank drops its last run when the
        next rank starts with an equal value, an exclusive scan of the counts gives global run
        offsets.  Each step appends its triple to a device buffer (one tiny library kernel, no
        host synchronisation); the ranks exchange the pending triples with ONE all-gather every
        STITCH_BATCH steps and at the end of the region -- the collective is latency-bound
        (24 bytes per rank and step), so it is batched rather than issued per step."""
        if dist is None:
            return
        lib.store_boundary_device(res, st_mine.data_ptr() + 24 * pending[0])
        pending[0] += 1
        if pending[0] == STITCH_BATCH:
            stitch_drain()

    def stitch_drain():
        """Exchange the pending triples; -> global run count of the last step."""
        if dist is None or pending[0] == 0:
            return last_total[0]
        n = pending[0]
        lib.synchronize()                              # the triples are written
        dist.all_gather_into_tensor(st_all[:world * n * 3], st_mine[:n * 3])
        allv = st_all[:world * n * 3].view(world, n, 3).cpu()
        counts = allv[:, n - 1, 2].clone()
        for r in range(world - 1):
            if allv[r, n - 1, 1] == allv[r + 1, n - 1, 0]:
                counts[r] -= 1
        pending[0] = 0
        last_total[0] = int(counts.sum())
        return last_total[0]

    if dist is not None:
        st_mine = torch.zeros(STITCH_BATCH * 3, dtype=torch.int64, device="cuda")
        st_all = torch.zeros(world * STITCH_BATCH * 3, dtype=torch.int64, device="cuda")
    resident = upload()
    srcs = [(s, []) for s in resident]

    step_no = [0]
    # The resident step goes through psk_eval_async: the kernels are queued and the call
    # returns; a result's run count is read two steps later, so the host work of step i+1
    # (plan, launches) overlaps the kernels of step i.  Every result is completed and checked
    # before the timed region ends.
    eval_resident = lib.prepare_eval(srcs, prog, A.F32, A.EQ_BITWISE, deferred=True)
    inflight = []
    expect_runs = [None]

    def complete(res):
        n = res.nruns                      # waits for that step's kernels
        if expect_runs[0] is None:
            expect_runs[0] = n
        assert n == expect_runs[0] and n > 0, f"step produced {n} runs, expected {expect_runs[0]}"

    def step_resident():
        res = eval_resident()
        stitch(res, step_no[0])
        step_no[0] += 1
        inflight.append(res)
        while len(inflight) > 2:
            complete(inflight.pop(0))
        return res

    def drain_resident():
        while inflight:
            complete(inflight.pop(0))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        lib.synchronize()

    for _ in range(args.warmup):
        res = step_resident()
    drain_resident()
    runs_out = res.nruns
    del res
    stitch_drain()

    # ---- timed region: K resident steps, device events on the library stream -------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    lib.profile(True)
    lib.launch_count(reset=True)
    t0 = time.time()
    lib.timer_start()
    for _ in range(args.steps):
        step_resident()
    drain_resident()                      # every step's result is complete and checked ...
    total_out = stitch_drain()            # ... and every step's exchange has completed inside the region
    ms_total = lib.timer_stop()
    barrier()
    t1 = time.time()
    launches = lib.launch_count()
    merge_ms, n_eval, prof_in, prof_out = lib.profile_read()
    lib.profile(False)
    clocks = sampler.stop(t0, t1)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    runs_in_step = N_OPERANDS * n_runs
    value = world * runs_in_step / (ms_step * 1e-3) / 1e9

    # ---- e2e: pinned host -> device -> eval -> runs back on the host ----------------------
    # Every step uploads its 4 operands from pinned host memory (320 MB), evaluates, and
    # downloads the result runs into pinned host memory.  The transfers use the library's
    # copy streams (psk_store_from_runs_async / psk_store_runs_async), so step i+1's upload
    # and step i's download overlap step i+1's kernels (PCIe is full duplex); nothing is
    # skipped or cached -- every byte moves every step, inside the timed region.
    outs = [(out_e, out_v),
            (torch.empty_like(out_e).pin_memory(), torch.empty_like(out_v).pin_memory())]

    def upload_async():
        return [lib.store_from_runs_async(pe.numpy(), pv.numpy()) for pe, pv in host]

    def run_e2e(n):
        pending = upload_async()
        prev = None
        n_out = 0
        verbose = os.environ.get("BENCH_E2E_VERBOSE") == "1"
        for i in range(n):
            if verbose:
                print(f"e2e step {i} starts at {time.perf_counter() * 1e3:.2f} ms", file=sys.stderr)
            nxt = upload_async() if i + 1 < n else None
            r = lib.eval([(s, []) for s in pending], prog, A.F32, A.EQ_BITWISE)
            n_out = r.nruns
            oe, ov = outs[i & 1]
            lib.store_runs_async(r, oe.numpy(), ov.numpy())
            prev = (r, pending)          # keep alive until their copies have been ordered
            pending = nxt
        lib.copy_synchronize()
        lib.synchronize()
        return n_out

    n_e2e = max(4, min(args.steps, 12))
    run_e2e(2)
    barrier()
    w0 = time.perf_counter()
    n_out = run_e2e(n_e2e)
    barrier()
    e2e_s = (time.perf_counter() - w0) / n_e2e
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * runs_in_step / e2e_s / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_merge_eval) -----------------------------------
    peak, peak_src = hbm_peak()
    avg_merge_ms = merge_ms / max(1, n_eval)
    alg_bytes = 8.0 * (runs_in_step + runs_out)      # (4 B end + 4 B value) per run read / written
    achieved = alg_bytes / (avg_merge_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "k_merge_eval", "kernel_ms": avg_merge_ms,
                "kernel_share_of_step": avg_merge_ms / ms_step, "algorithmic_bytes": alg_bytes,
                "peak_source": peak_src}

    # ---- cpu baseline: the reference itself on the host cores, bounded sample -------------
    try:
        cpu_val, cpu_ms, sample = time_reference(steps=3, warmup=1)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample,
               "ms_per_sample_step": cpu_ms}
    except Exception as ex:
        cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
               "sample": f"unavailable: {type(ex).__name__}: {ex}"}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "runs_out_per_step": runs_out, "output_gruns_per_s": world * runs_out / (ms_step * 1e-3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": 8 * runs_in_step, "d2h_bytes_per_step": 8 * int(n_out)},
        "gpu_launches": int(launches), "clocks": clocks,
    }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

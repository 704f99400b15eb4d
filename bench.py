#!/usr/bin/env python
"""bench.py -- structures/s of the fused feature + X^T X / X^T y build (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md 8d "config 2"): single-element gtinv polymlp, order 3,
maxl [4,4], cutoff 6 A, 10 radial functions, model_type 4 (F = 2030); synthetic 256-atom fcc 4x4x4
supercells (a = 4.05 A), Gaussian displacements sigma = 0.05 A, seed 20240+s, energy + forces + stress
rows (775 rows per structure), N(0,1) targets.  A step = one pass of the whole hot path (neighbour list
-> a_nlm -> invariants -> derivative rows -> polynomial X rows -> C += [X|y]^T [X|y]) over one batch of
structures per GPU.  One process per GPU; structures shard across ranks, partial accumulators are
combined with one NCCL reduce per step (weak scaling: per-GPU batch fixed).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # reference CPU path (oracle/_ref), rank 0 only
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402

METRIC = "structures/sec, fused feature + X^T X / X^T y build (fp64)"
WORKLOAD = ("config 2: gtinv order 3, maxl [4,4], rc 6 A, 10 radial fn, model_type 4 (F=2030); "
            "256-atom fcc 4x4x4 supercells, sigma 0.05 A, E+F+S rows (775/structure)")


def make_batch(n_st, first_seed):
    sts = [cases.fcc_supercell(seed=20240 + first_seed + s) for s in range(n_st)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    rows = n_st * 775
    rng = np.random.default_rng(first_seed)
    w = np.minimum(1.0, 1.0 / np.maximum(np.abs(rng.normal(size=rows)), 1e-12))  # force-like weights
    y = w * rng.normal(size=rows)
    return axis, pcs, tys, w, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def cpu_reference_rate(pd, n_st, steps=1, warmup=0):
    """Reference CPU path on the host cores: oracle/_ref (unmodified reference C++, OpenMP over structures,
    PyModel semantics) + weights + numpy x.T @ x, exactly as _compute_products_single_batch."""
    from oracle import ref

    cores = os.cpu_count() or 1
    ref.set_num_threads(cores)          # torchrun exports OMP_NUM_THREADS=1: use every host core, as the reference would
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)  # numpy's BLAS for x.T @ x
    except Exception:
        pass
    rm = ref.RefModel(pd)
    axis, pcs, tys, w, y = make_batch(n_st, 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        x = rm.build_x(axis, pcs, tys, [True] * n_st)
        xe = x[:n_st]
        _ = xe.sum(axis=0), np.square(xe).sum(axis=0)
        # reorder w/y (ours: PyModel layout already) and accumulate
        x *= w[:, None]
        xtx = x.T @ x
        xty = x.T @ y
        _ = y @ y
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del x, xtx, xty
    return n_st * len(times) / sum(times), ref.num_threads(), sum(times) / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from pypolymlp_b200.params import make_params_dict

    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    cores = os.cpu_count() or 1
    n_st = max(8, min(64, 2 * cores))   # two structures per OpenMP thread: every core busy, ~5-10 s per step
    k_steps, k_warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    rate, threads, step_s = cpu_reference_rate(pd, n_st, steps=k_steps, warmup=k_warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "structures/s", "n_gpus": args.gpus,
        "steps": k_steps, "warmup": k_warm, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "structures_per_step": n_st,
                   "note": "reference C++ (oracle/_ref, OpenMP over structures) + numpy x.T@x on host cores; "
                           "bounded sample of the workload, steps capped at 3"},
        "cpu_baseline": {"value": rate, "unit": "structures/s", "cores": threads, "kind": "reference",
                         "sample": f"{n_st} structures/step"},
        "e2e": {"value": rate, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--structures", type=int, default=256, help="structures per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes

    import torch
    import torch.distributed as dist

    from pypolymlp_b200 import fit
    from pypolymlp_b200._capi import StructureBatch, check, lib
    from pypolymlp_b200._capi import pd as pd_
    from pypolymlp_b200.libmlpcpp import PotentialXtX
    from pypolymlp_b200.params import make_params_dict

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    args.warmup = max(args.warmup, 3)
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    S = args.structures
    acc = PotentialXtX(pd, device=local_rank)
    ctx = acc.context
    F = acc.n_features
    axis, pcs, tys, w, y = make_batch(S, rank * S)
    batch_host = acc.stage(axis, pcs, tys, [True] * S, w, y)   # inputs resident in HBM
    stream = torch.cuda.ExternalStream(int(ctypes_stream(ctx)), device=local_rank)

    def barrier():
        torch.cuda.synchronize(local_rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    def step_device():
        acc.add_staged()
        if world > 1:
            fit.reduce_accumulator(acc, dst=0)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        # whichever is larger: the stream events cover the kernels, the wall clock also covers NCCL on its own stream
        ms = max(dev_ms, wall * 1e3) if world > 1 else dev_ms
        t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -----------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - l0
    value = world * S * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: host buffers in, X^T X / X^T y back on the host -----------
    w_h, y_h = np.ascontiguousarray(w), np.ascontiguousarray(y)

    def step_e2e():
        acc.reset()
        acc.add_batch(batch_host, w_h, y_h)
        if world > 1:
            fit.reduce_accumulator(acc, dst=0)
        if rank == 0:
            acc.finalize(copy=False)   # X^T X, X^T y ... in (pinned) host memory

    e2e_steps = max(2, min(args.steps, 4))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * S * e2e_steps / float(t.item())
    h2d = batch_host.h2d_bytes + w_h.nbytes + y_h.nbytes
    d2h = (F * F + 3 * F + 2) * 8

    # ---- secondary metric: atoms/s of E/F/S evaluation (BASELINE config 5), sharded by structure -----------
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast

    n_ev = int(os.environ.get("PM_BENCH_NEV", "32"))   # structures per evaluation call and GPU (config 5 holds 10 000)
    ev_sts = [cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777 + rank * n_ev + k) for k in range(n_ev)]
    ev_axis, ev_pcs, ev_tys = [x[0] for x in ev_sts], [x[1] for x in ev_sts], [x[2] for x in ev_sts]
    prop = PotentialPropertiesFast(pd, np.random.default_rng(12).normal(size=F) * 1e-3, device=local_rank)
    ev_batch = StructureBatch(ev_axis, ev_pcs, ev_tys, [True] * n_ev)
    ev_out = (np.zeros(n_ev), np.zeros((n_ev * 512, 3)), np.zeros((n_ev, 6)))

    def step_eval():
        check(lib().pm_eval(prop._ctx.handle, ctypes.byref(ev_batch.c), pd_(ev_out[0]), pd_(ev_out[1]), pd_(ev_out[2])))

    step_eval()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        step_eval()
    barrier()
    tev = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tev, op=dist.ReduceOp.MAX)
    eval_atoms_s = world * n_ev * 512 * 3 / float(tev.item())

    if rank == 0:
        # ---- roofline of the dominant kernel (SYRK, DMMA): one profiled step, CUDA events per stage -------
        acc.reset()
        ctx.profile(True)
        acc.add_staged()
        ctx.synchronize()
        prof = ctx.profile_get()
        ctx.profile(False)
        w_alg = acc.model.count_flops([256], [[13824]], True)
        # the dominant kernel timed live: event pairs around every SYRK launch of K unperturbed steps, on its stream
        ctx.profile(2)
        for _ in range(args.steps):
            acc.add_staged()
        ctx.synchronize()
        syrk_total_ms, syrk_launches = ctx.profile_get()["syrk"]
        ctx.profile(False)
        syrk_ms = syrk_total_ms / args.steps
        syrk_tflops = S * w_alg["syrk"] / (syrk_ms * 1e-3) * 1e-12
        peak = ctx.microbench(3, 8192)
        total_alg = S * sum(w_alg.values())
        stage_ms = {k: round(v[0], 3) for k, v in prof.items() if v[0] > 0}
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        # DRAM traffic of one SYRK launch from the committed ncu --set full capture (profiles/r01_ncu_kernels.json):
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (K5 runs per block of <= 32768 X rows)
        traffic, traffic_alg, rows_cap = None, None, 0
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_kernels.json")))["k_syrk_sk"]
            met, rows_cap = cap["metrics"], cap.get("rows_in_launch", 32768)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = sum(float(met[k]["value"]) * scale[met[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            fpad = (F + 1 + 127) // 128 * 128
            traffic_alg = rows_cap * fpad * 8.0 + fpad * fpad * 8.0 * 0.53 * 2   # X-tilde rows once + upper C tiles RMW
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "kernel": "k_syrk_sk (fp64 DMMA m8n8k4, stream-K SYRK)", "achieved": syrk_tflops, "peak": peak,
            "unit": "TFLOP/s", "frac": syrk_tflops / peak, "traffic": traffic,
            "traffic_note": "DRAM bytes of the captured SYRK launch (%d X rows; ncu --set full, " % (rows_cap if traffic else 0) +
                            "profiles/r01_ncu_kernels.json); algorithmic bytes of that launch in traffic_algorithmic",
            "traffic_algorithmic": traffic_alg,
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no fp64 entry; "
                           "tcgen05 has no f64 kind, DMMA peak == DFMA peak on B200)",
            "syrk_ms_per_step_live": syrk_ms, "syrk_launches_per_step": syrk_launches / args.steps,
            "algorithmic_flops_per_structure": w_alg, "stage_ms_one_step": stage_ms,
            "stage_note": "stage_ms_one_step comes from one extra step with a host sync after every stage (idle gaps "
                          "inflate it by ~8 %); achieved/frac use syrk_ms_per_step_live",
            "whole_step_tflops": total_alg / (ms / args.steps * 1e-3) * 1e-12,
            "whole_step_frac_of_peak": total_alg / (ms / args.steps * 1e-3) * 1e-12 / peak,
            "hbm_gbs_measured": hbm,
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref

            if ref.available():
                n_cpu = max(8, min(64, 2 * (os.cpu_count() or 1)))   # ~5-15 s of CPU work, every core busy
                rate, threads, _ = cpu_reference_rate(pd, n_cpu)
                cpu = {"value": rate, "unit": "structures/s", "cores": threads, "kind": "reference",
                       "sample": f"{n_cpu} structures of the same workload through oracle/_ref + numpy x.T@x"}
            else:
                cpu = {"value": None, "unit": "structures/s", "cores": 0, "kind": "reference",
                       "sample": "oracle/_ref not built on this box"}
        line = {
            "metric": METRIC, "value": value, "unit": "structures/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "structures_per_step_per_gpu": S, "n_features": F,
                       "l2": "no flush: each step streams ~200 MB of intermediates per structure through HBM "
                             "(51 GB per 256-structure step), far above the 126 MB L2; staged inputs are ~6 KB/structure",
                       "parallelism": f"structures sharded over {world} GPU(s), one NCCL reduce per step"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "structures/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": "pm_fit_reset + pm_fit_accumulate(host buffers) + pm_fit_finalize_view (X^T X, X^T y to pinned host memory) per step"},
            "roofline": roofline, "cpu_baseline": cpu,
            "eval": {"metric": "atoms/sec, E/F/S evaluation (config 5: F=2030 model, 512-atom fcc, sigma 0.03 A)",
                     "value": eval_atoms_s, "unit": "atoms/s", "structures_per_step_per_gpu": n_ev,
                     "note": "pm_eval through the C ABI: host arrays in, E/F/S back on the host (end to end)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ctypes_stream(ctx):
    from pypolymlp_b200._capi import lib

    return lib().pm_stream(ctx.handle)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: batched GP log-likelihood evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--epochs 512] [--batch 65536]

Workload (BASELINE.json metric / SURVEY.md 8d "C5"): synthetic two-image,
single-band light curve with N = 512 epochs, ExpSquaredKernel + UniformMean +
ConstantMagnification (P = 5), B = 65 536 theta per GPU drawn uniformly from the
SLSN-notebook prior box.  One step = one pass of the hot path over that batch.

Own arm: `value` is timed with CUDA events on the launching stream with theta
already resident in HBM (L2 flushed between steps); `e2e` is the same batch
through GP.loglikelihood_batch with HOST numpy buffers (pinned staging, H2D of
theta and D2H of logL/info inside the call).  For N > 1 ranks (torchrun), every
rank evaluates its own B theta (weak scaling), the logL vectors are all-gathered
over NCCL inside the step, and the time is the max over ranks.

`--impl reference` times the CPU path (the numpy/scipy oracle port of
gaussn/gausSN.py:135-160; the reference itself needs JAX, which is absent) on all
host cores, on a bounded sample of the same workload.
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

import benchdata  # noqa: E402

KERNEL, MEAN, LENS = "ExpSquaredKernel", "UniformMean", "ConstantMagnification"
METRIC = "gp_logl_evals_per_sec"
UNIT = "evals/s"


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    theta, ds = args
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        import contextlib
        ctx = contextlib.nullcontext()
    from oracle import gp_oracle
    with ctx:
        t0 = time.perf_counter()
        out = gp_oracle.loglikelihood_batch(theta, ds["x"], ds["y"], ds["yerr"], KERNEL, MEAN, LENS, ds["n_bands"],
                                            ds["n_images"], ds["indices"])
        return out, time.perf_counter() - t0


def cpu_baseline(ds, theta, per_core: int, cores: int | None = None):
    """Oracle port on `cores` processes (BLAS threads = 1 each); returns evals/s over all cores."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    n = per_core * cores
    sample = theta[:n]
    chunks = [sample[i::cores] for i in range(cores)]
    small = {k: ds[k] for k in ("x", "y", "yerr", "n_bands", "n_images", "indices")}
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[:1], small) for c in chunks])          # warm-up: imports, BLAS init
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(c, small) for c in chunks])
        wall = time.perf_counter() - t0
    per_core_rate = float(np.mean([len(c) / r[1] for c, r in zip(chunks, res)]))
    vals = np.empty(n)
    for i in range(cores):
        vals[i::cores] = res[i][0]
    return dict(value=n / wall, unit=UNIT, cores=cores, kind="port", per_core=per_core_rate, wall_s=wall,
                sample=f"{n} of the batch's theta ({per_core}/core), oracle/gp_oracle.py (numpy {np.__version__} + scipy LAPACK), "
                       f"one process per core, BLAS threads=1"), vals


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc, self.device = [], None, device

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        # "under load": samples drawing more than half of the maximum power seen
        load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return dict(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=float(max(power)))


# ----------------------------------------------------------------------------- arms
def fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json carries HBM and bf16 only, so the
    FP64 peak was measured on this pool's B200 with bench/peaks.cu (DMMA chain, DFMA
    chain, cuBLAS DGEMM; profiles/r01_fp64_peaks.json) and the maximum is used."""
    path = os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")
    try:
        with open(path) as f:
            pk = json.load(f)
        vals = {k: pk[k] for k in ("dmma_tflops_sustained", "dmma_tflops", "dfma_tflops") if k in pk}
        vals["dgemm_8192"] = pk.get("dgemm_tflops", {}).get("8192", 0.0)
        best = max(vals, key=vals.get)
        return float(vals[best]), f"measured on this pool's B200 by bench/peaks.cu ({best}; profiles/r01_fp64_peaks.json); MEASURED_PEAKS.json has no FP64 entry"
    except Exception:
        return 37.2, "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz (profiles/r01_fp64_peaks.json missing)"


def ncu_traffic(N, B):
    """dram bytes per launch from the committed ncu --set full capture of the same kernel, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        key = f"N{N}_B{B}"
        return t.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ds = benchdata.make_dataset(args.epochs, 1, 2)
    theta = benchdata.make_theta(args.batch, 2)
    cores = len(os.sched_getaffinity(0))
    # size one step at roughly 3 s of wall time: ~12 ms/eval/core at N = 512, scaling as N^3
    per_core = max(2, int(round(3.0 / (0.012 * (args.epochs / 512.0) ** 3))))
    per_core = min(per_core, max(2, args.batch // cores))
    times, n = [], per_core * cores
    base = None
    for s in range(args.warmup + args.steps):
        off = (s * n) % max(1, args.batch - n)
        base, _ = cpu_baseline(ds, theta[off:off + n], per_core, cores)
        if s >= args.warmup:
            times.append(base["wall_s"])
    total = float(np.sum(times))
    value = n * args.steps / total
    base["value"] = value
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=_config(args, n_step=n),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=base["sample"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def _config(args, n_step=None):
    return dict(workload=f"synthetic lensed SN light curve: N={args.epochs} epochs (1 band x 2 images), {KERNEL}+{MEAN}+{LENS}, "
                         f"P=5, B={args.batch} theta per GPU per step, theta uniform in the SLSN-notebook prior box",
                epochs=args.epochs, batch_per_gpu=args.batch, theta_per_step=n_step or args.batch * args.gpus,
                l2="flushed between timed steps (256 MiB memset outside the event pairs); the factor workspace alone exceeds L2",
                parallelism=f"theta-sharded x{args.gpus}", seeds=dict(data=benchdata.DATA_SEED, theta=benchdata.THETA_SEED))


def run_own(args):
    import torch
    import torch.distributed as dist
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; gaussn_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, B = args.epochs, args.batch
    ds = benchdata.make_dataset(N, 1, 2)
    theta_all = benchdata.make_theta(B * world, 2)
    theta_h = np.ascontiguousarray(theta_all[rank * B:(rank + 1) * B])      # this rank's shard
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]),
                   lensingmodels.ConstantMagnification([20.0, 0.5]), device=local)
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    gp._ensure_problem(gp.x, gp.y, gp.yerr)
    prob = gp._problem
    theta_d = torch.as_tensor(theta_h, device=dev)
    logl_d = torch.empty(B, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    fused, gathered, collective = None, None, "none (single GPU)"
    if world > 1:
        # NVLink-native gather: the kernel stores each logL into every rank's gathered buffer (peer-mapped symmetric
        # memory), followed by a barrier; falls back to an NCCL all-gather if symmetric memory cannot be set up here
        try:
            if args.gather == "nccl":
                raise RuntimeError("NCCL all-gather requested")
            from gaussn_b200 import sharding
            fused = sharding.FusedGather(B, dev)
            collective = "fused into the kernel epilogue: peer stores over NVLink into symmetric memory + barrier"
        except Exception as exc:  # noqa: BLE001
            gathered = torch.empty(B * world, dtype=torch.float64, device=dev)
            collective = f"NCCL all_gather_into_tensor of {B} doubles per rank ({type(exc).__name__}: {exc})"

    def step():
        if fused is not None:
            fused.run(prob, theta_d)
            return
        prob.logl_device(theta_d.data_ptr(), B, theta_d.shape[1], logl_d.data_ptr(), 0, 0, stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logl_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = prob.query(_engine.Q_LAUNCHES)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        for s in range(args.steps):
            flush.zero_()                      # evict L2 (outside the event pair)
            ev[s][0].record(stream)
            step()
            ev[s][1].record(stream)
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = prob.query(_engine.Q_LAUNCHES) - launches0
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = B * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (wall clock; the call is synchronous)
    gp.loglikelihood_batch(theta_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_h = gp.loglikelihood_batch(theta_h)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B * world * args.steps / e2e_s
    dev_vals = (fused.buf[rank * B:(rank + 1) * B] if fused is not None else logl_d).cpu().numpy()
    same = bool(np.array_equal(out_h, dev_vals, equal_nan=True))

    if rank == 0:
        flops = benchdata.algorithmic_flops(N, 1, True, KERNEL)
        peak, peak_src = fp64_peak()
        kernel_ms = float(np.mean(step_ms))      # one kernel launch per step; the all-gather (N>1) is microseconds
        achieved = flops * B / (kernel_ms * 1e-3) / 1e12
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", config=dict(_config(args), collective=collective), clocks=clocks.summary(),
                    gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(theta_h.nbytes), d2h_bytes_per_step=int(B * 12),
                             identical_to_device_path=same),
                    roofline=dict(bound="tensor", pipe="fp64 (DMMA.8x8x4, shared with DFMA)", achieved=achieved, peak=peak,
                                  unit="TFLOP/s", frac=achieved / peak, traffic=ncu_traffic(N, B), flops_per_eval=flops,
                                  peak_source=peak_src, kernel="gsn::logl_dmma_kernel<0, gsn::CfgWide>", kernel_ms=kernel_ms,
                                  algorithmic_hbm_bytes_per_eval=int(theta_h.shape[1] * 8 + 8)),
                    nan_fraction=float(np.mean(~np.isfinite(out_h))))
        if world == 1 and not args.no_cpu:
            per_core = max(2, int(round(1.0 / (0.012 * (N / 512.0) ** 3))))
            base, cpu_vals = cpu_baseline(ds, theta_h, per_core)
            fin = np.isfinite(cpu_vals)
            err = np.abs(out_h[:len(cpu_vals)][fin] - cpu_vals[fin])
            base["max_abs_dlogl_vs_gpu"] = float(err.max()) if fin.any() else None
            base["nan_pattern_agrees"] = bool(np.array_equal(np.isnan(cpu_vals), np.isnan(out_h[:len(cpu_vals)])))
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--epochs", type=int, default=512)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="multi-GPU result gather (default: fused peer stores)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()

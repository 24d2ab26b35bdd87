#!/usr/bin/env python
"""bench.py -- FP64 Mcell-updates/s of chiML's time-stepping hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (oracle/_ref/chiml_ref)

Workload (config.workload): BASELINE.md C5, weak scaling -- per GPU one y-slab of 2048 x 256 x 1024 grid points of the 3-D
anisotropic (oriented-dipole Lorentz) slab waveguide + two-level emitter sheet + CPML domain; N GPUs step a
2048 x 256N x 1024 grid.  One "step" = one leap-frog time step of the whole grid.  Inputs are synthetic (the JSON a user
would write, built by chiml_b200/inputs.py from the reference's input contract); all state is resident in HBM (every array
is > 4 GB, i.e. far larger than L2, so no flush is needed between steps).

`value`   : cell-updates/s with CUDA events around K steps of chiml_gpu_step_n, max over ranks.
`e2e`     : the same K steps driven the way the host driver of the reference does it -- one chiml_gpu_step_n(1) per step with
            the source amplitudes of that step in HOST memory, and the detector samples of that step read back to HOST
            memory -- wall clock on the host, copies inside the timed region.
`roofline`: the dominant kernel's algorithmic bytes per launch / its average launch duration (CUDA events around every
            launch, taken in the same timed region) against the measured HBM peak of MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the unmodified reference (compiled in place by oracle/Makefile) on the host cores, one
            thread per y-slab rank, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from chiml_b200 import inputs as I  # noqa: E402

METRIC = "fp64_cell_updates_per_s"
UNIT = "Mcell/s"
WORKLOAD_HAS_EMITTERS = True     # the two-level emitter sheet of C4/C5 is part of the workload
PLAN_TOOL = os.path.join(ROOT, "chiml_b200", "chiml_plan")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2048, help="grid points along x (whole grid)")
    ap.add_argument("--ny-per-gpu", type=int, default=256, help="grid points along y per GPU (y-slab height)")
    ap.add_argument("--nz", type=int, default=1024, help="grid points along z")
    ap.add_argument("--workload", default="c5", choices=["c5", "c1", "c2", "c3", "c4"],
                    help="c5 (default) is the bench contract; c1..c4 are the other BASELINE.md configurations at full size, for the record "
                         "(single GPU, no CPU baseline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the bench contract): --ny-per-gpu rows per GPU; strong: the fixed BASELINE C5 grid nx x --ny-total x nz "
                         "cut into N y-slabs (2048 x 2048 x 1024 needs >= 4 GPUs: ~52 GB of state per 2048 x 256 x 1024 slab)")
    ap.add_argument("--ny-total", type=int, default=2048, help="grid points along y of the fixed grid of --scaling strong")
    ap.add_argument("--no-halo-parity", action="store_true", help="N > 1: skip the bit-exactness check of the native halo before the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--keep", action="store_true", help="keep the scratch directory")
    return ap.parse_args()


OTHER = {"c1": ("C1: 2-D TE vacuum 512x512, Hz dipole, CPML 20, Hz detector", lambda steps: I.c1_te_vacuum(n=511, steps=steps)),
         "c2": ("C2: 2-D TM Drude nanorod 2048x2048, CPML 20, Ez line source, 4-edge flux box with 64 frequencies (running DFT every step)",
                lambda steps: I.c2_tm_drude(n=2047, steps=steps, nfreq=int(os.environ.get("CHIML_BENCH_C2_NFREQ", "64")))),
         "c3": ("C3: 3-D anisotropic (oriented-dipole Lorentz) slab waveguide 512^3, CPML all faces", lambda steps: I.c3_aniso_slab(n=511, steps=steps)),
         "c4": ("C4: 3-D 10x10 Au (6-pole) cubes + two-level emitter sheet (1e6 emitters), 768^3", lambda steps: I.c4_plasmonic_ml(n=767, steps=steps))}


# the reference on the host cores beside a --workload c1..c4 run: (grid points of the sample, timed steps, input builder).  C1 and C2 are run at
# their full size; C3 and C4 on a 192 x 512 x 192 sample of the same construction (2 x 2 cubes under a 180 x 180 emitter sheet for C4)
OTHER_SAMPLE = {"c1": ((512, 512, 1), 200, lambda nx, ny, nz, steps: I.c1_te_vacuum(n=nx - 1, steps=steps)),
                "c2": ((2048, 2048, 1), 20, lambda nx, ny, nz, steps: I.c2_tm_drude(n=nx - 1, steps=steps, nfreq=64)),
                "c3": ((192, 512, 192), 10, lambda nx, ny, nz, steps: I.c3_aniso_slab(n=nx - 1, ny=ny - 1, nz=nz - 1, steps=steps)),
                "c4": ((192, 512, 192), 10, lambda nx, ny, nz, steps: I.c4_plasmonic_ml(n=nx - 1, ny=ny - 1, nz=nz - 1, steps=steps, narray=2, sheet=180))}


def workload_cfg(nx_pts: int, ny_pts: int, nz_pts: int, steps: int):
    return I.c5_aniso_ml(nx=nx_pts - 1, ny=ny_pts - 1, nz=nz_pts - 1, steps=steps, sheet=WORKLOAD_HAS_EMITTERS, out="bench_out/c5")


def workload_name(nx, nyg, nz, n, scaling="weak"):
    sheet = " + two-level emitter sheet" if WORKLOAD_HAS_EMITTERS else ""
    kind = "weak-scaling slab" if scaling == "weak" else "strong scaling of the fixed grid"
    return (f"C5 {kind}: 3-D anisotropic (oriented-dipole Lorentz) slab waveguide{sheet} + CPML 20 cells, "
            f"{nx}x{nyg}x{nz} grid points per GPU ({nx}x{nyg * n}x{nz} total)")


# N > 1: fixtures stepped across the N ranks over the native halo right before the timed region and compared with the committed
# single-rank output of the unmodified reference (tests/golden/<case>.expect.npz): oriented-dipole slab through the CPML (node P_y
# ghost rows), Au cubes under an emitter sheet (emitter P_y rim, a slab that holds only the rim), finite oriented-dipole objects
# with different pole counts (slabs without node cells)
HALO_PARITY_CASES = ["aniso_slab3d", "c4_small", "aniso_mixed3d"]


def halo_parity(dist, rank, world, local):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import slab_gpu_worker
    import torch
    log = lambda m: print(m, file=sys.stderr)   # noqa: E731
    ok = True
    for case in HALO_PARITY_CASES:
        ok = slab_gpu_worker.run_case(case, rank, world, local, log=log) and ok
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    if int(flag.item()) != 1:
        raise SystemExit("bench.py: the native halo does not reproduce the single-rank reference fixtures -- refusing to time it")
    return {"result": "bit-identical", "cases": HALO_PARITY_CASES, "slabs": world,
            "against": "tests/golden/<case>.expect.npz (single-rank output of the unmodified reference)"}


# ---------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference on the host cores
# ---------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(sample_pts, steps: int, warmup: int, work: str, cfg_fn=None):
    """Runs oracle/_ref/chiml_ref (the reference's own sources compiled in place) on a `sample_pts` grid of the workload,
    one in-process rank (thread) per y-slab.  Returns (Mcell/s, ms_per_step, info).  cfg_fn(nx, ny, nz, steps) builds the input of a
    workload other than C5 at that grid (ny = 32 rows per rank; 2-D workloads pass nz = 1)."""
    if not os.path.exists(REF_BIN):
        raise RuntimeError(f"{REF_BIN} missing: run `make -C oracle ref` where /root/reference is present")
    cores = host_cores()
    ranks = max(1, min(cores, 16))
    # the reference's y-slab ranks must each be taller than the 20-cell CPML: 32 grid rows per rank
    nx, nz = sample_pts[0], sample_pts[2]
    ny = 32 * ranks
    if cfg_fn is not None and sample_pts[1]:
        ny = sample_pts[1]                      # the workload's own height (a multiple of the rank count, >= 32 rows per rank)
        ranks = max(1, min(ranks, ny // 32))
    cfg = (cfg_fn or workload_cfg)(nx, ny, nz, steps + warmup)
    os.makedirs(work, exist_ok=True)
    jpath = os.path.join(work, "ref_sample.json")
    I.write(cfg, jpath)
    t0 = time.time()
    r = subprocess.run([REF_BIN, "ref_sample.json", "--ranks", str(ranks), "--steps", str(steps), "--warmup", str(warmup), "--quiet", "--no-output"],
                       cwd=work, capture_output=True, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError(f"chiml_ref failed ({r.returncode}): {r.stderr[-2000:]}")
    out = json.loads(r.stdout.strip().splitlines()[-1])
    cells = nx * ny * nz
    sec = out["step_seconds"]
    return cells * steps / sec / 1e6, sec / steps * 1e3, {
        "cores": ranks, "kind": "reference", "cells": cells,
        "sample": f"same workload at {nx}x{ny}x{nz} grid points ({cells / 1e6:.1f} Mcell), {steps} timed steps after {warmup} warm-up, "
                  f"{ranks} y-slab ranks as threads on {cores} host cores; {wall:.0f} s wall including the reference's setup"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    work = tempfile.mkdtemp(prefix="chiml_bench_ref_")
    try:
        sample = (384, 0, 192)
        v, ms, info = run_reference(sample, args.steps, args.warmup, work)
        full_cells = args.nx * args.ny_per_gpu * args.gpus * args.nz
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.nx, args.ny_per_gpu, args.nz, args.gpus), "reference_sample": info["sample"],
                           "reference_sample_cells": info["cells"], "reference_sample_reduction": round(full_cells / info["cells"], 2),
                           "same_config": False},
                "cpu_baseline": dict(info, value=v, unit=UNIT),
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
    finally:
        if not args.keep:
            shutil.rmtree(work, ignore_errors=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------------------
def b200_arm(args):
    import numpy as np
    from chiml_b200 import capi, census, plan as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the engine has no CPU fallback")
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, W = args.steps, args.warmup
    nx, nyg, nz = args.nx, args.ny_per_gpu, args.nz
    work = tempfile.mkdtemp(prefix=f"chiml_bench_r{rank}_")
    try:
        # the input a user would write, then the host-side setup (C++) for this rank's y-slab
        total_steps = W + 3 * K + 8
        if args.workload != "c5":
            if world != 1:
                raise SystemExit("bench.py: --workload c1..c4 are single-GPU record runs")
            cfg = OTHER[args.workload][1](total_steps)
        else:
            if args.scaling == "strong":
                if args.ny_total % world:
                    raise SystemExit(f"bench.py: --scaling strong needs --ny-total ({args.ny_total}) divisible by the number of GPUs")
                nyg = args.ny_total // world
                if nx * nyg * nz > 2048 * 768 * 1024:
                    raise SystemExit(f"bench.py: a {nx}x{nyg}x{nz} slab (~{52 * nx * nyg * nz / (2048 * 256 * 1024):.0f} GB of state) does not fit one "
                                     "B200: strong scaling of the fixed C5 grid starts at 4 GPUs")
            cfg = workload_cfg(nx, nyg * world, nz, total_steps)
        jpath = os.path.join(work, "bench.json")
        I.write(cfg, jpath)
        t0 = time.time()
        r = subprocess.run([PLAN_TOOL, jpath, os.path.join(work, "bench"), "--ranks", str(world), "--only", str(rank)], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f"chiml_plan failed: {r.stderr[-2000:]}")
        plan = P.read_plan(os.path.join(work, f"bench.rank{rank}.plan"))
        cs = census.census(plan)
        sim = capi.GpuSim(plan, device=local)
        setup_s = time.time() - t0
        parity = None
        if world > 1:
            sim.halo_bind(dist, rank, world)
            if not args.no_halo_parity:
                parity = halo_parity(dist, rank, world, local)
        cells_local = cs.cells
        cells_total = allsum(float(cells_local))

        # warm-up
        sim.step_n(W)
        sim.sync()
        # ---- device-timed region: K steps, CUDA events on the engine's stream, per-kernel events inside ----
        sim.reset_kernel_stats()
        sim.set_kernel_timing(True)
        clocks = ClockSampler(local)
        clocks.start()
        l0 = sim.launch_count()
        barrier()
        ms = sim.step_n_timed(K)
        sim.sync()
        barrier()
        launches = sim.launch_count() - l0
        stats = sim.kernel_stats()
        sim.set_kernel_timing(False)
        ms = allmax(ms)
        value = cells_total * K / (ms * 1e-3) / 1e6

        # ---- end to end through the host-facing calls: per step, host amplitudes in, detector samples out ----
        nsrc = max(1, len(plan.sources))
        ndet = len(plan.detectors)
        det_buf = [np.empty(max(1, int(np.prod(capi.local_box(plan, d.loc, d.sz)[1]))) if capi.local_box(plan, d.loc, d.sz) else 1) for d in plan.detectors]
        det_next = [sim.steps_done // max(1, d.every) + 1 for d in plan.detectors]
        h2d = d2h = 0
        barrier()
        sim.sync()
        t0 = time.perf_counter()
        for k in range(K):
            amp = sim.src_amp(sim.steps_done, 1)            # host buffer of this step's source amplitudes
            sim.step_n(1, amp)
            h2d += amp.nbytes if plan.sources else 0
            for di in range(ndet):
                got = sim.detector_range(di, det_next[di], 1, det_buf[di])
                det_next[di] += got
                sim.consume_detector(di, det_next[di])      # read, then consume: the device ring stays at its initial size
                d2h += got * det_buf[di].nbytes
            if ndet == 0:
                sim.sync()
        sim.sync()
        barrier()
        e2e_s = allmax(time.perf_counter() - t0)
        h2d, d2h = allsum(float(h2d)), allsum(float(d2h))     # whole job: the source / detector cells live in one or two slabs
        clk = clocks.stop()
        e2e_value = cells_total * K / e2e_s / 1e6

        # ---- roofline of the dominant kernel, on the rank that moves the most bytes ----
        # (slabs with a neighbour lose a y-CPML face: the end slabs of a weak-scaling run carry more psi traffic than the middle ones,
        # and rank 0 of an N-GPU run less than the single GPU of N = 1)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        mine = {"rank": rank, "bytes_per_step": cs.bytes_per_step, "cells": cells_local, "stats": stats, "census": cs.as_dict()}
        ranks_info = [mine]
        if dist is not None:
            ranks_info = [None] * world
            dist.all_gather_object(ranks_info, mine)
        heavy = max(ranks_info, key=lambda r: r["bytes_per_step"])
        ms_step = ms / K

        def kernel_rows(st):
            rows = []
            for s_ in st:
                if s_["timed_launches"] <= 0:
                    continue
                t_step = s_["ms_total"] / K                       # device time of ALL launches of this kernel in one step
                rows.append({"name": s_["name"], "launches_per_step": s_["launches"] / K, "ms_per_step": t_step,
                             "avg_launch_ms": s_["ms_total"] / s_["timed_launches"], "share_of_step": t_step / ms_step,
                             "alg_GB_per_step": s_["alg_bytes_per_step"] / 1e9,
                             "alg_GBps": s_["alg_bytes_per_step"] / (t_step * 1e-3) / 1e9 if s_["alg_bytes_per_step"] else None})
            return rows
        kernels = kernel_rows(heavy["stats"])
        dom = max(kernels, key=lambda k_: k_["ms_per_step"])
        achieved = dom["alg_GBps"] or 0.0
        per_rank = [{"rank": r["rank"], "alg_bytes_per_step": r["bytes_per_step"], "achieved": r["bytes_per_step"] / (ms_step * 1e-3) / 1e9,
                     "frac": r["bytes_per_step"] / (ms_step * 1e-3) / 1e9 / peak} for r in ranks_info]
        step_bytes = heavy["bytes_per_step"]
        step_gbps = step_bytes / (ms_step * 1e-3) / 1e9
        sig = f"{args.workload}:{args.scaling}:{nx}x{nyg}x{nz}:n{world}"
        roofline = {"bound": "hbm", "kernel": dom["name"], "rank": heavy["rank"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src,
                    "how": "algorithmic bytes all launches of the kernel move in one step / their summed device time in one step (CUDA events around "
                           "every launch), on the rank with the most bytes per step",
                    "whole_step": {"alg_bytes_per_step_per_gpu": step_bytes, "achieved": step_gbps, "frac": step_gbps / peak,
                                   "per_rank": per_rank, "min_frac": min(r["frac"] for r in per_rank)},
                    "kernels": kernels}
        # DRAM traffic per launch is an ncu measurement: it is quoted only for the configuration it was captured on
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                tj = json.load(open(tr))
                if tj.get("_config") == sig:
                    roofline["traffic"] = tj.get(dom["name"])
                    roofline["traffic_source"] = tj.get("_comment")
            except Exception:
                pass

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(nx, nyg, nz, world, args.scaling) if args.workload == "c5" else OTHER[args.workload][0], "signature": sig, "l2": "inputs_larger_than_L2 (every state array > 4 GB)" if cells_local * 8 > 2.5e8 else "small grid: arrays may fit L2",
                           "census_rank0": cs.as_dict(), "device_GB_rank0": sim.device_bytes() / 1e9, "setup_s_rank0": round(setup_s, 1),
                           "fields": "zero initial state driven by the dipole source (reference behaviour); timing is data-independent"},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "ms_per_step": e2e_s / K * 1e3},
                "gpu_launches": int(launches),
                "roofline": roofline}
        if parity is not None:
            line["halo_parity"] = parity
        sim.close()
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                if args.workload == "c5":
                    v, cms, info = run_reference((256, 0, 128), 10, 2, work)
                else:
                    pts, st, fn = OTHER_SAMPLE[args.workload]
                    v, cms, info = run_reference(pts, st, 2, work, fn)
                    info["sample_reduction"] = round(cells_local / info["cells"], 2)
                line["cpu_baseline"] = dict(info, value=v, unit=UNIT)
            except Exception as e:   # the reference binary is test infrastructure; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        if rank == 0:
            print(json.dumps(line))
    finally:
        if not args.keep:
            shutil.rmtree(work, ignore_errors=True)
        if dist is not None:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(reference_arm(a) if a.impl == "reference" else b200_arm(a))

#!/usr/bin/env python
"""bench.py -- MPC QP solves/s of the batched sm_100a engine (BASELINE.json metric).

    python bench.py [--config 1..5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sweep fma|mma]
    N > 1:  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (record -> assembled QP -> 12 contact forces) over one batch of synthetic
problems.  `--config` picks the BASELINE.json workload; the default, and what the driver runs, is
config 2: B=4096 independent robots per GPU, trot, horizon 10 (weak scaling; at N>1 the step ends with ONE gather of
the [N*4096, 12] forces -- north_star: "a single ... all-gather of solved forces only when the batch is split".  The
default gather is the engine's own: the copy engines push every rank's forces into all ranks' gather buffers over NVLink
peer mappings and a device-side flag barrier closes the step, on a communication stream of its own; `--gather nccl` is
all_gather_into_tensor, `--gather peer` the solve kernel's fused peer-store epilogue).
Configs 4 and 5 are one 65536-problem batch cut into N contiguous shards (sharding.shard_bounds, strong scaling);
config 1 is one robot through the reference's own C interface, one MPC tick at a time.

The line printed by rank 0 carries
  value     whole-job solves/s with the records resident in HBM (device entry of the C ABI, batches in flight on the
            engine's slots), `serial` the same with one batch at a time,
  e2e       the same through the host entry (pinned host records -> H2D -> kernels -> D2H forces),
  roofline  algorithmic bytes of the dominant kernel / its CUDA-event duration vs the measured HBM peak,
  parity    computed IN THIS RUN: 256 sampled problems of the workload against the CPU oracle (fp64 truth and the
            reference-faithful fp32 path); `value` counts a batch's problems as solved only in the proportion that
            came back optimal,
  cpu_baseline  the CPU oracle (reference qpOASES when oracle/_ref exists) on the box's host cores.
`--impl reference` times that CPU path alone (all host cores) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_qp_solves_per_sec"
UNIT = "solves/s"
L2_BYTES = 126e6

# BASELINE.json configs[0..4].  per_gpu: every rank solves `batch` problems (weak scaling); otherwise
# `batch` is the whole job and is cut into contiguous shards (strong scaling).
CONFIGS = {
    1: dict(key="config1", h=10, batch=1, per_gpu=True,
            text="config1: 1 robot, trot gait, horizon=10, plane terrain, through the reference's C interface "
                 "(setup_problem ... get_solution), one MPC tick at a time"),
    2: dict(key="config2", h=10, batch=4096, per_gpu=True, seed=1234,
            text="config2: B=%d independent robots per GPU, trot, horizon=10, 12 forces, 20 friction-cone rows per "
                 "step, seed 1234+"),
    3: dict(key="config3", h=20, batch=4096, per_gpu=True, seed=2345,
            text="config3: B=%d per GPU, horizon=20, mixed gaits 0-11, randomised body inertia and mass, seed 2345+"),
    4: dict(key="config4", h=10, batch=65536, per_gpu=False, seed=3456,
            text="config4: B=%d, horizon=10, trot with stairs-terrain foothold perturbations, one batch sharded over "
                 "the GPUs + gather of the forces, seed 3456+"),
    5: dict(key="config5", h=16, batch=65536, per_gpu=False, seed=4567,
            text="config5: B=%d, horizon=16, galloping, one batch sharded over the GPUs + gather of the forces, "
                 "seed 4567+"),
}


def workload_text(cfg, batch):
    """The one string both arms print as config.workload."""
    return cfg["text"] % batch if "%d" in cfg["text"] else cfg["text"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_static_profile(cfg_id, solver="riccati"):
    """Figures that come from committed ncu captures / microbenchmarks, NOT from this run (labelled static):
    DRAM bytes per launch of the dominant kernel, its pipe utilisation, the measured fp64 peaks."""
    out = {"static": True}
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    traffic = None
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        e = (d.get("inverse_solver_config%d" % cfg_id) if solver == "inverse" else None) or d.get("config%d" % cfg_id)
        if e:
            traffic = e.get("dram_bytes_per_launch")
            out.update({k: e.get(k) for k in ("fp64_pipe_pct_of_peak", "issue_slots_pct_of_peak",
                                              "tensor_pipe_pct_of_peak", "source") if e.get(k) is not None})
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        with open(p) as f:
            out["fp64_peaks_measured"] = json.load(f)
    return traffic, out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker and the reported baseline, never the product path)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(records, horizon, target_seconds=12.0, workers=None):
    """Times the CPU oracle (reference qpOASES + fp32 assembly restatement) over all host cores.
    Returns (solves_per_s, dict)."""
    from oracle import oracle as O
    O.lib()
    kind = "reference" if O.have_reference_qpoases() else "port"
    backend = "reference" if kind == "reference" else "port"
    workers = workers or os.cpu_count() or 1
    # calibrate on a small slice, then size the sample for ~target_seconds of CPU work across the pool
    t0 = time.perf_counter()
    O.solve_batch(records[:64], horizon, 32, backend)
    per = (time.perf_counter() - t0) / 64
    n = int(min(max(target_seconds / per, 256), 65536))
    reps = (n + records.shape[0] - 1) // records.shape[0]
    sample = np.concatenate([records] * reps, 0)[:n] if reps > 1 else records[:n]
    t0 = time.perf_counter()
    O.solve_batch_parallel(sample, horizon, 32, backend, workers=workers)
    dt = time.perf_counter() - t0
    return n / dt, dict(kind=kind, cores=workers, sample="%d problems of the same workload, fp32 assembly + %s, "
                        "%d worker processes, %.1f s wall" % (n, "reference qpOASES 3.2 (oracle/_ref)" if kind == "reference"
                                                             else "oracle active-set port", workers, dt),
                        single_core_us_per_solve=per * 1e6)


def build_legacy_stub(oracle_side):
    """g++-compiles tools/legacy_tick_bench.cpp against the product library (include/convexMPC_interface.h) or, with
    -DLEGACY_ORACLE, against oracle/liboracle.so.  Returns (exe, extra argv)."""
    src = os.path.join(ROOT, "tools", "legacy_tick_bench.cpp")
    exe = os.path.join(tempfile.gettempdir(), "legacy_tick_bench_%s_%d" % ("oracle" if oracle_side else "gpu", os.getpid()))
    if oracle_side:
        from oracle import oracle as O
        O.lib()
        odir = os.path.join(ROOT, "oracle")
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-DLEGACY_ORACLE", src, "-o", exe, "-L", odir,
                               "-l:liboracle.so", "-Wl,-rpath," + odir, "-ldl"])
        return exe, [os.path.join(odir, "_ref", "libqpoases_ref.so")]
    from quadruped_ctrl_b200 import engine as E
    libdir = os.path.dirname(E.LIB_PATH)
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           src, "-o", exe, "-L", libdir, "-l:" + os.path.basename(E.LIB_PATH), "-Wl,-rpath," + libdir])
    return exe, []


def run_legacy_stub(exe, extra, ticks, h=10):
    """-> (median us, p95 us, mean us, forces [h, 12])."""
    try:
        r = subprocess.run([exe, str(ticks), str(h)] + extra, capture_output=True, text=True, timeout=1200)
    finally:
        os.unlink(exe)
    if r.returncode != 0:
        raise SystemExit("legacy_tick_bench failed: " + r.stderr[-500:])
    lines = r.stdout.splitlines()
    med, p95, mean = [float(x) for x in lines[0].split()[1:4]]
    forces = np.array([[float(x) for x in l.split()[3:15]] for l in lines[1:1 + h]])
    return med, p95, mean, forces, r.stderr


def parity_block(eng, rec, h, n_sample=256):
    """256 sampled problems of the workload, solved by the engine (host entry, whole 12h solution) and by the CPU
    oracle in both precisions; SURVEY 8d criterion.  Runs inside bench.py so that every reported number sits beside
    the parity of the very build that produced it."""
    from oracle import oracle as O
    from quadruped_ctrl_b200 import engine as E
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(rec.shape[0], size=min(n_sample, rec.shape[0]), replace=False))
    sub = np.ascontiguousarray(rec[idx])
    forces, sol, status = eng.solve_host(sub, want_solution=True)
    backend = O.default_backend()
    o64 = O.solve_batch(sub, h, 64, backend)
    o32 = O.solve_batch(sub, h, 32, backend)

    def rel(a, b):
        return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1.0)

    code = E.status_code(status)
    ok64 = o64["rc"] == 0
    e64 = rel(sol, o64["sol"])
    cloud = rel(o32["forces"], o64["forces"])
    e32 = rel(forces.astype(np.float64), o32["forces"])
    well = ok64 & (o32["rc"] == 0) & (cloud <= 1e-5)   # the reference's own fp32 answer is within 1e-5
    passed = (code == 0) & np.where(well, e32 <= 1e-4, True) & np.where(ok64, e64 <= 1e-9, True)
    return {
        "sampled": int(len(idx)), "oracle_backend": backend,
        "status_optimal": int((code == 0).sum()),
        "vs_oracle64_max_rel_12h": float(e64[ok64].max()) if ok64.any() else None,
        "vs_oracle32_max_rel_forces_well_conditioned": float(e32[well].max()) if well.any() else None,
        "well_conditioned": int(well.sum()),
        "reference_fp32_cloud_max": float(cloud[ok64].max()) if ok64.any() else None,
        "vs_oracle32_max_rel_forces_all": float(e32[ok64].max()) if ok64.any() else None,
        "reference_failed_nwsr": int((~ok64).sum()),
        "passed": int(passed.sum()),
        "criterion": "status optimal; 12h solution within 1e-9 of reference qpOASES on the fp64-assembled QP; first-step "
                     "forces within 1e-4 of the reference-faithful fp32 path wherever that path is itself within 1e-5 "
                     "of the fp64 answer (elsewhere the fp32 path's own rounding cloud is larger than the target)",
    }


def count_classes(rec, h, classes):
    """Problems per size class, from the gait tables (the classify kernel's rule)."""
    go = 4 * (48 + 12 * h)
    gait = rec[:, go:go + 4 * h].astype(np.float32)
    fmax = rec.view(np.float32)[:, 46:47]
    ub = gait * fmax
    ns = (~((ub < 0.01) & (ub > -0.01))).sum(1)
    nv = 3 * ns
    caps = [c["nv_cap"] for c in classes]
    idx = np.zeros(len(nv), np.int64)
    for i in range(len(caps) - 1):
        idx += nv > caps[i]
    return np.bincount(idx, minlength=len(caps)), nv


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    h = cfg["h"]
    if args.config == 1:
        exe, extra = build_legacy_stub(True)
        ticks = max(200, 50 * (args.steps + args.warmup))
        med, p95, mean, _, _ = run_legacy_stub(exe, extra, ticks)
        from oracle import oracle as O
        kind = "reference" if O.have_reference_qpoases() else "port"
        value = 1e6 / mean
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e-3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 assembly / f64 QP (reference CPU path)",
            "data": "synthetic", "config": {"workload": workload_text(cfg, 1),
                                            "note": "%d ticks, one after the other, 1 host core" % ticks},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "%d MPC ticks through the oracle's legacy C interface; median %.1f us, p95 %.1f us"
                                       % (ticks, med, p95)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return
    from quadruped_ctrl_b200 import workloads as W
    batch = args.batch or cfg["batch"]
    rec = W.CONFIGS[cfg["key"]](min(batch, 4096), h, cfg["seed"])
    vals = []
    info = None
    per_step_target = 4.0
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference_leg(rec, h, target_seconds=per_step_target * (os.cpu_count() or 1))
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * batch / value, "higher_is_better": True,
        "scaling": "weak" if cfg["per_gpu"] else "strong",
        "vs_baseline": None, "dtype": "f32 assembly / f64 QP (reference CPU path)", "data": "synthetic",
        "config": {"workload": workload_text(cfg, batch),
                   "note": "each step is a bounded sample of the workload on all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_config1(args):
    """One robot through the reference's C interface: every tick is a batch of one on the GPU (host classifies the
    problem, one kernel launch, one copy back)."""
    import torch
    from quadruped_ctrl_b200 import engine as E
    from quadruped_ctrl_b200 import records as R
    from quadruped_ctrl_b200 import workloads as W
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return   # replicas only: a single robot never leaves GPU 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU leg)")
    cfg = CONFIGS[1]
    h = 10
    ticks = max(500, 100 * args.steps)
    sampler = ClockSampler(0)
    sampler.start()
    exe, extra = build_legacy_stub(False)
    med, p95, mean, f_gpu, err = run_legacy_stub(exe, extra, ticks)
    clocks = sampler.stop()
    # parity of those very ticks: the ten gait phases of config 1 against the oracle
    from oracle import oracle as O
    rec = W.config1(h)
    o64 = O.solve_batch(rec, h, 64)
    o32 = O.solve_batch(rec, h, 32)
    # the stub's phase k is the gait table of iteration k, i.e. workloads.config1 record k
    e64 = np.linalg.norm(f_gpu - o64["forces"], axis=1) / np.maximum(np.linalg.norm(o64["forces"], axis=1), 1.0)
    e32 = np.linalg.norm(f_gpu - o32["forces"], axis=1) / np.maximum(np.linalg.norm(o32["forces"], axis=1), 1.0)
    cloud = np.linalg.norm(o32["forces"] - o64["forces"], axis=1) / np.maximum(np.linalg.norm(o64["forces"], axis=1), 1.0)
    # device-resident figure: the same problem as a batch of one through the batched entry
    eng = E.MpcBatch(h, 1, 0)
    eng.set_sweep_variant(args.sweep)
    eng.set_solver(args.solver)
    d = torch.from_numpy(rec[:1]).cuda()
    f = torch.empty((1, 12), dtype=torch.float32, device="cuda")
    st = torch.empty((1,), dtype=torch.int32, device="cuda")
    for _ in range(20):
        eng.solve_device(d, forces=f, status=st)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.kernel_launches()
    n = 200
    ev0.record()
    for _ in range(n):
        eng.solve_device(d, forces=f, status=st)
    ev1.record()
    torch.cuda.synchronize()
    dev_us = ev0.elapsed_time(ev1) * 1e3 / n
    launches = (eng.kernel_launches() - l0) // n
    eng.set_timing(True)
    eng.solve_device(d, forces=f, status=st)
    torch.cuda.synchronize()
    k_ms = max(eng.last_class_kernel_ms(c) for c in range(len(eng.classes())))
    peak, peak_src = load_peaks()
    traffic, on_chip = load_static_profile(1, args.solver)
    achieved = R.algorithmic_bytes(h) / (k_ms * 1e-3) / 1e9
    exe_o, extra_o = build_legacy_stub(True)
    cmed, cp95, cmean, _, _ = run_legacy_stub(exe_o, extra_o, 400)
    line = {
        "metric": METRIC, "value": 1e6 / dev_us, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_us * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_text(cfg, 1),
                   "note": "value: batch of one through mpc_batch_solve_device, back to back (launch-bound); e2e: %d "
                           "ticks through the legacy C interface from a C++ caller (tools/legacy_tick_bench.cpp)" % ticks,
                   "l2": "latency benchmark of a single problem: inputs are 720 bytes, L2 state is irrelevant",
                   "solver": args.solver, "sweep": args.sweep},
        "e2e": {"value": 1e6 / mean, "unit": UNIT, "h2d_bytes_per_step": R.record_stride(h),
                "d2h_bytes_per_step": 12 * 4 + 4 + 12 * h * 8, "us_per_tick_median": med, "us_per_tick_p95": p95,
                "api": "setup_problem / update_x_drag / update_solver_settings / update_problem_data_floats / "
                       "12 x get_solution (ConvexMPCLocomotion.cpp:630-674)"},
        "gpu_launches": int(launches * args.steps),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "on_chip": on_chip, "kernel_ms": k_ms,
                     "kernel": "the problem's size-class kernel, one CTA",
                     "algorithmic_bytes_per_solve": R.algorithmic_bytes(h),
                     "note": "one problem on one SM: latency-bound by construction"},
        "parity": {"sampled": 10, "vs_oracle64_max_rel_forces": float(e64.max()),
                   "vs_oracle32_max_rel_forces": float(e32.max()), "reference_fp32_cloud_max": float(cloud.max()),
                   "passed": int(((e64 <= 1e-6) & (e32 <= 1e-4)).sum()),
                   "criterion": "forces of the ten gait phases returned by get_solution(0..11): within 1e-4 of the "
                                "reference-faithful fp32 path and 1e-6 (float output) of the fp64 answer"},
        "clocks": clocks,
        "cpu_baseline": {"value": 1e6 / cmean, "unit": UNIT, "cores": 1, "kind": "reference" if O.have_reference_qpoases() else "port",
                         "sample": "400 MPC ticks through the oracle's legacy C interface (fp32 assembly + qpOASES), "
                                   "median %.1f us, p95 %.1f us" % (cmed, cp95),
                         "single_core_us_per_solve": cmean},
    }
    print(json.dumps(line))
    eng.close()


def run_ours(args):
    if args.config == 1:
        return run_config1(args)
    import torch
    import torch.distributed as dist
    from quadruped_ctrl_b200 import engine as E
    from quadruped_ctrl_b200 import records as R
    from quadruped_ctrl_b200 import sharding as S
    from quadruped_ctrl_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    h = cfg["h"]
    total = args.batch or cfg["batch"]
    if cfg["per_gpu"]:
        B, lo = total, rank * total            # every rank its own batch (weak scaling)
        job = world * total
    else:
        lo, hi = S.shard_bounds(total, world, rank)   # one batch, contiguous shards (strong scaling)
        if (hi - lo) * world != total:
            raise SystemExit("bench.py: --batch must divide by the number of GPUs for the sharded configs")
        B, job = hi - lo, total
    gen = W.CONFIGS[cfg["key"]]
    eng = E.MpcBatch(h, B, local_rank)
    eng.set_sweep_variant(args.sweep)
    eng.set_solver(args.solver)
    classes = eng.classes()
    stride = eng.stride
    # ---- synthetic inputs: distinct batches per rank, rotated so that every step reads cold (non-L2) inputs ----
    n_sets = max(3, min(96, int(np.ceil(2.2 * L2_BYTES / (B * stride)))))
    host_sets = [gen(B, h, cfg["seed"] + 1000 * rank + i) for i in range(n_sets)]
    dev_sets = [torch.from_numpy(s).to(dev) for s in host_sets]
    per_class, _ = count_classes(host_sets[0], h, classes)
    # Consecutive steps are independent batches: they rotate over the engine's scratch slots / streams, so the tail
    # of step i (a few CTAs still solving) shares the GPU with the head of step i+1.
    nq = max(1, min(args.inflight, E.SLOTS))
    streams = [torch.cuda.Stream(dev) for _ in range(nq)]
    peer = world > 1 and args.gather == "peer"
    # Output buffers rotate over MORE sets than there are scratch slots when the batch is gathered over NCCL: a solve
    # then waits for the gather that read its output set 2*nq steps ago instead of nq -- the ranks are separate
    # processes whose launch jitter otherwise stalls every slot on the slowest rank's previous gather.
    push = world > 1 and args.gather == "push"
    nbuf = nq * (args.outbufs if (world > 1 and not peer and not push) else 1)
    if push:   # the engine has one gather region per scratch slot (six): a rank may run that many batches ahead of the
        nbuf = max(nq, min(E.SLOTS, nq * args.pushbufs))   # slowest rank's gather instead of nq

    forces2 = [torch.empty((B, 12), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    status2 = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    gathered2 = [torch.empty((world * B, 12), dtype=torch.float32, device=dev) for _ in range(nbuf)] if world > 1 else None
    if peer or push:
        # peer: the solve kernel stores every force straight into all ranks' gather buffers over NVLink (fused epilogue);
        # push: the copy engines do it after the solve (no SM takes part) -- both end with a device-side flag barrier
        try:
            eng.setup_peer_gather(world * B, rank * B)
            ipc_ok = 1
        except E.MpcError as exc:      # no CUDA IPC between the ranks (container restrictions): NCCL gather instead
            print("bench.py: peer mappings unavailable (%s) -- falling back to --gather nccl" % exc, file=sys.stderr)
            ipc_ok = 0
        flag = torch.tensor([ipc_ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # every rank takes the same path
        if int(flag.item()) == 1:
            eng.set_gather_fused(peer)
            gathered2 = eng.gather_views   # one region per scratch slot
        else:
            peer = push = False
            args.gather = "nccl"
            nbuf = nq * args.outbufs
            forces2 = [torch.empty((B, 12), dtype=torch.float32, device=dev) for _ in range(nbuf)]
            status2 = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(nbuf)]
            gathered2 = [torch.empty((world * B, 12), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    # NCCL gather on a stream of its own: the next batch's kernels never queue behind the collective
    nogather = world > 1 and args.gather == "none"   # diagnosis only: what the step costs without its gather
    comm = torch.cuda.Stream(dev, priority=-1 if args.comm_priority else 0) if world > 1 and not peer and not nogather else None
    gather_done = [torch.cuda.Event() for _ in range(nbuf)]
    tiny_gather = bool(os.environ.get("MPC_DIAG_TINY_GATHER"))
    if os.environ.get("MPC_DIAG_CTAS_PER_SM"):
        eng.set_ctas_per_sm_limit(int(os.environ["MPC_DIAG_CTAS_PER_SM"]))

    def step(i, overlap=True):
        q = (i % nq) if overlap else 0
        j = (i % nbuf) if overlap else 0
        st = streams[q]
        with torch.cuda.stream(st):
            if comm is not None:
                st.wait_event(gather_done[j])   # the previous gather out of this output set has finished
            eng.solve_device(dev_sets[i % n_sets], forces=forces2[j], status=status2[j], stream=st, slot=q)
            if world > 1 and not nogather:
                if peer:
                    eng.gather_sync(stream=st, slot=q)  # device-side flag exchange over NVLink
                else:
                    comm.wait_stream(st)
                    with torch.cuda.stream(comm):
                        if push:
                            eng.gather_push(forces2[j], slot=j, stream=comm)
                        elif tiny_gather:   # diagnosis only (MPC_DIAG_TINY_GATHER): the rendezvous without the payload
                            dist.all_gather_into_tensor(gathered2[j].view(-1)[:2 * world], forces2[j].view(-1)[:2])
                        else:
                            dist.all_gather_into_tensor(gathered2[j], forces2[j])
                        gather_done[j].record(comm)
                    if not overlap:
                        st.wait_stream(comm)

    def fork():
        cur = torch.cuda.current_stream(dev)
        for st in streams + ([comm] if comm is not None else []):
            st.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream(dev)
        for st in streams + ([comm] if comm is not None else []):
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(n, first, overlap):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        eng.timing_mark()
        ev0.record()
        fork()
        h0 = time.perf_counter()
        for i in range(n):
            step(first + i, overlap)
        host_ms[0] = (time.perf_counter() - h0) * 1e3 / n   # host time to QUEUE one step (no synchronisation inside)
        join()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fork()
    for i in range(args.warmup):
        step(i)
    join()
    barrier()
    # which size class dominates this workload: every class kernel timed alone (events around each, one batch at a
    # time); the timed runs below then carry events around that kernel only
    eng.set_timing(True)
    eng.timing_mark()
    for i in range(4):
        step(i, False)
    join()
    barrier()
    k_alone = [eng.timing_collect(c) for c in range(len(classes))]
    # (among the classes the host counts problems in: the overflow class only receives re-queued problems, a handful
    # of CTAs whose latency says nothing about the batch)
    dominant = int(np.argmax([ms if per_class[c] > 0 else -1.0 for c, (ms, _) in enumerate(k_alone)]))
    eng.set_timed_class(dominant)

    def optimal_fraction():
        bad = 0
        for st_ in status2:
            bad += int(((st_ & 0xff) != 0).sum().item())
        return 1.0 - bad / float(len(status2) * B)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.kernel_launches()
    total_ms = timed(args.steps, args.warmup, True)
    host_queue_ms = host_ms[0]
    launches = eng.kernel_launches() - launches0
    k_timed = eng.timing_collect(dominant)
    frac_ok = optimal_fraction()   # statuses of the LAST batch on every slot, i.e. of timed steps
    value = frac_ok * job * args.steps / (total_ms * 1e-3)
    # ---- the same steps one batch at a time on one stream: latency-style figure and the kernels timed alone ----
    n_serial = min(args.steps, 32)
    serial_ms = timed(n_serial, args.warmup, False)
    serial_value = frac_ok * job * n_serial / (serial_ms * 1e-3)
    k_ms, k_n = eng.timing_collect(dominant)
    # ---- check of what was timed: the last step's forces against a fresh solve of the same batch, the gather against
    #      its inputs (exact integer checksum), every status optimal ----
    last = args.warmup + n_serial - 1
    ref_f, _, ref_st = eng.solve_device(dev_sets[last % n_sets])
    torch.cuda.synchronize()
    results_ok = bool(torch.equal(ref_f, forces2[0])) and frac_ok == 1.0
    gather_ok = None
    if world > 1 and not nogather:
        mine = forces2[0].view(torch.int32).to(torch.int64).sum()
        tot = mine.clone()
        dist.all_reduce(tot)
        g = gathered2[0]
        gather_ok = bool(torch.equal(g[rank * B:(rank + 1) * B], forces2[0])) and \
            int(g.view(torch.int32).to(torch.int64).sum().item()) == int(tot.item())
        flag = torch.tensor([int(gather_ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(flag.item())

    # ---- end to end through the host entry: pinned host records -> H2D -> kernels -> D2H, every step.
    # The slotted host API is used the way a caller with a stream of batches uses it: submit step i, then collect step
    # i-2.  Every step's inputs come from pinned host memory and its forces + status are read back inside the timed
    # region; at N>1 the gather takes the slot's device forces (no second upload).
    pinned_sets = [torch.from_numpy(s).pin_memory() for s in host_sets[:min(8, n_sets)]]
    nslots = max(2, min(args.e2e_slots, E.SLOTS))
    depth = nslots - 1
    out_f = [eng.host_buffers(q)[1] for q in range(nslots)]
    out_s = [eng.host_buffers(q)[3] for q in range(nslots)]
    dev_f = [eng.device_forces(q) for q in range(nslots)] if world > 1 else None
    checksum = [0.0]
    e2e_bad = [0]

    e2e_gathered = [None] * nslots   # event: the gather out of the slot's device forces has finished
    dev_f_B = [t[:B] for t in dev_f] if dev_f is not None else None
    e2e_comm = comm if comm is not None else (torch.cuda.Stream(dev) if world > 1 else None)

    def collect(slot):
        eng.wait_host(slot)                       # results stay in the slot's pinned buffers
        checksum[0] += float(out_f[slot][0, 2])
        e2e_bad[0] += int((out_s[slot][:B] & 0xff != 0).sum())
        if world > 1 and not nogather:            # the forces are complete on the device (wait_host synchronised)
            if e2e_gathered[slot] is None:
                e2e_gathered[slot] = torch.cuda.Event()
            if push:
                eng.gather_push(dev_f_B[slot], slot=slot, stream=e2e_comm)
            else:
                with torch.cuda.stream(e2e_comm):
                    dist.all_gather_into_tensor(gathered2[0], dev_f_B[slot])
            e2e_gathered[slot].record(e2e_comm)

    def e2e_run(n):
        for i in range(n):
            if e2e_gathered[i % nslots] is not None:
                e2e_gathered[i % nslots].synchronize()   # (long done) before the slot's device forces are overwritten
            eng.submit_host(i % nslots, pinned_sets[i % len(pinned_sets)].numpy(), zero_copy=True)
            if i >= depth:
                collect((i - depth) % nslots)
        for j in range(max(0, n - depth), n):
            collect(j % nslots)
        torch.cuda.synchronize()

    eng.set_timing(False)
    e2e_run(max(args.warmup, 2 * len(pinned_sets) + 4))   # every pinned set has been through the DMA path once
    e2e_bad[0] = 0
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = (1.0 - e2e_bad[0] / float(B * args.steps)) * job * args.steps / float(t.item())
    # ---- the same from TICK records (SURVEY 8f N1 + N2 on the host path): 272 bytes per robot cross the bus and the
    #      problem records are built on the device.  Trot configs, N = 1 (synthetic trot ticks of the same shape). ----
    e2e_ticks = None
    if world == 1 and args.config in (2, 4):
        tick_gen = W.config2_ticks if args.config == 2 else W.config4_ticks
        tick_sets = [torch.from_numpy(tick_gen(B, h, cfg["seed"] + i)).pin_memory() for i in range(4)]
        tick_bad = [0]

        def ticks_run(n):
            for i in range(n):
                eng.submit_host_ticks(i % nslots, tick_sets[i % len(tick_sets)].numpy(), zero_copy=True)
                if i >= depth:
                    q = (i - depth) % nslots
                    eng.wait_host(q)
                    tick_bad[0] += int((out_s[q][:B] & 0xff != 0).sum())
            for j in range(max(0, n - depth), n):
                eng.wait_host(j % nslots)
                tick_bad[0] += int((out_s[j % nslots][:B] & 0xff != 0).sum())
            torch.cuda.synchronize()

        ticks_run(max(args.warmup, 8))
        tick_bad[0] = 0
        t0 = time.perf_counter()
        ticks_run(args.steps)
        dt_ticks = time.perf_counter() - t0
        e2e_ticks = {"value": (1.0 - tick_bad[0] / float(B * args.steps)) * B * args.steps / dt_ticks, "unit": UNIT,
                     "h2d_bytes_per_step": B * 272, "d2h_bytes_per_step": B * 48 + B * 4 + B * 16,
                     "api": "mpc_batch_submit_host_ticks / mpc_batch_wait_host: tick records in, problem records built "
                            "on the device (workloads.config%d_ticks: the same seeded problems as tick records)" % args.config}
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = load_peaks()
        traffic, on_chip = load_static_profile(args.config, args.solver)
        n_dom = int(per_class[dominant])
        alg_bytes = R.algorithmic_bytes(h) * max(n_dom, 1)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        nv_dom = min(classes[dominant]["nv_cap"], 12 * h)
        ric_dom = classes[dominant]["threads"] == 32   # the dominant class runs the Riccati kernel (one warp per problem)
        flops = R.algorithmic_flops_riccati(h, nv_dom) if ric_dom else R.algorithmic_flops(h, nv_dom)
        on_chip.update(algorithmic_fp64_flops_per_solve=flops, flop_model="riccati" if ric_dom else "inverse",
                       fp64_tflops_achieved=flops * n_dom / (k_ms * 1e-3) / 1e12)
        parity = parity_block(eng, host_sets[0], h)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if cfg["per_gpu"] else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(cfg, total), "problems_per_gpu": B,
                       "l2": "inputs rotate over %d distinct record sets (%.0f MB > 126 MB L2), no flush needed"
                             % (n_sets, n_sets * B * stride / 1e6),
                       "pipelining": "steps rotate over %d of the engine's scratch slots / streams (independent "
                                     "batches); ms_per_step is therefore an inverse throughput, `serial` is the "
                                     "one-batch-at-a-time figure" % nq,
                       "collective": ("none (N=1)" if world == 1 else
                                      "peer stores from the solve kernel + device-side flag barrier (no NCCL on the path)" if peer else
                                      "copy-engine gather: one DMA per peer of this rank's [B,12] forces over NVLink into every rank's gather "
                                      "buffer + device-side flag barrier, on a communication stream of its own (no SM takes part)" if push else
                                      "all_gather_into_tensor of [N*B,12] fp32 forces on a communication stream of its own"),
                       "solver": eng.solver(), "sweep": eng.sweep_variant(), "classes": classes,
                       "problems_per_class": [int(x) for x in per_class]},
            "serial": {"value": serial_value, "unit": UNIT, "ms_per_batch": serial_ms / n_serial, "steps": n_serial,
                       "note": "one batch at a time on one stream (BASELINE configs are single batches)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * stride,
                    "d2h_bytes_per_step": B * 48 + B * 4,
                    "api": "mpc_batch_submit_host / mpc_batch_wait_host (%d slots, %d batches submitted ahead; "
                           "page-locked caller buffers handed over zero-copy with mpc_batch_submit_host_pinned)" % (nslots, depth)},
            "gpu_launches": launches, "host_queue_ms_per_step": host_queue_ms,
            "results_ok": results_ok, "optimal_fraction": frac_ok,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "on_chip": on_chip,
                         "kernel": "size class %d (nv <= %d, %d threads): %d of the batch's %d problems"
                                   % (dominant, classes[dominant]["nv_cap"], classes[dominant]["threads"], n_dom, B),
                         "kernel_ms": k_ms, "kernel_launches_timed": k_n,
                         "kernel_ms_in_timed_region": k_timed[0],
                         "class_kernel_ms_alone": [ms for ms, _ in k_alone],
                         "algorithmic_bytes_per_solve": R.algorithmic_bytes(h),
                         "note": "on-chip fp64/latency bound by construction: H and g never leave shared memory, so "
                                 "the HBM fraction is tiny; see DESIGN.md for the fp64-pipe figures"},
            "parity": parity,
            "clocks": clocks,
        }
        if gather_ok is not None:
            line["gather_ok"] = gather_ok
        if e2e_ticks is not None:
            line["e2e_ticks"] = e2e_ticks
        if world == 1 and not args.no_cpu:
            v, info = cpu_reference_leg(host_sets[0][:4096], h, target_seconds=1.5 * (os.cpu_count() or 1))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"],
                                    "single_core_us_per_solve": info["single_core_us_per_solve"]}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json workload (1..5)")
    ap.add_argument("--batch", type=int, default=0, help="override the config's batch size")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("MPC_INFLIGHT", "3")), help="batches in flight on the device-resident path (<= 3)")
    ap.add_argument("--e2e-slots", type=int, default=6, help="scratch slots the end-to-end leg rotates over (<= 6)")
    ap.add_argument("--sweep", default=os.environ.get("MPC_SWEEP", "fma"), choices=["fma", "mma"],
                    help="inversion of the register-resident classes: FP64 FMA pipe or FP64 tensor pipe (DMMA)")
    ap.add_argument("--solver", default=os.environ.get("MPC_SOLVER", "riccati"), choices=["riccati", "inverse"],
                    help="riccati: sweeps over the horizon, no condensed Hessian (default); inverse: explicit inverse of "
                         "the reduced condensed Hessian")
    ap.add_argument("--pushbufs", type=int, default=2, help="N>1, push gather: output sets / gather regions per scratch slot")
    ap.add_argument("--outbufs", type=int, default=1, help="N>1, NCCL gather: output sets per scratch slot")
    ap.add_argument("--comm-priority", type=int, default=1, help="N>1: 1 = the NCCL gather runs on a high-priority stream")
    ap.add_argument("--gather", default="push", choices=["push", "nccl", "peer", "none"],
                    help="N>1: push = copy-engine gather over NVLink peer mappings + device-side flag barrier (default); "
                         "nccl = all_gather_into_tensor; peer = the solve kernel's fused peer-store epilogue; none = diagnosis")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

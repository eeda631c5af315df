#!/usr/bin/env python
"""bench.py -- MPC QP solves/s of the batched sm_100a engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    N > 1:  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (record -> assembled QP -> 12 contact forces) over one batch of
synthetic problems.  Workload at N=1: BASELINE.json configs[1] -- B=4096 independent robots, trot,
horizon 10 (workloads.config2, seed 1234).  At N>1 every rank solves its own 4096-problem shard (weak
scaling) and the step ends with the all-gather of the [N*4096, 12] forces (north_star: "a single NCCL
all-gather of solved forces only when the batch is split").

The line printed by rank 0 carries
  value     whole-job solves/s with the records resident in HBM (device entry of the C ABI),
  e2e       the same through the host entry (pinned host records -> H2D -> kernels -> D2H forces),
  roofline  algorithmic bytes of the dominant kernel / its CUDA-event duration vs the measured HBM peak,
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
HORIZON = 10
BATCH = 4096
N_SETS = 96          # distinct record sets rotated through so that every step reads cold (non-L2) inputs
L2_BYTES = 126e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """(dram bytes per launch, on-chip pipe figures) of the dominant kernel from the committed ncu capture
    (profiles/roofline_traffic.json), or (None, None)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), {"fp64_pipe_pct_of_peak": d.get("fp64_pipe_pct_of_peak"),
                                                "issue_slots_pct_of_peak": d.get("issue_slots_pct_of_peak"),
                                                "source": d.get("source")}
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from quadruped_ctrl_b200 import workloads as W
    rec = W.config2(args.batch, HORIZON, 1234)
    vals = []
    info = None
    per_step_target = 4.0
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference_leg(rec, HORIZON, target_seconds=per_step_target * (os.cpu_count() or 1))
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 assembly / f64 QP (reference CPU path)", "data": "synthetic",
        "config": {"workload": "config2: B=%d independent robots, trot, horizon=%d, seed 1234" % (args.batch, HORIZON),
                   "note": "each step is a bounded sample of the workload on all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from quadruped_ctrl_b200 import engine as E
    from quadruped_ctrl_b200 import records as R
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

    B, h = args.batch, HORIZON
    eng = E.MpcBatch(h, B, local_rank)
    eng.set_timed_class(0)   # CUDA events around the dominant kernel only (size class 0 holds every trot problem)
    classes = eng.classes()
    stride = eng.stride
    # ---- synthetic inputs: N_SETS distinct batches per rank (rotation keeps every step's inputs out of L2) ----
    n_sets = max(2, min(N_SETS, int(np.ceil(2.2 * L2_BYTES / (B * stride)))))
    host_sets = [W.config2(B, h, 1234 + 1000 * rank + i) for i in range(n_sets)]
    dev_sets = [torch.from_numpy(s).to(dev) for s in host_sets]
    # Consecutive steps are independent batches, so they alternate between two of the engine's scratch slots on two
    # streams: the tail of step i (a few CTAs still solving) shares the GPU with the head of step i+1.
    nq = max(1, min(args.inflight, E.SLOTS))   # batches in flight on the device-resident path
    streams = [torch.cuda.Stream(dev) for _ in range(nq)]
    forces2 = [torch.empty((B, 12), dtype=torch.float32, device=dev) for _ in range(nq)]
    status2 = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(nq)]
    forces, status = forces2[0], status2[0]
    gathered2 = [torch.empty((world * B, 12), dtype=torch.float32, device=dev) for _ in range(nq)] if world > 1 else None
    gathered = gathered2[0] if world > 1 else None
    peer = world > 1 and args.gather == "peer"
    if peer:  # fused: the solve kernel stores every force straight into all ranks' gather buffers over NVLink
        gathered = eng.setup_peer_gather(world * B, rank * B)
        gathered2 = eng.gather_views   # one region per scratch slot
    overlap = not args.serial

    def step(i):
        q = (i % nq) if overlap else 0
        if overlap:
            with torch.cuda.stream(streams[q]):
                eng.solve_device(dev_sets[i % n_sets], forces=forces2[q], status=status2[q], stream=streams[q], slot=q)
                if world > 1:
                    if peer:
                        eng.gather_sync(stream=streams[q], slot=q)  # device-side flag exchange over NVLink
                    else:
                        dist.all_gather_into_tensor(gathered2[q], forces2[q])
            return
        eng.solve_device(dev_sets[i % n_sets], forces=forces, status=status)
        if world > 1:
            if peer:
                eng.gather_sync()  # device-side flag exchange over NVLink: everybody's stores have landed
            else:
                dist.all_gather_into_tensor(gathered, forces)

    def fork():
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fork()
    for i in range(args.warmup):
        step(i)
    join()
    barrier()
    # all problems must have solved to optimality before anything is timed
    for st_ in (status2 if overlap else status2[:1]):
        codes = (st_.cpu().numpy() & 0xff)
        assert (codes == 0).all(), "non-optimal status in warm-up: %s" % np.bincount(codes)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.kernel_launches()
    dominant = 0  # size class 0 (nv <= 60) holds every trot problem of this workload
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    eng.timing_mark()
    ev0.record()
    fork()
    for i in range(args.steps):
        step(args.warmup + i)
    join()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches() - launches0
    # mean duration of the dominant kernel over the timed steps (events recorded on the launching stream); with
    # overlapping steps two launches share the GPU, so each one's duration is longer than its share of the step
    k_ms_timed, k_n = eng.timing_collect(dominant)
    # the same kernel timed alone: the same steps once more, one at a time on one stream (roofline figure)
    k_ms = k_ms_timed
    if overlap:
        eng.timing_mark()
        for i in range(min(args.steps, 32)):
            eng.solve_device(dev_sets[(args.warmup + i) % n_sets], forces=forces, status=status)
        torch.cuda.synchronize()
        k_ms, _ = eng.timing_collect(dominant)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host entry: pinned host records -> H2D -> kernels -> D2H, every step.
    # The slotted host API is used the way a caller with a stream of batches uses it: submit step i, then
    # collect step i-2, so one step's transfers and the host's work overlap the other steps' kernels.  Every step's inputs come from
    # pinned host memory and every step's forces + status are read back to the host inside the timed region.
    pinned_sets = [torch.from_numpy(s).pin_memory() for s in host_sets[:8]]
    nslots = E.SLOTS
    depth = nslots - 1          # batches submitted ahead of the one being collected
    out_f = [eng.host_buffers(q)[1] for q in range(nslots)]
    out_s = [eng.host_buffers(q)[3] for q in range(nslots)]
    checksum = [0.0]

    def collect(slot):
        eng.wait_host(slot)                       # results stay in the slot's pinned buffers
        checksum[0] += float(out_f[slot][0, 2]) + float(out_s[slot][0])
        if world > 1:
            forces.copy_(torch.from_numpy(out_f[slot][:B]), non_blocking=True)
            dist.all_gather_into_tensor(gathered, forces)

    def e2e_run(n):
        for i in range(n):
            eng.submit_host(i % nslots, pinned_sets[i % len(pinned_sets)].numpy())
            if i >= depth:
                collect((i - depth) % nslots)
        for j in range(max(0, n - depth), n):
            collect(j % nslots)
        torch.cuda.synchronize()

    e2e_run(max(args.warmup, 2 * len(pinned_sets) + 4))   # every pinned set has been through the DMA path once
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = load_peaks()
        traffic, on_chip = load_traffic()
        alg_bytes = R.algorithmic_bytes(h) * B
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        if on_chip is not None:
            # secondary figure: fp64 rate of the dominant kernel against the fp64 peak measured on this pool's B200s
            # with tools/microbench/dfma.cu (36.8 TFLOP/s); every problem of this workload has nv = 60
            flops = R.algorithmic_flops(h, 60) * B
            on_chip.update(fp64_tflops_achieved=flops / (k_ms * 1e-3) / 1e12, fp64_tflops_peak_measured=36.8,
                           algorithmic_fp64_flops_per_solve=R.algorithmic_flops(h, 60))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: B=%d independent robots per GPU, trot, horizon=%d, 12 forces, 20 "
                                   "friction-cone rows per step, seed 1234+" % (B, h),
                       "l2": "inputs rotate over %d distinct record sets (%.0f MB > 126 MB L2), no flush needed"
                             % (n_sets, n_sets * B * stride / 1e6),
                       "pipelining": ("steps rotate over %d of the engine's scratch slots / streams (independent batches)" % nq
                                      if overlap else "one stream, steps strictly one after another"),
                       "collective": ("none (N=1)" if world == 1 else
                                      "peer stores from the solve kernel + device-side flag barrier (no NCCL on the path)" if peer else
                                      "all_gather_into_tensor of [N*B,12] fp32 forces"),
                       "classes": classes},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * stride,
                    "d2h_bytes_per_step": B * 48 + B * 4,
                    "api": "mpc_batch_submit_host / mpc_batch_wait_host (%d slots, %d batches submitted ahead; page-locked host buffers read in place)" % (nslots, depth)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "on_chip": on_chip,
                         "kernel": "mpc_solve_pipe_kernel<128,...> (size class nv<=60, two problems in flight per CTA)", "kernel_ms": k_ms, "kernel_launches_timed": k_n,
                         "kernel_ms_in_timed_region": k_ms_timed,
                         "algorithmic_bytes_per_solve": R.algorithmic_bytes(h),
                         "note": "on-chip fp64/latency bound by construction: H and g never leave shared memory, so "
                                 "the HBM fraction is tiny; see DESIGN.md for the fp64-pipe figures"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            v, info = cpu_reference_leg(host_sets[0], h, target_seconds=1.5 * (os.cpu_count() or 1))
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--serial", action="store_true", help="one stream, no overlap between consecutive steps")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight on the device-resident path (<= 3)")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "peer"],
                    help="N>1: NCCL all-gather of the forces (default) or the kernel's fused peer-store epilogue")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

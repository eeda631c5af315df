"""ctypes binding of libquadruped_mpc_b200.so (include/mpc_batch.h) -- the batched sm_100a MPC engine.

PyTorch is used for device memory and streams only; every solve runs the hand-written CUDA
kernels in csrc/.  There is no CPU implementation behind this module: importing works
anywhere (so host logic can be tested), but creating an engine without the built library or
without an sm_100 device raises.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import records as R

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPC_LIB_PATH: development aid (A/B runs of experimental builds of the same library); never a CPU path
LIB_PATH = os.environ.get("MPC_LIB_PATH") or os.path.join(_HERE, "libquadruped_mpc_b200.so")
_LIB = None

MPC_OK, MPC_E_ARG, MPC_E_CUDA, MPC_E_NOMEM, MPC_E_NODEVICE = 0, -1, -2, -3, -4
SLOTS = 6  # MPC_BATCH_SLOTS of include/mpc_batch.h: scratch slots per engine (batches that can be in flight)
STATUS_OPTIMAL, STATUS_MAX_ITER, STATUS_BAD_INPUT, STATUS_NOT_PD, STATUS_NO_STANCE = 0, 1, 2, 3, 4

# every symbol include/mpc_batch.h and include/convexMPC_interface.h declare
BATCH_SYMBOLS = ["mpc_record_stride", "mpc_record_gait_offset", "mpc_batch_create", "mpc_batch_destroy",
                 "mpc_batch_solve_device", "mpc_batch_solve_device_slot", "mpc_batch_solve_host", "mpc_batch_submit_host", "mpc_batch_submit_host_pinned", "mpc_batch_wait_host",
                 "mpc_batch_submit_host_ticks", "mpc_batch_host_state", "mpc_batch_slots",
                 "mpc_batch_assemble_device", "mpc_batch_build_records_device", "mpc_batch_solve_ticks_device",
                 "mpc_batch_gait_state_device", "mpc_batch_leg_commands_device",
                 "mpc_batch_set_gather_peers", "mpc_batch_gather_alloc", "mpc_batch_gather_connect",
                 "mpc_batch_gather_buffer", "mpc_batch_gather_buffer_slot", "mpc_batch_gather_sync", "mpc_batch_gather_sync_slot", "mpc_batch_set_gather_fused", "mpc_batch_gather_push_slot", "mpc_batch_set_max_iterations", "mpc_batch_set_warm_start", "mpc_batch_warm_stride", "mpc_batch_set_sweep_variant", "mpc_batch_sweep_variant", "mpc_batch_set_solver", "mpc_batch_solver", "mpc_batch_set_timing", "mpc_batch_set_timed_class", "mpc_batch_set_phase_clock_buffer", "mpc_batch_set_ctas_per_sm_limit",
                 "mpc_batch_num_classes", "mpc_batch_class_info", "mpc_batch_kernel_launches",
                 "mpc_batch_last_solve_kernel_ms", "mpc_batch_last_class_kernel_ms", "mpc_batch_timing_mark",
                 "mpc_batch_timing_collect", "mpc_batch_host_buffers", "mpc_batch_device_buffers",
                 "mpc_batch_last_error", "mpc_last_error", "mpc_batch_horizon"]
LEGACY_SYMBOLS = ["setup_problem", "update_problem_data", "update_problem_data_floats", "get_solution",
                  "update_solver_settings", "_Z13update_x_dragf", "mpc_last_status", "mpc_last_iterations",
                  "mpc_set_robot", "mpc_shutdown", "mpc_legacy_record", "mpc_jcqp_requested"]


class MpcError(RuntimeError):
    pass


def build(force=False):
    """Compiles csrc/ for sm_100a into libquadruped_mpc_b200.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    deps = [os.path.join(src, f) for f in ("mpc_engine.cu", "mpc_core.h", "mpc_riccati.h", "mpc_ticks.h", "mpc_legs.h", "convexMPC_interface.cpp",
                                           "Makefile")]
    deps += [os.path.join(_HERE, "..", "include", f) for f in ("mpc_batch.h", "convexMPC_interface.h")]
    stale = force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(map(os.path.getmtime, deps))
    if stale:
        subprocess.check_call(["make", "-C", src, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    """Loads the shared library and declares the C signatures.  Raises when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise MpcError("libquadruped_mpc_b200.so is not built (run __graft_entry__.build() or make -C "
                       "quadruped_ctrl_b200/csrc); this package has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, f32, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
    L.mpc_record_stride.argtypes = [i32]
    L.mpc_record_stride.restype = ctypes.c_size_t
    L.mpc_record_gait_offset.argtypes = [i32]
    L.mpc_record_gait_offset.restype = ctypes.c_size_t
    L.mpc_batch_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32]
    L.mpc_batch_destroy.argtypes = [vp]
    L.mpc_batch_destroy.restype = None
    L.mpc_batch_solve_device.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.mpc_batch_solve_device_slot.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp]
    L.mpc_batch_gather_buffer_slot.argtypes = [vp, i32]
    L.mpc_batch_gather_buffer_slot.restype = vp
    L.mpc_batch_gather_sync_slot.argtypes = [vp, i32, vp]
    L.mpc_batch_solve_host.argtypes = [vp, vp, i32, vp, vp, vp]
    L.mpc_batch_submit_host.argtypes = [vp, i32, vp, i32, i32]
    L.mpc_batch_submit_host_pinned.argtypes = [vp, i32, vp, i32, i32]
    L.mpc_batch_wait_host.argtypes = [vp, i32, vp, vp, vp]
    L.mpc_batch_submit_host_ticks.argtypes = [vp, i32, vp, i32, i32, i32]
    L.mpc_batch_host_state.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    if L.mpc_batch_slots() != SLOTS:
        raise MpcError("engine.SLOTS (%d) does not match the library's MPC_BATCH_SLOTS (%d)" % (SLOTS, L.mpc_batch_slots()))
    L.mpc_batch_assemble_device.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.mpc_batch_build_records_device.argtypes = [vp, vp, i32, vp, vp, vp]
    L.mpc_batch_solve_ticks_device.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.mpc_batch_gait_state_device.argtypes = [vp, vp, i32, vp, vp, i32, vp]
    L.mpc_batch_leg_commands_device.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.mpc_batch_set_gather_peers.argtypes = [vp, ctypes.POINTER(vp), i32, i32]
    L.mpc_batch_gather_alloc.argtypes = [vp, i32, vp]
    L.mpc_batch_gather_connect.argtypes = [vp, vp, i32, i32, i32]
    L.mpc_batch_gather_sync.argtypes = [vp, vp]
    L.mpc_batch_set_gather_fused.argtypes = [vp, i32]
    L.mpc_batch_gather_push_slot.argtypes = [vp, i32, vp, i32, vp]
    L.mpc_batch_gather_buffer.argtypes = [vp]
    L.mpc_batch_gather_buffer.restype = vp
    L.mpc_batch_set_max_iterations.argtypes = [vp, i32]
    L.mpc_batch_set_warm_start.argtypes = [vp, vp, vp, i32]
    L.mpc_batch_set_sweep_variant.argtypes = [vp, i32]
    L.mpc_batch_sweep_variant.argtypes = [vp]
    L.mpc_batch_set_solver.argtypes = [vp, i32]
    L.mpc_batch_solver.argtypes = [vp]
    L.mpc_batch_set_timing.argtypes = [vp, i32]
    L.mpc_batch_set_timed_class.argtypes = [vp, i32]
    L.mpc_batch_set_phase_clock_buffer.argtypes = [vp, vp]
    L.mpc_batch_set_ctas_per_sm_limit.argtypes = [vp, i32]
    L.mpc_batch_num_classes.argtypes = [vp]
    L.mpc_batch_class_info.argtypes = [vp, i32, ctypes.POINTER(i32)]
    L.mpc_batch_kernel_launches.argtypes = [vp]
    L.mpc_batch_kernel_launches.restype = ctypes.c_long
    L.mpc_batch_last_solve_kernel_ms.argtypes = [vp]
    L.mpc_batch_last_solve_kernel_ms.restype = f32
    L.mpc_batch_last_class_kernel_ms.argtypes = [vp, i32]
    L.mpc_batch_last_class_kernel_ms.restype = f32
    L.mpc_batch_timing_mark.argtypes = [vp]
    L.mpc_batch_timing_mark.restype = None
    L.mpc_batch_timing_collect.argtypes = [vp, i32, ctypes.POINTER(f32), ctypes.POINTER(i32)]
    L.mpc_batch_host_buffers.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                         ctypes.POINTER(vp)]
    L.mpc_batch_device_buffers.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.mpc_batch_last_error.argtypes = [vp]
    L.mpc_batch_last_error.restype = ctypes.c_char_p
    L.mpc_last_error.restype = ctypes.c_char_p
    L.mpc_batch_horizon.argtypes = [vp]
    # legacy interface (include/convexMPC_interface.h)
    fp, dp, ip = ctypes.POINTER(f32), ctypes.POINTER(f64), ctypes.POINTER(i32)
    L.setup_problem.argtypes = [f64, i32, f64, f64]
    L.setup_problem.restype = None
    L.update_problem_data_floats.argtypes = [fp, fp, fp, fp, fp, f32, fp, fp, f32, ip]
    L.update_problem_data_floats.restype = None
    L.update_problem_data.argtypes = [dp, dp, dp, dp, dp, f64, dp, dp, f64, ip]
    L.update_problem_data.restype = None
    L.update_solver_settings.argtypes = [i32, f64, f64, f64, f64, f64]
    L.update_solver_settings.restype = None
    L.get_solution.argtypes = [i32]
    L.get_solution.restype = f64
    L._Z13update_x_dragf.argtypes = [f32]
    L._Z13update_x_dragf.restype = None
    L.mpc_set_robot.argtypes = [fp, f32]
    L.mpc_set_robot.restype = None
    L.mpc_shutdown.restype = None
    L.mpc_legacy_record.argtypes = [vp]
    _LIB = L
    return L


def _torch():
    import torch
    return torch


class MpcBatch:
    """One engine per (device, horizon).  Problems are packed records (records.pack_records)."""

    def __init__(self, horizon, max_batch, device=0):
        L = lib()
        self._L = L
        self.horizon = int(horizon)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.stride = R.record_stride(self.horizon)
        assert self.stride == L.mpc_record_stride(self.horizon)
        h = ctypes.c_void_p()
        rc = L.mpc_batch_create(ctypes.byref(h), self.device, self.horizon, self.max_batch)
        if rc != MPC_OK:
            raise MpcError("mpc_batch_create failed (rc=%d): %s" % (rc, L.mpc_last_error().decode()))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.mpc_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != MPC_OK:
            raise MpcError("%s failed (rc=%d): %s" % (what, rc, self._L.mpc_batch_last_error(self._h).decode()))

    # ---- configuration ------------------------------------------------------------------
    def set_max_iterations(self, n):
        self._check(self._L.mpc_batch_set_max_iterations(self._h, int(n)), "set_max_iterations")

    def new_warm_cache(self, robots):
        """Zeroed device cache for `robots` robots (cuda int32 [robots, stride])."""
        torch = _torch()
        return torch.zeros((int(robots), int(self._L.mpc_batch_warm_stride())), dtype=torch.int32,
                           device=torch.device("cuda", self.device))

    def set_warm_start(self, cache, robot_ids=None, shift=1):
        """Warm start across ticks: cache from new_warm_cache() (None: off), robot_ids cuda int32 [B]
        (None: problem index), shift = horizon steps the gait advanced since the cached solve."""
        self._warm_keep = (cache, robot_ids)
        self._check(self._L.mpc_batch_set_warm_start(self._h, cache.data_ptr() if cache is not None else None,
                                                     robot_ids.data_ptr() if robot_ids is not None else None,
                                                     int(shift)), "set_warm_start")

    def set_sweep_variant(self, variant):
        """0 / "fma": rank-1 sweep on the FP64 FMA pipe; 1 / "mma": grouped sweep on the FP64 tensor pipe (DMMA)."""
        v = {"fma": 0, "mma": 1}.get(variant, variant)
        self._check(self._L.mpc_batch_set_sweep_variant(self._h, int(v)), "set_sweep_variant")

    def sweep_variant(self):
        return "mma" if self._L.mpc_batch_sweep_variant(self._h) == 1 else "fma"

    def set_solver(self, solver):
        """1 / "riccati" (default): Riccati sweeps, the condensed Hessian is never formed (csrc/mpc_riccati.h);
        0 / "inverse": explicit inverse of the reduced condensed Hessian (csrc/mpc_core.h)."""
        v = {"inverse": 0, "riccati": 1}.get(solver, solver)
        self._check(self._L.mpc_batch_set_solver(self._h, int(v)), "set_solver")

    def solver(self):
        return "riccati" if self._L.mpc_batch_solver(self._h) == 1 else "inverse"

    def set_timing(self, on):
        self._check(self._L.mpc_batch_set_timing(self._h, int(bool(on))), "set_timing")

    def set_phase_clock_buffer(self, tensor):
        """tensor: cuda int64 [max_batch, 24] (or None): per-problem clock64() stamps at phase boundaries."""
        self._phase_buf = tensor
        self._check(self._L.mpc_batch_set_phase_clock_buffer(self._h, tensor.data_ptr() if tensor is not None else None),
                    "set_phase_clock_buffer")

    def set_ctas_per_sm_limit(self, n):
        self._check(self._L.mpc_batch_set_ctas_per_sm_limit(self._h, int(n)), "set_ctas_per_sm_limit")

    def last_solve_kernel_ms(self):
        return float(self._L.mpc_batch_last_solve_kernel_ms(self._h))

    def set_timed_class(self, idx):
        self._check(self._L.mpc_batch_set_timed_class(self._h, int(idx)), "set_timed_class")

    def last_class_kernel_ms(self, idx):
        return float(self._L.mpc_batch_last_class_kernel_ms(self._h, int(idx)))

    def timing_mark(self):
        self._L.mpc_batch_timing_mark(self._h)

    def timing_collect(self, idx):
        """(mean ms, n) of size class idx's solve kernel over the solves since timing_mark() (last 256 at most)."""
        ms, n = ctypes.c_float(), ctypes.c_int()
        self._check(self._L.mpc_batch_timing_collect(self._h, int(idx), ctypes.byref(ms), ctypes.byref(n)),
                    "timing_collect")
        return float(ms.value), int(n.value)

    def host_buffers(self, slot=0):
        """numpy views of slot `slot`'s pinned staging buffers: (records [max_batch, stride] u8,
        forces [max_batch, 12] f32, solution [max_batch, 12h] f64, status [max_batch] i32).  Passing these to
        solve_host / submit_host / wait_host skips the pageable->pinned copies."""
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        self._check(self._L.mpc_batch_host_buffers(self._h, int(slot), *[ctypes.byref(p) for p in ptrs]),
                    "host_buffers")
        B, NU = self.max_batch, 12 * self.horizon

        def view(p, nbytes, dtype, shape):
            buf = (ctypes.c_char * nbytes).from_address(p.value)
            return np.frombuffer(buf, dtype=dtype).reshape(shape)

        return (view(ptrs[0], B * self.stride, np.uint8, (B, self.stride)),
                view(ptrs[1], B * 48, np.float32, (B, 12)),
                view(ptrs[2], B * NU * 8, np.float64, (B, NU)),
                view(ptrs[3], B * 4, np.int32, (B,)))

    def device_forces(self, slot=0):
        """Slot `slot`'s device forces buffer of the host entry as a cuda tensor [max_batch, 12] f32 (valid after
        wait_host(slot) until the next submit on the slot)."""
        torch = _torch()
        pf, ps = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self._L.mpc_batch_device_buffers(self._h, int(slot), ctypes.byref(pf), ctypes.byref(ps)),
                    "device_buffers")

        class _Buf:
            __cuda_array_interface__ = {"shape": (self.max_batch, 12), "typestr": "<f4", "data": (int(pf.value), False),
                                        "version": 2, "strides": None}
        keep = getattr(self, "_dev_keep", [])
        keep.append(_Buf())
        self._dev_keep = keep
        return torch.as_tensor(keep[-1], device=torch.device("cuda", self.device))

    def kernel_launches(self):
        return int(self._L.mpc_batch_kernel_launches(self._h))

    def classes(self):
        out = []
        info = (ctypes.c_int * 6)()
        for i in range(self._L.mpc_batch_num_classes(self._h)):
            self._L.mpc_batch_class_info(self._h, i, info)
            out.append(dict(nv_cap=info[0], m_cap=info[1], threads=info[2], grid=info[3], smem=info[4],
                            in_smem=bool(info[5])))
        return out

    def set_gather_peers(self, peer_ptrs, rank_offset):
        n = len(peer_ptrs)
        arr = (ctypes.c_void_p * max(n, 1))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
        self._check(self._L.mpc_batch_set_gather_peers(self._h, arr, n, int(rank_offset)), "set_gather_peers")

    def setup_peer_gather(self, world_batch, rank_offset, group=None):
        """Fused shard-and-gather: allocates this rank's [world_batch, 12] gather buffer, exchanges CUDA IPC
        handles over torch.distributed and arms the solve kernel's peer-store epilogue.  Returns the local
        gather buffer as a cuda tensor (valid after a solve on every rank + one barrier)."""
        torch = _torch()
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handle = (ctypes.c_char * 64)()
        self._check(self._L.mpc_batch_gather_alloc(self._h, int(world_batch), ctypes.addressof(handle)),
                    "mpc_batch_gather_alloc")
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).cuda(self.device)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        blob = allh.cpu().numpy().tobytes()
        self._check(self._L.mpc_batch_gather_connect(self._h, blob, world, rank, int(rank_offset)),
                    "mpc_batch_gather_connect")
        self._gather_keepalive = []
        views = []
        for slot in range(SLOTS):
            ptr = self._L.mpc_batch_gather_buffer_slot(self._h, slot)

            class _Buf:
                __cuda_array_interface__ = {"shape": (int(world_batch), 12), "typestr": "<f4",
                                            "data": (int(ptr), False), "version": 2, "strides": None}
            self._gather_keepalive.append(_Buf())
            views.append(torch.as_tensor(self._gather_keepalive[-1], device=torch.device("cuda", self.device)))
        self.gather_views = views   # one [world_batch, 12] view per scratch slot
        return views[0]

    def set_gather_fused(self, on):
        """True (default after setup_peer_gather): the solve kernels store the forces into every rank's gather buffer
        themselves; False: the gather is made by gather_push (copy engines)."""
        self._check(self._L.mpc_batch_set_gather_fused(self._h, int(bool(on))), "set_gather_fused")

    def gather_push(self, forces, slot=0, stream=None):
        """Copy-engine gather of this rank's forces [B, 12] (cuda float32) into region `slot` of every rank's gather
        buffer + the device-side cross-rank barrier, queued on `stream` (default: current)."""
        torch = _torch()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self._L.mpc_batch_gather_push_slot(self._h, int(slot), forces.data_ptr(), int(forces.shape[0]),
                                                       st.cuda_stream), "mpc_batch_gather_push_slot")

    def gather_sync(self, stream=None, slot=0):
        """Device-side cross-rank barrier of the fused gather for scratch slot `slot`, queued on `stream`
        (default: current)."""
        torch = _torch()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._check(self._L.mpc_batch_gather_sync_slot(self._h, int(slot), st.cuda_stream),
                    "mpc_batch_gather_sync_slot")

    # ---- solves ---------------------------------------------------------------------------
    def solve_device(self, records, forces=None, solution=None, status=None, want_solution=False,
                     want_status=True, stream=None, slot=0):
        """records: cuda uint8 tensor [B, stride] on this engine's device.  Asynchronous on `stream`
        (default: torch's current stream).  `slot` (0..SLOTS-1) picks the engine's device scratch: solves may
        overlap when they use different slots on different streams.  Returns (forces [B,12] f32, solution [B,12h] f64 | None,
        status [B] int32 | None), all cuda tensors."""
        torch = _torch()
        assert records.is_cuda and records.dtype == torch.uint8 and records.is_contiguous()
        B = records.shape[0]
        assert records.shape[1] == self.stride, (records.shape, self.stride)
        dev = records.device
        assert dev.index == self.device, "records live on cuda:%s, the engine on cuda:%d" % (dev.index, self.device)
        if forces is None:
            forces = torch.empty((B, 12), dtype=torch.float32, device=dev)
        if solution is None and want_solution:
            solution = torch.empty((B, 12 * self.horizon), dtype=torch.float64, device=dev)
        if status is None and want_status:
            status = torch.empty((B,), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self._L.mpc_batch_solve_device_slot(self._h, int(slot), records.data_ptr(), B, forces.data_ptr(),
                                                 solution.data_ptr() if solution is not None else None,
                                                 status.data_ptr() if status is not None else None, st.cuda_stream)
        self._check(rc, "mpc_batch_solve_device_slot")
        return forces, solution, status

    def solve_host(self, records, want_solution=False, out_forces=None, out_status=None, out_solution=None):
        """records: numpy uint8 [B, stride] in host memory.  Synchronous: H2D, kernels, D2H.
        Returns (forces [B,12] f32, solution [B,12h] f64 | None, status [B] int32)."""
        records = np.ascontiguousarray(records, np.uint8)
        B = records.shape[0]
        assert records.shape[1] == self.stride
        forces = out_forces if out_forces is not None else np.empty((B, 12), np.float32)
        want_solution = want_solution or out_solution is not None
        sol = out_solution if out_solution is not None else (
            np.empty((B, 12 * self.horizon), np.float64) if want_solution else None)
        status = out_status if out_status is not None else np.empty((B,), np.int32)
        rc = self._L.mpc_batch_solve_host(self._h, records.ctypes.data, B, forces.ctypes.data,
                                          sol.ctypes.data if want_solution else None, status.ctypes.data)
        self._check(rc, "mpc_batch_solve_host")
        return forces, sol, status

    def submit_host(self, slot, records, want_solution=False, zero_copy=False):
        """Queues H2D + kernels + D2H for `records` (numpy uint8 [B, stride]) on slot 0..SLOTS-1; returns at once.
        The records are copied before the call returns, so the array may be reused immediately.  zero_copy=True
        (opt-in): `records` is page-locked memory that the DMA engine reads in place -- it must stay untouched
        until wait_host(slot)."""
        records = np.ascontiguousarray(records, np.uint8)
        assert records.shape[1] == self.stride
        fn = self._L.mpc_batch_submit_host_pinned if zero_copy else self._L.mpc_batch_submit_host
        rc = fn(self._h, int(slot), records.ctypes.data, records.shape[0], int(bool(want_solution)))
        self._check(rc, "mpc_batch_submit_host")

    def submit_host_ticks(self, slot, ticks, want_solution=False, zero_copy=False):
        """The host entry from tick records (numpy [B, 68] of 32-bit words, ticks.pack_ticks): 272 bytes per robot
        cross the bus, the problem records are built on the device.  Collect with wait_host(slot); the controller
        state written back by the builder is host_state(slot)[:B] afterwards."""
        ticks = np.ascontiguousarray(ticks)
        assert ticks.itemsize * ticks.shape[1] == 272
        rc = self._L.mpc_batch_submit_host_ticks(self._h, int(slot), ticks.ctypes.data, ticks.shape[0],
                                                 int(bool(want_solution)), int(bool(zero_copy)))
        self._check(rc, "mpc_batch_submit_host_ticks")

    def host_state(self, slot=0):
        """(state [max_batch, 4] f32, ticks [max_batch, 68] f32): the slot's pinned controller-state output of the tick
        entry and its own pinned tick buffer."""
        ps, pt = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self._L.mpc_batch_host_state(self._h, int(slot), ctypes.byref(ps), ctypes.byref(pt)), "host_state")
        B = self.max_batch

        def view(p, nbytes, shape):
            buf = (ctypes.c_char * nbytes).from_address(p.value)
            return np.frombuffer(buf, dtype=np.float32).reshape(shape)

        return view(ps, B * 16, (B, 4)), view(pt, B * 272, (B, 68))

    def wait_host(self, slot, out_forces=None, out_solution=None, out_status=None):
        """Waits for slot `slot` and copies its results into the given numpy arrays (None: left in the slot's
        pinned buffers, see host_buffers)."""
        rc = self._L.mpc_batch_wait_host(self._h, int(slot),
                                         out_forces.ctypes.data if out_forces is not None else None,
                                         out_solution.ctypes.data if out_solution is not None else None,
                                         out_status.ctypes.data if out_status is not None else None)
        self._check(rc, "mpc_batch_wait_host")

    def build_records_device(self, ticks, records=None, want_state=True, stream=None):
        """ticks: cuda tensor [B, 68] of 32-bit words (ticks.pack_ticks).  Builds the problem records on the
        device.  Returns (records uint8 [B, stride], state_out float32 [B, 4] | None)."""
        torch = _torch()
        assert ticks.is_cuda and ticks.is_contiguous() and ticks.element_size() * ticks.shape[1] == 272
        B = ticks.shape[0]
        dev = ticks.device
        if records is None:
            records = torch.empty((B, self.stride), dtype=torch.uint8, device=dev)
        state = torch.empty((B, 4), dtype=torch.float32, device=dev) if want_state else None
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self._L.mpc_batch_build_records_device(self._h, ticks.data_ptr(), B, records.data_ptr(),
                                                    state.data_ptr() if state is not None else None, st.cuda_stream)
        self._check(rc, "mpc_batch_build_records_device")
        return records, state

    def solve_ticks_device(self, ticks, want_solution=False, want_state=True, stream=None):
        """Build + solve on the device: tick records in, (forces, solution | None, status, state_out | None) out."""
        torch = _torch()
        assert ticks.is_cuda and ticks.is_contiguous() and ticks.element_size() * ticks.shape[1] == 272
        B = ticks.shape[0]
        dev = ticks.device
        forces = torch.empty((B, 12), dtype=torch.float32, device=dev)
        sol = torch.empty((B, 12 * self.horizon), dtype=torch.float64, device=dev) if want_solution else None
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        state = torch.empty((B, 4), dtype=torch.float32, device=dev) if want_state else None
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self._L.mpc_batch_solve_ticks_device(self._h, ticks.data_ptr(), B, forces.data_ptr(),
                                                  sol.data_ptr() if sol is not None else None, status.data_ptr(),
                                                  state.data_ptr() if state is not None else None, st.cuda_stream)
        self._check(rc, "mpc_batch_solve_ticks_device")
        return forces, sol, status, state

    def gait_state_device(self, gait, want_table=False, stream=None):
        """gait records (cuda int32 [B, 12], see gait.pack_gait_records) -> (state [B, 10] float32 with
        word 0 an int32, table uint8 [B, 4*max nIterations] | None)."""
        torch = _torch()
        assert gait.is_cuda and gait.dtype == torch.int32 and gait.shape[1] == 12 and gait.is_contiguous()
        B = gait.shape[0]
        state = torch.empty((B, 10), dtype=torch.float32, device=gait.device)
        table = None
        stride = 0
        if want_table:
            stride = 4 * int(gait[:, 2].max().item())
            table = torch.zeros((B, stride), dtype=torch.uint8, device=gait.device)
        st = stream if stream is not None else torch.cuda.current_stream(gait.device)
        rc = self._L.mpc_batch_gait_state_device(self._h, gait.data_ptr(), B, state.data_ptr(),
                                                 table.data_ptr() if table is not None else None, stride, st.cuda_stream)
        self._check(rc, "mpc_batch_gait_state_device")
        return state, table

    def leg_commands_device(self, legs, forces, stream=None):
        """leg records (cuda float32 [B, 100], see legs.pack_leg_records) + solved forces [B, 12] ->
        (f_ff [B, 12] body-frame feed-forward forces, tau [B, 12] joint torques)."""
        torch = _torch()
        assert legs.is_cuda and legs.is_contiguous() and legs.element_size() * legs.shape[1] == 400
        assert forces.is_cuda and forces.dtype == torch.float32 and forces.is_contiguous()
        B = legs.shape[0]
        f_ff = torch.empty((B, 12), dtype=torch.float32, device=legs.device)
        tau = torch.empty((B, 12), dtype=torch.float32, device=legs.device)
        st = stream if stream is not None else torch.cuda.current_stream(legs.device)
        rc = self._L.mpc_batch_leg_commands_device(self._h, legs.data_ptr(), forces.data_ptr(), B, f_ff.data_ptr(),
                                                   tau.data_ptr(), st.cuda_stream)
        self._check(rc, "mpc_batch_leg_commands_device")
        return f_ff, tau

    def assemble_device(self, records, stream=None):
        """Parity entry: the reduced QP only.  Returns (nv [B] int32, H [B,12h,12h] f64, g [B,12h] f64)."""
        torch = _torch()
        B = records.shape[0]
        NU = 12 * self.horizon
        dev = records.device
        nv = torch.zeros((B,), dtype=torch.int32, device=dev)
        H = torch.zeros((B, NU, NU), dtype=torch.float64, device=dev)
        g = torch.zeros((B, NU), dtype=torch.float64, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self._L.mpc_batch_assemble_device(self._h, records.data_ptr(), B, nv.data_ptr(), H.data_ptr(),
                                               g.data_ptr(), st.cuda_stream)
        self._check(rc, "mpc_batch_assemble_device")
        return nv, H, g


def status_code(status):
    return np.asarray(status) & 0xff


def status_iterations(status):
    return (np.asarray(status).astype(np.uint32)) >> 8

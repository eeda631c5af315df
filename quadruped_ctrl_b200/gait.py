"""Contact-table generator: host mirror of the reference's OffsetDurationGait.

Semantics follow /root/reference/src/MPC_Ctrl/Gait.cpp:142-166 (getMpcTable) and
:187-193 (setIterations); the gait catalogue follows
/root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:23-41 (14-segment gaits) and
:149-172 (gait-number mapping).  Used to synthesise bench / test inputs; the MPC
engine itself only consumes the resulting 0/1 table.
"""
import numpy as np

# name -> (offsets, durations) for nSegment = 14 (ConvexMPCLocomotion.cpp:25-41; int truncation as Vec4<int>)
GAITS_14 = {
    "trotting": ((0, 7, 7, 0), (7, 7, 7, 7)),
    "bounding": ((7, 7, 0, 0), (6, 6, 6, 6)),
    "pronking": ((0, 0, 0, 0), (6, 6, 6, 6)),
    "jumping": ((0, 0, 0, 0), (3, 3, 3, 3)),
    "galloping": ((0, 4, 7, 11), (7, 7, 7, 7)),
    "standing": ((0, 0, 0, 0), (14, 14, 14, 14)),
    "trotRunning": ((0, 7, 7, 0), (6, 6, 6, 6)),
    "walking": ((0, 7, 3, 10), (10, 10, 10, 10)),
    "walking2": ((0, 7, 7, 0), (10, 10, 10, 10)),
    "pacing": ((7, 0, 7, 0), (7, 7, 7, 7)),
}
# gaitNumber -> name (ConvexMPCLocomotion.cpp:149-172); 0,3,6 and anything else fall through to trotting
GAIT_NUMBER = {1: "bounding", 2: "pronking", 4: "standing", 5: "trotRunning", 7: "galloping", 8: "pacing",
               9: "trotting", 10: "walking", 11: "walking2"}


def gait_by_number(n):
    return GAIT_NUMBER.get(int(n), "trotting")


def rescale(offsets, durations, n_segments, base=14):
    """Rescales a 14-segment gait to n_segments by round(x*n/14)."""
    f = n_segments / float(base)
    off = tuple(int(np.floor(o * f + 0.5)) % n_segments for o in offsets)
    dur = tuple(min(n_segments, max(0, int(np.floor(d * f + 0.5)))) for d in durations)
    return off, dur


def mpc_table(n_segments, offsets, durations, iteration):
    """[n_segments, 4] 0/1 contact table at gait iteration `iteration` (Gait.cpp:142-166)."""
    off = np.asarray(offsets, np.int64)
    dur = np.asarray(durations, np.int64)
    i = np.arange(n_segments)[:, None]
    it = (i + int(iteration) + 1) % n_segments
    progress = it - off[None, :]
    progress = np.where(progress < 0, progress + n_segments, progress)
    return (progress < dur[None, :]).astype(np.int32)


def mpc_tables(n_segments, offsets, durations, iterations):
    """Vectorised mpc_table: offsets/durations [B,4] (or [4]), iterations [B] -> [B, n_segments*4]."""
    iterations = np.asarray(iterations, np.int64)
    B = iterations.shape[0]
    off = np.broadcast_to(np.asarray(offsets, np.int64), (B, 4))
    dur = np.broadcast_to(np.asarray(durations, np.int64), (B, 4))
    i = np.arange(n_segments)[None, :, None]
    it = (i + iterations[:, None, None] + 1) % n_segments
    progress = it - off[:, None, :]
    progress = np.where(progress < 0, progress + n_segments, progress)
    return (progress < dur[:, None, :]).astype(np.int32).reshape(B, n_segments * 4)


def set_iterations(n_segments, iterations_per_mpc, current_iteration):
    """(iteration, phase) bookkeeping of Gait.cpp:187-193."""
    iteration = (current_iteration // iterations_per_mpc) % n_segments
    phase = float(current_iteration % (iterations_per_mpc * n_segments)) / float(iterations_per_mpc * n_segments)
    return iteration, phase

"""EKF smoother over an ensemble -- host mirror of MavStateEstimator::EKFSmoothBackwardsPass
(MSE/mav_state_est.cpp:98-189) on top of rbis_smooth_plan / rbis_batch_smooth_backward (include/rbis_batch.h).

The reference keeps every update's posterior in its history node; an ensemble keeps them in the snapshot ring:
`forward_program` appends a SNAPSHOT after every update (slot u for history entry u, slot 0 for the reset at the front of
the history), `smooth` plans the backwards traversal on the host and runs it on the device, and `posterior_slot[u]` then
names the slot that holds update u's smoothed posterior.
"""
import ctypes as C

import numpy as np

from . import capi

STEP_DTYPE = np.dtype([("cur_slot", "<i4"), ("cur_pred_slot", "<i4"), ("out_slot", "<i4"), ("reserved", "<i4")])


def forward_program(events, utime0=0):
    """events: in-order (kind, stream, row, utime, dt) -> (ops with a SNAPSHOT after every update, is_ins [n], slot [n]);
    history entry 0 is the reset (the state the ensemble starts from), entries 1.. are the events."""
    ops = [(capi.OP_SNAPSHOT, 0, 0, utime0, 0.0)]
    is_ins, slot = [0], [0]
    for u, e in enumerate(events, 1):
        ops.append(tuple(e))
        ops.append((capi.OP_SNAPSHOT, 0, u, e[3], 0.0))
        is_ins.append(1 if e[0] == capi.OP_IMU else 0)
        slot.append(u)
    return ops, np.asarray(is_ins, dtype=np.uint8), np.asarray(slot, dtype=np.int32)


def plan(is_ins, slot):
    """rbis_smooth_plan -> (next_pred_slot, next_slot, steps [n_steps] of STEP_DTYPE, alias [n])."""
    lib = capi.load()
    is_ins = np.ascontiguousarray(is_ins, dtype=np.uint8)
    slot = np.ascontiguousarray(slot, dtype=np.int32)
    n = len(is_ins)
    assert len(slot) == n
    steps = np.zeros(max(n, 1), dtype=STEP_DTYPE)
    alias = np.zeros(n, dtype=np.int32)
    a, b = C.c_int32(0), C.c_int32(0)
    k = lib.rbis_smooth_plan(n, is_ins.ctypes.data, slot.ctypes.data, C.addressof(a), C.addressof(b), steps.ctypes.data,
                             alias.ctypes.data)
    if k < 0:
        capi.check(int(k))
    return a.value, b.value, steps[:k].copy(), alias


def smooth(batch, is_ins, slot, dt):
    """Plan + run the backward pass on `batch` (an RBISBatch whose ring holds the forward posteriors).  Returns alias."""
    np_slot, n_slot, steps, alias = plan(is_ins, slot)
    batch.smooth_backward(np_slot, n_slot, steps, dt)
    return alias

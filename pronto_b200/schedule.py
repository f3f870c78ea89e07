"""Host mirror of the delayed-measurement planner (rbis_planner_* in include/rbis_batch.h).

All logic lives in librbis_b200.so (csrc/rbis_planner.cpp); this is the ctypes wrapper."""
import ctypes as C

import numpy as np

from . import capi
from .batch import OP_DTYPE


class Planner:
    """Batch stand-in for MavStateEstimator::addUpdate's history (MSE/mav_state_est.cpp:28-80)."""

    def __init__(self, utime0=0, snapshot_slots=4, snapshot_period_us=100_000, snapshot_phase_us=0,
                 history_span_us=0):
        self.lib = capi.load()
        self.h = C.c_void_p()
        capi.check(self.lib.rbis_planner_create(C.byref(self.h), int(utime0), int(snapshot_slots),
                                                int(snapshot_period_us), int(snapshot_phase_us), int(history_span_us)))
        self.snapshot_slots = int(snapshot_slots)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.rbis_planner_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def add_update(self, kind, stream, row, utime, dt=0.0, roll_forward=True):
        """-> True if accepted, False if discarded as too old."""
        op = capi.Op(int(kind), int(stream), int(row), int(utime), float(dt))
        rc = self.lib.rbis_planner_add_update(self.h, C.byref(op), int(bool(roll_forward)))
        if rc < 0:
            capi.check(rc)
        return rc == 0

    def take(self):
        """Pending op program as a structured array (OP_DTYPE), cleared from the planner."""
        n = int(self.lib.rbis_planner_pending(self.h))
        out = np.zeros(n, dtype=OP_DTYPE)
        got = C.c_int64(0)
        capi.check(self.lib.rbis_planner_take(self.h, out.ctypes.data, n, C.byref(got)))
        assert got.value == n
        return out

    def counters(self):
        c = (C.c_int64 * 6)()
        capi.check(self.lib.rbis_planner_counters(self.h, c))
        return dict(zip(("accepted", "discarded", "rewinds", "replayed", "snapshots", "retained"), [int(x) for x in c]))


def program_from_arrivals(arrivals, snapshot_slots=4, snapshot_period_us=100_000, snapshot_phase_us=0, utime0=0,
                          history_span_us=0, roll_forward_every=1):
    """arrivals: (kind, stream, row, utime, dt) in ARRIVAL order -> (op program, planner counters)."""
    p = Planner(utime0, snapshot_slots, snapshot_period_us, snapshot_phase_us, history_span_us)
    arrivals = list(arrivals)
    for i, (kind, stream, row, utime, dt) in enumerate(arrivals):
        rf = ((i + 1) % roll_forward_every == 0) or i == len(arrivals) - 1
        p.add_update(kind, stream, row, utime, dt, roll_forward=rf)
    ops, cnt = p.take(), p.counters()
    p.close()
    return ops, cnt

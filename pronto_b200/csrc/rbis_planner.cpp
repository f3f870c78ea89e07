// rbis_planner.cpp -- host-side planner for delayed measurements (see include/rbis_batch.h).
//
// Restates the ordering semantics of the reference's history driver for a whole ensemble:
//   MSE/update_history.cpp:16-54   time-ordered multimap, hinted insert (equal keys keep arrival
//                                  order), discard of updates older than the oldest retained entry
//   MSE/mav_state_est.cpp:28-80    out-of-order insert moves the unprocessed start back; the roll
//                                  forward replays every update from there to the head; truncation
//                                  to utime_history_span behind the newest update
// The reference keeps one posterior per history node; here the device keeps a small ring of ensemble
// snapshots, so a rewind restores the nearest snapshot at or before the insertion point and replays
// from there.
#include <algorithm>
#include <climits>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/rbis_batch.h"

int rbis_set_error(int code, const char* fmt, ...);  // rbis_batch.cu

struct rbis_planner {
  struct Snap {
    int64_t pos;    // number of history entries (base included) applied when the snapshot was taken
    int64_t utime;  // utime of entry pos-1
    int32_t slot;
  };
  std::vector<rbis_op_t> hist;  // hist[0] is the base (reset, or the oldest retained update): never re-applied
  std::vector<Snap> snaps;      // ascending pos
  std::vector<int32_t> free_slots;
  std::vector<rbis_op_t> pending;
  int64_t valid = 1;            // hist[0..valid-1]: applied on the device and still in that order
  int64_t device_count = 1;     // number of entries the device head has applied (>= valid)
  bool need_restore = false;    // the device head reflects an order that is no longer the history's
  int32_t n_slots = 0;
  int64_t period = 0, phase = 0, span = 0;
  int64_t counters[6] = {0, 0, 0, 0, 0, 0};

  bool snapshot_time(int64_t u) const {
    if (period <= 0) return false;
    int64_t r = (u - phase) % period;
    return r == 0;
  }
  const Snap* snap_at_or_before(int64_t pos) const {
    const Snap* best = nullptr;
    for (const Snap& s : snaps)
      if (s.pos <= pos && (!best || s.pos > best->pos)) best = &s;
    return best;
  }
  bool has_snap_at(int64_t pos) const {
    for (const Snap& s : snaps)
      if (s.pos == pos) return true;
    return false;
  }
  void emit_snapshot(int64_t pos) {
    if (n_slots <= 0) return;
    const int64_t u = hist[(size_t)pos - 1].utime;
    int32_t slot = -1;
    // a snapshot of the same utime is superseded (the later one sits after more updates stamped u)
    for (size_t i = 0; i < snaps.size(); i++)
      if (snaps[i].utime == u && snaps[i].pos < pos) {
        slot = snaps[i].slot;
        snaps.erase(snaps.begin() + (long)i);
        break;
      }
    if (slot < 0 && !free_slots.empty()) {
      slot = free_slots.back();
      free_slots.pop_back();
    }
    if (slot < 0) {  // ring is full: the oldest snapshot goes
      size_t o = 0;
      for (size_t i = 1; i < snaps.size(); i++)
        if (snaps[i].pos < snaps[o].pos) o = i;
      slot = snaps[o].slot;
      snaps.erase(snaps.begin() + (long)o);
    }
    snaps.push_back({pos, u, slot});
    rbis_op_t op;
    std::memset(&op, 0, sizeof(op));
    op.kind = RBIS_OP_SNAPSHOT;
    op.row = slot;
    op.utime = u;
    pending.push_back(op);
    counters[4]++;
  }
  void roll_forward() {
    int64_t start = valid;
    if (need_restore) {
      const Snap* s = snap_at_or_before(valid);  // existence guaranteed by add_update
      rbis_op_t op;
      std::memset(&op, 0, sizeof(op));
      op.kind = RBIS_OP_RESTORE;
      op.row = s->slot;
      op.utime = s->utime;
      pending.push_back(op);
      counters[2]++;
      counters[3] += device_count - s->pos;  // updates the device had applied and now applies again
      start = s->pos;
      need_restore = false;
    }
    if (start == 1 && n_slots > 0 && !has_snap_at(1) && (int64_t)hist.size() > 1) emit_snapshot(1);
    for (int64_t i = start; i < (int64_t)hist.size(); i++) {
      const int64_t u = hist[(size_t)i - 1].utime;
      if (i > 1 && hist[(size_t)i].utime > u && snapshot_time(u) && !has_snap_at(i)) emit_snapshot(i);
      pending.push_back(hist[(size_t)i]);
    }
    valid = device_count = (int64_t)hist.size();
    // truncation, mav_state_est.cpp:74-77 + update_history.cpp:44-54: the last entry at or before
    // (newest - span) becomes the new base
    if (span > 0) {
      const int64_t cutoff = hist.back().utime - span;
      int64_t keep = -1;
      for (int64_t i = 0; i < (int64_t)hist.size() && hist[(size_t)i].utime <= cutoff; i++) keep = i;
      if (keep > 0) {
        hist.erase(hist.begin(), hist.begin() + keep);
        valid -= keep;
        device_count -= keep;
        std::vector<Snap> kept;
        for (const Snap& s : snaps) {
          if (s.pos - keep >= 1) kept.push_back({s.pos - keep, s.utime, s.slot});
          else free_slots.push_back(s.slot);
        }
        snaps.swap(kept);
      }
    }
    counters[5] = (int64_t)hist.size();
  }
};

extern "C" {

int rbis_planner_create(rbis_planner_t** out, int64_t utime0, int32_t snapshot_slots, int64_t snapshot_period_us,
                        int64_t snapshot_phase_us, int64_t history_span_us) {
  if (!out) return rbis_set_error(RBIS_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (snapshot_slots < 0 || snapshot_period_us < 0 || history_span_us < 0)
    return rbis_set_error(RBIS_ERR_INVALID, "snapshot_slots, snapshot_period_us and history_span_us must be >= 0");
  rbis_planner* p = new (std::nothrow) rbis_planner();
  if (!p) return rbis_set_error(RBIS_ERR_ALLOC, "host allocation failed");
  rbis_op_t base;
  std::memset(&base, 0, sizeof(base));
  base.kind = -1;  // the reset update the estimator was created with (MSE/mav_state_est.cpp:12-16)
  base.utime = utime0;
  p->hist.push_back(base);
  p->n_slots = snapshot_slots;
  p->period = snapshot_period_us;
  p->phase = snapshot_phase_us;
  p->span = history_span_us;
  for (int32_t s = snapshot_slots - 1; s >= 0; s--) p->free_slots.push_back(s);
  p->counters[5] = 1;
  *out = p;
  // An update can only be rewound to while a snapshot at or before it is retained, i.e. within ~ slots x period of the head;
  // the reference replays anything newer than the oldest history entry (the whole span).  Not an error -- a shorter window buys
  // shorter replays -- but the caller should know: the text is left in rbis_last_error().
  if (history_span_us > 0 && (int64_t)snapshot_slots * snapshot_period_us < history_span_us)
    rbis_set_error(0, "rbis_planner_create: rewind window %lld us (snapshot_slots x snapshot_period_us) is shorter than history_span_us %lld: "
                      "updates delayed beyond it are discarded where the reference would replay them",
                   (long long)((int64_t)snapshot_slots * snapshot_period_us), (long long)history_span_us);
  return 0;
}

int rbis_planner_destroy(rbis_planner_t* p) {
  delete p;
  return 0;
}

int rbis_planner_add_update(rbis_planner_t* p, const rbis_op_t* update, int roll_forward) {
  if (!p || !update) return rbis_set_error(RBIS_ERR_INVALID, "null planner or update");
  if (update->kind != RBIS_OP_IMU && update->kind != RBIS_OP_MEAS)
    return rbis_set_error(RBIS_ERR_INVALID, "only IMU and measurement updates enter the history (kind %d)", update->kind);
  // hinted insert at end(): the position after the last entry with utime <= update->utime
  int64_t pos = (int64_t)p->hist.size();
  while (pos > 0 && p->hist[(size_t)pos - 1].utime > update->utime) pos--;
  int rc = 0;
  if (pos == 0) {
    rc = 1;  // before the first in history: discarded (update_history.cpp:28-39)
  } else if (pos < p->valid && !p->snap_at_or_before(pos)) {
    rc = 1;  // no retained snapshot precedes it: nothing to rewind to
  }
  if (rc == 1) {
    p->counters[1]++;
  } else {
    p->hist.insert(p->hist.begin() + pos, *update);
    if (pos < p->valid) {
      // out of order with respect to what the device has applied: snapshots taken after the
      // insertion point no longer describe a prefix of the history
      std::vector<rbis_planner::Snap> kept;
      for (const auto& s : p->snaps) {
        if (s.pos <= pos) kept.push_back(s);
        else p->free_slots.push_back(s.slot);
      }
      p->snaps.swap(kept);
      p->valid = pos;
      p->need_restore = true;
    }
    p->counters[0]++;
  }
  if (roll_forward && ((int64_t)p->hist.size() > p->valid || p->need_restore)) p->roll_forward();
  return rc;
}

int64_t rbis_planner_pending(const rbis_planner_t* p) { return p ? (int64_t)p->pending.size() : 0; }

int rbis_planner_take(rbis_planner_t* p, rbis_op_t* out, int64_t cap, int64_t* n_out) {
  if (!p || !n_out) return rbis_set_error(RBIS_ERR_INVALID, "null planner or n_out");
  const int64_t n = (int64_t)p->pending.size();
  if (n > 0 && (!out || cap < n)) return rbis_set_error(RBIS_ERR_INVALID, "program has %lld ops, capacity %lld", (long long)n, (long long)cap);
  if (n > 0) std::memcpy(out, p->pending.data(), (size_t)n * sizeof(rbis_op_t));
  *n_out = n;
  p->pending.clear();
  return 0;
}

int rbis_planner_counters(const rbis_planner_t* p, int64_t out[6]) {
  if (!p || !out) return rbis_set_error(RBIS_ERR_INVALID, "null planner or out");
  std::memcpy(out, p->counters, sizeof(p->counters));
  return 0;
}

// EKFSmoothBackwardsPass's traversal on indices (mav_state_est.cpp:98-189): `it` plays current_it, -1 plays the
// `--begin()` stop mark.
int64_t rbis_smooth_plan(int64_t n, const uint8_t* is_ins, const int32_t* slot, int32_t* next_pred_slot, int32_t* next_slot,
                         rbis_smooth_step_t* steps, int32_t* alias) {
  if (n <= 0 || !is_ins || !slot || !next_pred_slot || !next_slot || !steps || !alias)
    return rbis_set_error(RBIS_ERR_INVALID, "null argument or empty history");
  bool any_ins = false;
  for (int64_t u = 0; u < n; u++) {
    alias[u] = slot[u];
    any_ins = any_ins || is_ins[u];
  }
  if (!any_ins) return rbis_set_error(RBIS_ERR_STATE, "history holds no IMU process step");
  int64_t it = n - 1;  // :108-109  latest processed update
  bool measurement_cur_step = false;
  int64_t next = -1;
  int64_t cu = it;
  while (!is_ins[cu]) {  // :116-124  rewind through the trailing measurements
    cu = it;
    if (!measurement_cur_step) {
      next = cu;
      measurement_cur_step = true;
    }
    it--;
  }
  const int64_t next_pred = cu;  // :126-127
  if (!measurement_cur_step) next = cu;
  it--;  // :130
  *next_pred_slot = slot[next_pred];
  *next_slot = slot[next];
  measurement_cur_step = false;
  int64_t cur = -1, n_steps = 0;
  while (it >= 0) {  // :141-187 (it < -1 can only follow the extra decrement on a two-entry history)
    cu = it;
    if (is_ins[cu]) {
      if (!measurement_cur_step) cur = cu;
      rbis_smooth_step_t st;
      st.cur_slot = slot[cur];
      st.cur_pred_slot = slot[cu];
      st.out_slot = slot[cu];
      st.reserved = 0;
      steps[n_steps++] = st;
      // :163-169  the measurement updates that follow this IMU step take the smoothed posterior
      for (int64_t f = cu + 1; f < n && !is_ins[f]; f++) alias[f] = slot[cu];
      measurement_cur_step = false;
    } else if (!measurement_cur_step) {
      cur = cu;
      measurement_cur_step = true;
    }
    it--;
  }
  return n_steps;
}

}  // extern "C"

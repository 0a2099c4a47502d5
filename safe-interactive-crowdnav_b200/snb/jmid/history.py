"""History join + resampling of the predictor wrapper (sicnav_diffusion/JMID/mid_sim_wrapper.py:244-298) for the B = 1 plugin
object, without pandas: the per-agent [x, y, t] lists are joined on their exact time stamps, binned from the END into windows of
int(round(time_step * 100)) "nanoseconds" of t * 100 (the reference's pd.to_datetime(time * 100) trick, :289-292), the last row of
every window is kept, empty windows are filled by linear interpolation between their filled neighbours, and the newest
past_num_frames rows are returned.  In the simulator every frame is time_step apart and this is the identity; it matters when the
caller records poses faster than time_step or misses frames.

Host code by design: a few dozen rows per call on the one-environment path; the batched device path (ForecasterBatch) pushes one
frame per env step into its rings and never needs it.
"""
import numpy as np


def resample_histories(prev_states, prev_robot_states, time_step, num_hist_frames):
    """prev_states: H lists of [x, y, t]; prev_robot_states: list of [x, y, t].  -> (humans [H, F, 2], robot [F, 2]) float64,
    F <= num_hist_frames, oldest first.  Rows whose time stamp is missing from any agent's list are dropped (the left joins +
    dropna of :262-267); time stamps are assumed unique within a list."""
    H = len(prev_states)
    cols = [np.asarray(p, np.float64).reshape(-1, 3) for p in prev_states] + [np.asarray(prev_robot_states, np.float64).reshape(-1, 3)]
    times = cols[0][:, 2]
    keep = np.ones(len(times), bool)
    idx = []
    for c in cols[1:]:
        pos = {t: k for k, t in enumerate(c[:, 2].tolist())}          # exact float equality, like the index join
        ix = np.array([pos.get(t, -1) for t in times.tolist()], np.int64)
        keep &= ix >= 0
        idx.append(ix)
    if not keep.any():
        return np.zeros((H, 0, 2)), np.zeros((0, 2))
    times = times[keep]
    rows = np.empty((len(times), H + 1, 2), np.float64)
    rows[:, 0] = cols[0][keep, :2]
    for a, (c, ix) in enumerate(zip(cols[1:], idx), start=1):
        rows[:, a] = c[ix[keep], :2]
    order = np.argsort(times, kind="stable")                           # sort_values(by="time")
    times, rows = times[order], rows[order]
    t_ns = np.trunc(times * 100).astype(np.int64)                      # pd.to_datetime(float): whole nanoseconds, toward zero
    w = int(round(time_step * 100))
    end = int(t_ns.max())
    k = (end - t_ns) // w                                              # window (end - (k+1) w, end - k w], closed right, origin = "end"
    n = int(k.max()) + 1
    out = np.full((n, H + 1, 2), np.nan)
    for r in range(len(times)):                                        # rows are time-sorted: the last row of a window wins
        out[n - 1 - int(k[r])] = rows[r]
    filled = np.flatnonzero(~np.isnan(out[:, 0, 0]))
    for a, b in zip(filled[:-1], filled[1:]):                          # interpolate(method="linear"): equally spaced rows
        for m in range(a + 1, b):
            out[m] = out[a] + (out[b] - out[a]) * ((m - a) / (b - a))
    out = out[filled[0]:][-num_hist_frames:]
    return np.ascontiguousarray(out[:, :H].transpose(1, 0, 2)), np.ascontiguousarray(out[:, H])

"""Oracle: event-window selection and accumulation.

Test infrastructure (see oracle/__init__.py).  Restates
  model/nerf.py:162-205     closed time window [low, up] over sorted-by-time events
  utils/event_utils.py:247-259  accumulate_events_on_gpu: COO (ys, xs, pol) summed
                                in fp32 then added to a float64 zero image (Q10)
"""
import numpy as np
import torch


def select_window(events, low_t, up_t):
    """events: dict of numpy arrays x, y, ts, pol.  Closed on both ends (Q16)."""
    keep = np.where((low_t <= events["ts"]) * (events["ts"] <= up_t))
    return {k: events[k][keep] for k in ("x", "y", "ts", "pol")}


def accumulate(height, width, xs, ys, pol):
    """-> float64 [height, width] polarity sums."""
    idx = torch.tensor(np.array([ys, xs]), dtype=torch.long)
    val = torch.tensor(np.asarray(pol), dtype=torch.float32)
    dense = torch.sparse_coo_tensor(idx, val, torch.Size([height, width])).to_dense()
    return torch.zeros(height, width, dtype=torch.float64) + dense

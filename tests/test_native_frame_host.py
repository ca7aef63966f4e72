"""Host-side logic of the sync-free frame (gms_b200.trainer.NativeFrame) without a GPU: the per-view prediction of the
binning capacity and the harvesting of the mapped (N, flag) ring -- the part of SURVEY.md section 7 step 4 ("sync-free
frame") that is plain Python.  The device side (sentinel keys, overflow => background frame) is covered by the -m gpu tests."""
import collections

import numpy as np

from gms_b200.trainer import NativeFrame


def _bare_frame():
    fr = NativeFrame.__new__(NativeFrame)          # no CUDA objects: only the fields the host logic touches
    fr.capacity, fr.capacity_override = 1, None
    fr._n_np = np.zeros((NativeFrame.RING, 2), dtype=np.int32)
    fr._pending = collections.deque()
    fr._view_n, fr._n_max, fr._n_last, fr._frame_no, fr.overflows = {}, 0, 0, 0, 0
    return fr


def test_capacity_prediction_per_view():
    fr = _bare_frame()
    fr._note("a", 4_000_000)
    # a camera never seen: sized from the largest N so far, generously
    assert fr._predict_capacity("b") == int(4_000_000 * 1.25) + (1 << 18)
    # second visit of a camera whose N is known: 8 % + 64k
    assert fr._predict_capacity("a") == int(4_000_000 * 1.08) + (1 << 16)
    # a camera that moved 10 % between its last two visits gets three times that as margin
    fr._note("a", 4_400_000)
    drift = 400_000 / 4_400_000
    assert fr._predict_capacity("a") == int(4_400_000 * (1.0 + 3.0 * drift)) + (1 << 16)
    assert fr._n_max == 4_400_000 and fr._n_last == 4_400_000
    fr.capacity_override = 12345
    assert fr._predict_capacity("a") == 12345 and fr._predict_capacity("zzz") == 12345


def test_harvest_is_in_order_counts_overflows_and_stops_at_unfinished_frames():
    fr = _bare_frame()
    fr._note("a", 1000)
    # three frames in flight: slots 0, 1, 2; the device has finished the first two (slot value != -1)
    for slot, (key, cap) in enumerate((("a", 1200), ("b", 1500), ("a", 1300))):
        fr._n_np[slot] = (-1, 0)
        fr._pending.append((slot, key, cap))
    fr._n_np[0] = (1100, 0)
    fr._n_np[1] = (1800, 1)          # N above that frame's capacity: an overflow (the host derives it from N > capacity)
    fr._harvest()
    assert [p[0] for p in fr._pending] == [2]                      # slot 2 still -1: stays pending, nothing behind it is read
    assert fr._view_n["a"] == (1100, 1000) and fr._view_n["b"] == (1800, 0)
    assert fr.overflows == 1 and fr._n_max == 1800 and fr._n_last == 1800
    # the overflowed camera's next visit is sized from its TRUE N
    assert fr._predict_capacity("b") >= 1800
    fr._n_np[2] = (1250, 0)
    fr._harvest()
    assert not fr._pending and fr._view_n["a"] == (1250, 1100) and fr.overflows == 1

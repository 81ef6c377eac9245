"""Host-side bookkeeping of the temporal path (romp_b200/temporal.py; simple_romp/romp/main.py:117-157): which One-Euro
filter slot a detection uses.  The filters themselves are pinned to the reference in tests/test_oracle_golden.py::test_one_euro
and run on the GPU (tests/test_gpu_temporal.py); the association replaces the third-party norfair tracker (parity unpinned)."""
import numpy as np

from romp_b200.temporal import MAX_TRACKS_PER_SIGNAL, NearestCenterTracker, TemporalState


def test_tracker_keeps_ids_within_the_reference_threshold_and_ages_out():
    tr = NearestCenterTracker(distance_threshold=200.0, max_age=2)
    ids0, fresh0 = tr.update([[100.0, 100.0], [400.0, 100.0]])
    assert ids0 == [1, 2] and fresh0 == [1, 2]
    ids1, fresh1 = tr.update([[410.0, 120.0], [90.0, 95.0]])           # order of the detections swapped, small motion
    assert ids1 == [2, 1] and fresh1 == []
    ids2, fresh2 = tr.update([[100.0, 400.0]])                          # 300 px away from both: a new person
    assert ids2 == [3] and fresh2 == [3]
    for _ in range(3):                                                  # tracks 1 and 2 unseen for > max_age frames are dropped
        tr.update([[100.0, 400.0]])
    assert set(tr.tracks) == {3}


def test_slots_follow_tracks_and_new_tracks_reset_their_filter_state():
    st = TemporalState(show_largest=False)
    cams = np.array([[0.8, 0.1, 0.2], [0.5, 0.6, 0.7]], np.float32)     # tracker points = cam[[2, 1]] * 512 (main.py:138)
    slots, ids, reset = st.assign(cams, signal_ID=0)
    assert list(slots) == [0, 1] and list(ids) == [1, 2]
    assert set(reset) == set(range(MAX_TRACKS_PER_SIGNAL))              # first frame of a signal: its whole slot block is reset
    slots2, ids2, reset2 = st.assign(cams[::-1].copy(), signal_ID=0)     # same two people, detections in the other order
    assert list(slots2) == [1, 0] and list(ids2) == [2, 1] and reset2 == []
    far = np.array([[0.5, -0.9, -0.9]], np.float32)                      # a third person appears: next free slot, reset only that one
    slots3, ids3, reset3 = st.assign(far, signal_ID=0)
    assert list(slots3) == [2] and list(ids3) == [3] and reset3 == [2]
    slots4, _, reset4 = st.assign(cams, signal_ID=7)                     # another signal gets its own block of slots
    assert list(slots4) == [MAX_TRACKS_PER_SIGNAL, MAX_TRACKS_PER_SIGNAL + 1]
    assert set(reset4) == set(range(MAX_TRACKS_PER_SIGNAL, 2 * MAX_TRACKS_PER_SIGNAL))


def test_show_largest_smooths_only_the_largest_person():
    st = TemporalState(show_largest=True)
    cams = np.array([[0.3, 0.0, 0.0], [0.9, 0.1, 0.1], [0.5, 0.2, 0.2]], np.float32)
    slots, ids, _ = st.assign(cams, signal_ID=0)                        # main.py:128-134: argmax of the cam scale
    assert list(slots) == [-1, 0, -1] and ids is None


def test_two_detections_on_one_track_do_not_share_a_filter():
    st = TemporalState(show_largest=False)
    st.assign(np.array([[0.5, 0.0, 0.0]], np.float32), 0)
    slots, ids, _ = st.assign(np.array([[0.5, 0.0, 0.0], [0.5, 0.01, 0.01]], np.float32), 0)   # both within 200 px of track 1
    assert ids[0] == ids[1] == 1 and slots[0] == 0 and slots[1] == -1

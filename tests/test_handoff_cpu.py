"""Host logic of the legacy hand-off format; no GPU, no compute."""
import numpy as np
import pytest

from deepi2p_b200 import handoff, synthetic


def _write(tmp_path, n_frames=3, n_points=257):
    recs = []
    for i in range(n_frames):
        s = synthetic.make_sample(100 + i, n_points=n_points + i)      # ragged on purpose
        name = "%06d_%02d" % (i * 30, i % 2)
        P34 = s["P_gt"][:3] if i % 2 == 0 else s["P_gt"]               # both shapes occur in the wild
        handoff.save_record(str(tmp_path), name, s["points"], s["pred"], s["gt"], s["pred"], s["gt"], s["K"], P34)
        recs.append((name, s))
    return recs


def test_roundtrip_matches_reference_reader(tmp_path):
    recs = _write(tmp_path)
    names = handoff.list_records(str(tmp_path))
    assert names == sorted(n for n, _ in recs)
    b = handoff.load_batch(str(tmp_path), names)
    assert b["xyz"].dtype == np.float32 and b["label"].dtype == np.int8 and b["K"].shape == (3, 9)
    for s_idx, name in enumerate(names):
        smp = dict(recs)[name]
        n = smp["points"].shape[1]
        assert b["n_pts"][s_idx] == n
        # what registration_lsq.py:291-298 would have read
        raw = np.load(tmp_path / (name + "_pc_label.npy"))
        assert raw.shape[0] == 7
        np.testing.assert_array_equal(raw[0:3].astype(np.float64), smp["points"].astype(np.float64))
        np.testing.assert_array_equal(raw[3].astype(np.int64), smp["pred"])
        np.testing.assert_array_equal(b["xyz"][s_idx, :, :n], smp["points"].astype(np.float32))
        np.testing.assert_array_equal(b["label"][s_idx, :n], smp["pred"].astype(np.int8))
        assert (b["label"][s_idx, n:] == -1).all() and (b["xyz"][s_idx, :, n:] == 0).all()
        np.testing.assert_array_equal(b["K"][s_idx].reshape(3, 3), smp["K"])
        np.testing.assert_array_equal(b["P_gt"][s_idx], smp["P_gt"])    # 3x4 files get their last row back


def test_label_row_selection_and_unknown_labels(tmp_path):
    s = synthetic.make_sample(5, n_points=64)
    odd = s["pred"].copy()
    odd[:4] = 2                                                          # neither inside nor outside: ignored
    handoff.save_record(str(tmp_path), "000000_00", s["points"], odd, s["gt"], s["pred"], s["gt"], s["K"], s["P_gt"])
    b = handoff.load_batch(str(tmp_path), which="coarse_prediction")
    assert (b["label"][0, :4] == -1).all()
    g = handoff.load_batch(str(tmp_path), which="coarse_label")
    np.testing.assert_array_equal(g["label"][0], s["gt"].astype(np.int8))
    with pytest.raises(KeyError):
        handoff.load_batch(str(tmp_path), which="nope")


def test_enu2cam_is_the_reference_transform(tmp_path):
    s = synthetic.make_sample(6, n_points=32)
    handoff.save_record(str(tmp_path), "000000_00", s["points"], s["pred"], s["gt"], s["pred"], s["gt"], s["K"],
                        s["P_gt"])
    pc, _, _, P = handoff.load_record(str(tmp_path), "000000_00", enu2cam=True)
    p = s["points"]
    np.testing.assert_array_equal(pc, np.stack([p[0], -p[2], p[1]]))     # registration_lsq.py:242-245
    # the composite P * [pc;1] is unchanged by the change of axes (:247)
    h = np.vstack([p.astype(np.float64), np.ones((1, p.shape[1]))])
    h2 = np.vstack([pc.astype(np.float64), np.ones((1, p.shape[1]))])
    np.testing.assert_allclose(P @ h2, s["P_gt"] @ h, atol=1e-9)


def test_summarize_follows_result_analysis():
    t = np.array([0.5, 3.0, 1.0, 9.0])
    r = np.array([1.0, 2.0, 7.0, 9.0])
    cost = np.array([10.0, 5.0, 2.0, 0.0])                               # last frame: cost <= 1e-6 -> dropped
    ok = ((t < 2) & (r < 5)).astype(np.int32)
    s = handoff.summarize(t, r, cost, ok)
    assert s["n"] == 3
    assert s["rte_mean"] == pytest.approx(np.mean(t[:3])) and s["rre_sigma"] == pytest.approx(np.std(r[:3]))
    assert s["success_rate"] == pytest.approx(1.0 / 3.0)
    assert handoff.summarize(t, r, np.zeros(4), ok)["n"] == 0

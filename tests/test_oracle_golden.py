"""Pins the oracle (oracle/resample.c + oracle/ref_path.py) against outputs of the reference's own
code (tests/golden/, written by oracle/gen_golden.py) and the SURVEY Appendix-D known answers."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_path, resample


@pytest.fixture(scope="module")
def crops(golden_dir):
    return np.load(os.path.join(golden_dir, "crops.npz"))


def test_c_oracle_square_crop_matches_reference(crops, golden_frames):
    n = len(crops["ok"])
    assert n > 200
    bad = []
    for i in range(n):
        frame = golden_frames[int(crops["frame_id"][i])]
        box, pad = tuple(crops["box"][i]), int(crops["padding"][i])
        try:
            ok, crop = resample.square_crop(frame, box, 128, pad)
            zd = 0
        except ZeroDivisionError:
            ok, crop, zd = False, None, 1
        if ok != bool(crops["ok"][i]) or zd != int(crops["zero_div"][i]):
            bad.append((i, "status"))
        elif ok and hashlib.sha256(crop.tobytes()).hexdigest() != str(crops["sha256"][i]):
            bad.append((i, "bytes"))
    assert not bad, bad[:10]


def test_library_port_square_crop_matches_reference(crops, golden_frames):
    for i in range(0, len(crops["ok"]), 3):
        if int(crops["zero_div"][i]):
            continue
        frame = golden_frames[int(crops["frame_id"][i])]
        ok, crop = ref_path.square_crop_libs(frame, tuple(crops["box"][i]), 128, int(crops["padding"][i]))
        assert ok == bool(crops["ok"][i])
        if ok:
            assert hashlib.sha256(crop.tobytes()).hexdigest() == str(crops["sha256"][i])


def test_full_crops_and_appendix_d(crops, golden_frames):
    for k, i in enumerate(crops["full_idx"]):
        ok, crop = resample.square_crop(golden_frames[int(crops["frame_id"][i])], tuple(crops["box"][i]), 128, int(crops["padding"][i]))
        assert ok and np.array_equal(crop, crops["full"][k])
    d1 = (0.673046875, 0.5368055555555555, 0.12890625, 0.2625)
    ok, c = resample.square_crop(golden_frames[0], d1, 128, 30)  # Appendix D3
    assert ok and int(c.sum()) == 6265750 and hashlib.sha256(c.tobytes()).hexdigest()[:16] == "96519aed41f98a17"
    assert list(c[0, 0]) == [114, 105, 160] and list(c[64, 64]) == [116, 125, 116] and list(c[127, 127]) == [122, 101, 126]
    ok, c = resample.square_crop(golden_frames[0], d1, 128, 0)
    assert int(c.sum()) == 6273125 and hashlib.sha256(c.tobytes()).hexdigest()[:16] == "d9d1b387db993224"
    ok, c = resample.square_crop(golden_frames[1], d1, 128, 30)  # Appendix D4
    assert int(c.sum()) == 7073522 and hashlib.sha256(c.tobytes()).hexdigest()[:16] == "aa056564e5580d3a"
    ok, c = resample.square_crop(golden_frames[1], d1, 128, 0)
    assert int(c.sum()) == 7117315 and hashlib.sha256(c.tobytes()).hexdigest()[:16] == "ef7e74d9e783187c"
    ok, c = resample.square_crop(golden_frames[1], (0.5, 0.5, 196 / 1920 + 1e-9, 100 / 1080), 128, 0)  # D5: 127-row quirk
    assert ok and c.shape == (128, 128, 3) and int(c[126].max()) == 254 and int(c[127].max()) == 0 and int(c.sum()) == 6311152


def test_edge_cases_appendix_d6(golden_frames):
    f = golden_frames[1]
    assert resample.square_crop(f, (1.3, 0.5, 0.128, 0.2625), 128, 30) == (False, None)
    ok, c = resample.square_crop(f, (0.02, 0.5, 0.128, 0.2625), 128, 30)
    assert ok and int(c[:, 0].max()) == 0 and int(c[:, 127].max()) == 0  # re-centred letterbox
    with pytest.raises(ZeroDivisionError):
        resample.square_crop(f, (0.5, 0.5, 0.0, 0.0), 128, 30)
    assert resample.square_crop(f, (0.5, 0.5, 0.0, 0.0), 128, 0) == (False, None)


def test_bbox_geometry_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "bbox.npz"))
    recs = json.loads(str(g["records"]))
    got = np.array([ref_path.fighter_box(r) for r in recs])
    assert np.array_equal(got, g["boxes"])
    assert tuple(got[-1]) == (0.673046875, 0.5368055555555555, 0.12890625, 0.2625)  # Appendix D1
    assert tuple(g["yolo_pixels"][-1]) == (1292, 579, 247, 283)


def test_windows_match_reference(golden_dir):
    w = json.load(open(os.path.join(golden_dir, "windows.json")))
    for (a, b, c, d, e), out in zip(w["args"], w["out"]):
        assert ref_path.middle_out(a, b, c, d, e) == out
    assert ref_path.middle_out(100, 7, 3, 1000, 1) == [73, 88, 97, 100, 103, 112, 127]  # Appendix D2
    assert ref_path.middle_out(2, 7, 3, 10, 1) == [1, 1, 1, 2, 5, 9, 9]
    with pytest.raises(AssertionError):
        ref_path.middle_out(3, 4, 1, 10)


def test_timeline_matches_reference(golden_dir):
    tl = json.load(open(os.path.join(golden_dir, "timeline.json")))
    for off, want in tl.items():
        gt = ref_path.load_ground_truth(os.path.join(golden_dir, "sample_log.jsonl"), log_offset=int(off))
        got = [[[r["num_frames_left"], r["fighter_id"], r["pos_x"]] for r in fr] for fr in gt]
        assert got == want


def test_model_matches_reference_default_init(golden_dir):
    from workloads import weights

    g = np.load(os.path.join(golden_dir, "model.npz"))
    actions = json.loads(str(g["actions"]))
    sd = weights.default_state_dict(0)
    assert sorted(sd.keys()) == sorted("model." + k if not k.startswith("model.") else k for k in json.loads(str(g["keys"])) if "accuracy" not in k)
    m = ref_path.RefCNNActionDetector(actions, 7).eval()
    m.load_state_dict(sd)
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 15347815  # Appendix D7
    x = torch.rand((3, 7, 3, 128, 128), generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        lp = m(x)
    assert np.allclose(lp.numpy(), g["logp"], atol=2e-5)
    assert torch.argmax(lp, dim=1).tolist() == g["pred"].tolist() == [59, 59, 59]


def test_resformer_restatement_matches_reference_golden(golden_dir):
    """SURVEY 8f rank 2 groundwork: oracle/ref_resformer.py (module composition and the written-out encoder) against
    log-probs recorded from the reference's ResnetTransformerDetector, incl. its attention across the batch axis."""
    import json
    import os

    import torch

    from oracle.ref_resformer import RefResnetTransformerDetector
    from playaid_core_b200.anim_ontology import ACTIONS

    g = np.load(os.path.join(golden_dir, "resformer.npz"))
    torch.manual_seed(0)
    m = RefResnetTransformerDetector(ACTIONS, sequence_length=7).eval()
    assert sorted(m.state_dict().keys()) == json.loads(str(g["keys"]))
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    assert np.allclose(m.model.freq_encoding.numpy(), g["freq_encoding"], atol=1e-7)
    x = torch.rand((2, 7, 3, 128, 128), generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        y2, y1 = m(x).numpy(), m(x[:1]).numpy()
        e2, e1 = m.model.forward_explicit(x).numpy(), m.model.forward_explicit(x[:1]).numpy()
    assert y2.shape == (2, 7, 63)
    assert np.abs(y2 - g["logp_b2"]).max() < 2e-5 and np.abs(y1 - g["logp_b1"]).max() < 2e-5
    assert np.abs(e2 - g["logp_b2"]).max() < 5e-5 and np.abs(e1 - g["logp_b1"]).max() < 5e-5
    assert np.abs(g["logp_b2"][:1] - g["logp_b1"]).max() > 1e-3      # the batch-composition dependence is real


def test_cfg4_slice_golden_labels_give_the_recorded_stats(golden_dir):
    """BASELINE cfg4 consumer side: the committed fp32-oracle labels of the 2 048-frame slice, through ai_output.yaml ->
    the REFERENCE's load_timeline_from_ai_output / update_fighters_from_timeline / Stats, reproduce the recorded
    Stats.stats digests. Needs /root/reference (build container only); the GPU test asserts label identity, which
    carries this result over to the GPU label stream."""
    import os

    import pytest

    if not os.path.isdir("/root/reference/playaid"):
        pytest.skip("reference tree not present")
    from oracle.gen_cfg4_golden import SLICE, reference_stats_digests, slice_boxes

    g = np.load(os.path.join(golden_dir, "cfg4_slice.npz"))
    assert g["label"].shape == (SLICE, 2) and len(np.unique(g["label"])) >= 20
    sha600, sha_slice = reference_stats_digests(g["label"].astype(np.int64), g["prob"], slice_boxes(SLICE))
    assert sha600 == str(g["stats_sha256_first600"]) and sha_slice == str(g["stats_sha256_slice"])

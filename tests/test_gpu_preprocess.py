"""GPU parity of the fused preprocess kernel (through the C-ABI) against the reference's golden
crops and the C oracle: bit-exact bytes for every resample regime, statuses, output formats."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _run_u8(torch, frames_np, boxes, frame_ids, padding, out_size=128, swap=False):
    from playaid_core_b200 import _lib
    from playaid_core_b200.preprocess import crop_records, preprocess_crops

    frames = torch.from_numpy(np.ascontiguousarray(frames_np)).cuda()
    H, W = frames.shape[1:3]
    rec = torch.from_numpy(crop_records(boxes, frame_ids, W, H)).cuda()
    out, status = preprocess_crops(frames, rec, out_size, padding, swap_rb=swap, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    torch.cuda.synchronize()
    return out.cpu().numpy(), status.cpu().numpy()


def test_golden_crops_bit_exact(torch_cuda, golden_dir, golden_frames):
    g = np.load(os.path.join(golden_dir, "crops.npz"))
    frames = np.stack(golden_frames)
    n = len(g["ok"])
    bad = []
    for pad in (0, 30):
        sel = np.nonzero(g["padding"] == pad)[0]
        out, status = _run_u8(torch_cuda, frames, g["box"][sel], g["frame_id"][sel], pad)
        for j, i in enumerate(sel):
            want_ok, zd = bool(g["ok"][i]), int(g["zero_div"][i])
            if zd:
                if status[j] != -2:
                    bad.append((int(i), "zero_div", int(status[j])))
            elif want_ok != (status[j] == 1):
                bad.append((int(i), "status", int(status[j])))
            elif want_ok and hashlib.sha256(out[j].tobytes()).hexdigest() != str(g["sha256"][i]):
                bad.append((int(i), "bytes", int(g["sum"][i]) - int(out[j].sum())))
    assert not bad, f"{len(bad)}/{n} golden crops differ: {bad[:12]}"


def test_random_boxes_vs_c_oracle(torch_cuda, golden_frames):
    from oracle import resample

    rng = np.random.default_rng(123)
    frames = np.stack(golden_frames)
    n = 600
    boxes = np.stack([rng.uniform(-0.05, 1.05, n), rng.uniform(-0.05, 1.05, n), rng.uniform(0.005, 0.6, n), rng.uniform(0.005, 0.9, n)], 1)
    fids = rng.integers(0, 3, n)
    for pad in (0, 30, 7):
        out, status = _run_u8(torch_cuda, frames, boxes, fids, pad)
        bad = []
        for i in range(n):
            try:
                ok, crop = resample.square_crop(frames[fids[i]], tuple(boxes[i]), 128, pad)
                want = 1 if ok else 0
            except ZeroDivisionError:
                crop, want = None, -2
            if status[i] != want:
                bad.append((i, "status", int(status[i]), want))
            elif want == 1 and not np.array_equal(out[i], crop):
                bad.append((i, "bytes", int(np.abs(out[i].astype(int) - crop).max())))
        assert not bad, f"pad={pad}: {len(bad)} mismatches {bad[:10]}"
        assert (status == -7).sum() == 0     # every window of a 1080p frame is computed (the reference computes any size)


def test_huge_boxes_are_computed(torch_cuda, golden_frames):
    """Boxes up to the whole frame (and beyond its edges): windows of up to 1 920 x 1 080 pixels go through the tensor-core
    path in row parts or through the streaming kernel's large-window pass -- never status -7 (fighter.py:336-355 computes
    any size)."""
    from oracle import resample

    rng = np.random.default_rng(77)
    frames = np.stack(golden_frames)
    n = 48
    boxes = np.stack([rng.uniform(0.2, 0.8, n), rng.uniform(0.2, 0.8, n), rng.uniform(0.5, 1.0, n), rng.uniform(0.5, 1.0, n)], 1)
    boxes[0] = (0.5, 0.5, 1.0, 1.0)
    boxes[1] = (0.5, 0.5, 0.999, 0.3)
    boxes[2] = (0.25, 0.75, 0.3, 0.999)
    fids = rng.integers(0, 3, n)
    for pad in (30, 0):
        out, status = _run_u8(torch_cuda, frames, boxes, fids, pad)
        for i in range(n):
            ok, crop = resample.square_crop(frames[fids[i]], tuple(boxes[i]), 128, pad)
            assert status[i] == (1 if ok else 0), (pad, i, int(status[i]))
            if ok:
                assert np.array_equal(out[i], crop), (pad, i)


def test_other_output_sizes(torch_cuda, golden_frames):
    from oracle import resample

    frames = np.stack(golden_frames[:2])
    rng = np.random.default_rng(5)
    boxes = np.stack([rng.uniform(0.1, 0.9, 40), rng.uniform(0.1, 0.9, 40), rng.uniform(0.02, 0.3, 40), rng.uniform(0.03, 0.4, 40)], 1)
    fids = rng.integers(0, 2, 40)
    for out_size in (64, 100, 224):
        out, status = _run_u8(torch_cuda, frames, boxes, fids, 30, out_size=out_size)
        for i in range(40):
            ok, crop = resample.square_crop(frames[fids[i]], tuple(boxes[i]), out_size, 30)
            assert ok == (status[i] == 1)
            if ok:
                assert np.array_equal(out[i], crop), (out_size, i)


def test_output_formats(torch_cuda, golden_frames):
    """BGR->RGB, HWC->CHW, /255 exactly like ai_runner.py:448,461-463; bf16 hi/lo planes; mean/std."""
    torch = torch_cuda
    from playaid_core_b200 import _lib
    from playaid_core_b200.preprocess import crop_records, preprocess_crops

    frames = torch.from_numpy(np.stack(golden_frames)).cuda()
    rng = np.random.default_rng(9)
    boxes = np.stack([rng.uniform(0.1, 0.9, 16), rng.uniform(0.1, 0.9, 16), rng.uniform(0.05, 0.3, 16), rng.uniform(0.05, 0.4, 16)], 1)
    rec = torch.from_numpy(crop_records(boxes, rng.integers(0, 3, 16), 1920, 1080)).cuda()
    u8, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    # cvtColor(BGR2RGB) + permute + .float()/255.0 evaluated on the CPU like the reference (torch's CUDA
    # scalar division multiplies by a reciprocal and differs in the last bit for 126 of 256 values)
    want = (u8.cpu().flip(-1).permute(0, 3, 1, 2).float() / 255.0).cuda()
    f32, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_F32, layout=_lib.LAYOUT_NCHW)
    assert torch.equal(f32, want)
    nhwc, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_F32, layout=_lib.LAYOUT_NHWC)
    assert torch.equal(nhwc.permute(0, 3, 1, 2), want)
    b4, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_BF16, layout=_lib.LAYOUT_NHWC4)
    assert torch.equal(b4[..., :3].permute(0, 3, 1, 2), want.to(torch.bfloat16)) and float(b4[..., 3].abs().max()) == 0.0
    x2, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_BF16X2, layout=_lib.LAYOUT_NHWC4)
    hi, lo = x2[0, ..., :3].permute(0, 3, 1, 2).float(), x2[1, ..., :3].permute(0, 3, 1, 2).float()
    assert torch.equal(hi, want.to(torch.bfloat16).float())
    assert float((hi + lo - want).abs().max()) < 2e-5
    h2, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_F16X2, layout=_lib.LAYOUT_NHWC4)
    hh, hl = h2[0, ..., :3].permute(0, 3, 1, 2), h2[1, ..., :3].permute(0, 3, 1, 2)
    assert h2.dtype == torch.float16 and torch.equal(hh, want.half()) and float((hh.float() + hl.float() - want).abs().max()) < 2e-7
    h1, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_F16, layout=_lib.LAYOUT_NHWC4)
    assert torch.equal(h1[..., :3].permute(0, 3, 1, 2), want.half())
    pp, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_F16X2, layout=_lib.LAYOUT_NHWC4P)
    assert tuple(pp.shape) == (2, 16, 128, 136, 4) and torch.equal(pp[:, :, :, 4:132], h2)   # conv1-ready padded layout
    assert float(pp[:, :, :, :4].abs().max()) == 0.0 and float(pp[:, :, :, 132:].abs().max()) == 0.0
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    ms, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, mean=mean, std=std, dtype=_lib.DTYPE_F32, layout=_lib.LAYOUT_NCHW)
    m = torch.tensor(mean, device="cuda").view(1, 3, 1, 1)
    s = torch.tensor(std, device="cuda").view(1, 3, 1, 1)
    assert torch.equal(ms.cpu(), (want.cpu() - m.cpu()) / s.cpu())
    sw_u8, _ = preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NCHW)
    assert torch.equal(sw_u8, u8.flip(-1).permute(0, 3, 1, 2))


def test_yolocrop_square_crop_dropin(torch_cuda, golden_frames):
    """Same call as the reference: crop.square_crop(frame, 128, padding=30) -> (ok, ndarray)."""
    from playaid_core_b200.fighter import YoloCrop

    d1 = YoloCrop(0.673046875, 0.5368055555555555, 0.12890625, 0.2625)
    ok, c = d1.square_crop(golden_frames[0], 128, padding=30)
    assert ok and c.shape == (128, 128, 3) and c.dtype == np.uint8
    assert int(c.sum()) == 6265750 and hashlib.sha256(c.tobytes()).hexdigest()[:16] == "96519aed41f98a17"  # Appendix D3
    ok, c = YoloCrop(0.5, 0.5, 196 / 1920 + 1e-9, 100 / 1080).square_crop(golden_frames[1], 128)
    assert ok and int(c[127].max()) == 0 and int(c.sum()) == 6311152  # Appendix D5
    assert YoloCrop(1.3, 0.5, 0.128, 0.2625).square_crop(golden_frames[1], 128, padding=30) == (False, None)
    with pytest.raises(ZeroDivisionError):
        YoloCrop(0.5, 0.5, 0.0, 0.0).square_crop(golden_frames[1], 128, padding=30)


def test_full_size_batch_properties(torch_cuda):
    """BASELINE cfg2 batch (256 frames x 2 fighters at 1080p): every crop valid, sampled crops equal
    the oracle, and the kernel is a pure function of (frame, box) -- identical on a second launch and
    when the same crops are requested in reversed order."""
    torch = torch_cuda
    from oracle import resample
    from playaid_core_b200 import _lib
    from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
    from playaid_core_b200.preprocess import crop_records, preprocess_crops
    from workloads import synthetic

    N = 256
    recs = synthetic.synth_log_records(N, 2, seed=2024)
    boxes = boxes_from_records([r for f in recs for r in f]).reshape(N, 2, 4)
    px = yolo_pixels_batch(boxes, 1920, 1080)
    frames = synthetic.synth_frames(np.arange(N), px, device="cuda")
    fid = np.repeat(np.arange(N), 2)
    rec = crop_records(boxes.reshape(-1, 4), fid, 1920, 1080)
    rec_d = torch.from_numpy(rec).cuda()
    a, st = preprocess_crops(frames, rec_d, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    b, _ = preprocess_crops(frames, rec_d, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    assert torch.equal(a, b)
    rev = torch.from_numpy(rec[::-1].copy()).cuda()
    c, _ = preprocess_crops(frames, rev, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    assert torch.equal(a, c.flip(0))
    st = st.cpu().numpy()
    assert (st == 1).all(), np.unique(st, return_counts=True)
    a_np = a.cpu().numpy()
    for i in range(0, 2 * N, 16):
        f = frames[fid[i]].cpu().numpy()
        ok, crop = resample.square_crop(f, tuple(boxes.reshape(-1, 4)[i]), 128, 30)
        assert ok and np.array_equal(a_np[i], crop), i


def test_window_staging_matches_device_frames(torch_cuda, golden_dir, golden_frames):
    """pa_stage_windows: crops cut from a staging buffer that holds only the window bytes (pulled from
    pinned host frames, rest of the buffer poisoned) are the bytes cut from fully resident frames."""
    torch = torch_cuda
    from playaid_core_b200 import _lib
    from playaid_core_b200.preprocess import crop_records, preprocess_crops, stage_windows

    g = np.load(os.path.join(golden_dir, "crops.npz"))
    frames_np = np.ascontiguousarray(np.stack(golden_frames))
    H, W = frames_np.shape[1:3]
    host = torch.from_numpy(frames_np).pin_memory()
    dev_full = torch.from_numpy(frames_np).cuda()
    for pad in (0, 30):
        sel = np.nonzero(g["padding"] == pad)[0]
        base = 5   # the match-wide record table carries global frame numbers
        rec_global = crop_records(g["box"][sel], g["frame_id"][sel] + base, W, H)
        rec_local = rec_global.copy(); rec_local[:, 0] -= base
        rec_g, rec_l = torch.from_numpy(rec_global).cuda(), torch.from_numpy(rec_local).cuda()
        staged = torch.full(tuple(host.shape), 0xAB, dtype=torch.uint8, device="cuda")
        stage_windows(host, rec_g, staged, padding=pad, frame_base=base)
        a, sa = preprocess_crops(staged, rec_l, 128, pad, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
        b, sb = preprocess_crops(dev_full, rec_l, 128, pad, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
        torch.cuda.synchronize()
        assert torch.equal(sa, sb)
        assert torch.equal(a, b), f"padding {pad}: {(a != b).flatten(1).any(1).sum().item()} crops differ"


def test_empty_record_lists_are_no_ops(torch_cuda, golden_frames):
    """Zero crops / zero log records: valid calls that launch nothing (a chunk in which no fighter is on screen)."""
    torch = torch_cuda
    from playaid_core_b200 import _lib
    from playaid_core_b200.fighter import boxes_from_records_device
    from playaid_core_b200.preprocess import preprocess_crops, stage_windows

    frames = torch.from_numpy(np.stack(golden_frames)).cuda()
    rec = torch.empty((0, _lib.BOX_STRIDE), dtype=torch.int32, device="cuda")
    before = _lib.Context.get().launch_count()
    out, st = preprocess_crops(frames, rec, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    assert out.shape[0] == 0 and st.shape[0] == 0
    host = frames.cpu().pin_memory()
    stage_windows(host, rec, torch.empty_like(frames), padding=30, frame_base=0)
    boxes, recs = boxes_from_records_device(np.empty((0, _lib.LOG_STRIDE), np.float64), 1920, 1080)
    assert boxes.shape[0] == 0 and recs.shape[0] == 0
    torch.cuda.synchronize()
    assert _lib.Context.get().launch_count() == before


def test_window_staging_overlapping_windows(torch_cuda):
    """Windows of consecutive records in one frame share bytes; the staging kernel pulls the shared part once (cutting a
    window back where the previous record's window reaches over its left or right end). Every byte of every window must
    still arrive: overlaps to the left / right / above / below, containment both ways, identical windows, a different
    frame in between."""
    torch = torch_cuda
    from playaid_core_b200.preprocess import stage_windows

    Hh, Ww, pad = 540, 960, 30
    rng = np.random.default_rng(7)
    host_np = rng.integers(0, 256, (3, Hh, Ww, 3), dtype=np.uint8)
    host = torch.from_numpy(host_np).pin_memory()
    # crop records: frame, cx, cy, w, h (+ padding to the record stride)
    base_boxes = [(0, 300, 250, 200, 180), (0, 380, 260, 200, 180), (0, 220, 240, 200, 180), (0, 300, 330, 200, 180),
                  (0, 300, 170, 200, 180), (0, 300, 250, 80, 60), (0, 300, 250, 320, 300), (0, 300, 250, 320, 300),
                  (1, 300, 250, 200, 180), (0, 310, 255, 200, 180), (2, 20, 20, 200, 180), (2, 60, 30, 120, 100),
                  (2, 940, 520, 200, 180), (2, 900, 500, 200, 180)]
    from playaid_core_b200 import _lib
    rec = np.zeros((len(base_boxes), _lib.BOX_STRIDE), np.int32)
    for i, b in enumerate(base_boxes):
        rec[i, :5] = b
    staged = torch.full(tuple(host.shape), 0xAB, dtype=torch.uint8, device="cuda")
    stage_windows(host, torch.from_numpy(rec).cuda(), staged, padding=pad, frame_base=0)
    torch.cuda.synchronize()
    got = staged.cpu().numpy()
    for f, cx, cy, w, h in base_boxes:
        sd = max(w, h); half = sd // 2
        y0, y1 = max(cy - half - pad, 0), min(cy + half + pad, Hh)
        x0, x1 = max(cx - half - pad, 0), min(cx + half + pad, Ww)
        assert np.array_equal(got[f, y0:y1, x0:x1], host_np[f, y0:y1, x0:x1]), (f, cx, cy, w, h)


def test_match_stream_host_modes_agree(torch_cuda):
    """MatchStream fed pinned host chunks (window staging on the copy stream, and in-place reads) labels
    the clip exactly like the device-resident path."""
    torch = torch_cuda
    from playaid_core_b200.action_detector import ActionDetector
    from playaid_core_b200.anim_ontology import ACTIONS
    from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
    from playaid_core_b200.models.cnn_action_detector import CNNActionDetector
    from workloads import synthetic, weights

    n, Hh, Ww = 96, 540, 960
    recs = synthetic.synth_log_records(n, 2, seed=11)
    boxes = boxes_from_records([r for f in recs for r in f]).reshape(n, 2, 4)
    frames = synthetic.synth_frames(np.arange(n), yolo_pixels_batch(boxes, Ww, Hh), H=Hh, W=Ww, device="cuda", seed=5)
    host = frames.cpu().pin_memory()
    model = CNNActionDetector(ACTIONS, sequence_length=7, precision="f16", device="cuda").eval()
    model.load_state_dict(weights.default_state_dict(0))
    det = ActionDetector(model)
    outs = {}
    for mode in ("device", "stage", "inplace"):
        det.host_mode = mode if mode != "device" else "stage"
        st = det.stream(boxes, Hh, Ww)
        src = frames if mode == "device" else host
        for c in range(0, n, 32):
            st.push(src[c : c + 32])
        torch.cuda.synchronize()
        outs[mode] = (st.label.clone(), st.logp.clone(), st.status.clone())
    for mode in ("stage", "inplace"):
        assert torch.equal(outs[mode][2], outs["device"][2])
        assert torch.equal(outs[mode][0], outs["device"][0])
        assert torch.equal(outs[mode][1], outs["device"][1]), mode


def test_unaligned_frame_geometry(torch_cuda):
    """Frames whose rows are not 16-byte multiples (1001 px wide: 3003-byte pitch) and pitched views of wider
    frames take the byte-granular load paths of pa_preprocess and pa_stage_windows: still bit-exact vs the C oracle."""
    torch = torch_cuda
    from oracle import resample
    from playaid_core_b200 import _lib
    from playaid_core_b200.preprocess import crop_records, preprocess_crops, stage_windows

    rng = np.random.default_rng(5)
    Hh, Ww = 477, 1001
    yy, xx = np.mgrid[0:Hh, 0:Ww]
    base = np.stack([(xx * 3 + yy) % 256, (xx + yy * 2) % 256, (xx * yy // 7) % 256], -1).astype(np.uint8)
    frames = np.stack([base, np.roll(base, 37, 1), (rng.integers(0, 256, base.shape)).astype(np.uint8)])
    n = 160
    boxes = np.stack([rng.uniform(0.0, 1.0, n), rng.uniform(0.0, 1.0, n), rng.uniform(0.02, 0.5, n), rng.uniform(0.02, 0.7, n)], 1)
    fids = rng.integers(0, 3, n)
    out, status = _run_u8(torch, frames, boxes, fids, 30)
    # a pitched view: the same pixels inside wider rows (row pitch 3072 B = 16-byte multiple, but x offsets are not)
    wide = torch.zeros((3, Hh, 1024, 3), dtype=torch.uint8, device="cuda")
    wide[:, :, :Ww] = torch.from_numpy(frames).cuda()
    view = wide[:, :, :Ww]
    rec = torch.from_numpy(crop_records(boxes, fids, Ww, Hh)).cuda()
    out_v, status_v = preprocess_crops(view, rec, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    # staging from pinned host frames with the unaligned pitch
    host = torch.from_numpy(frames).pin_memory()
    staged = torch.full(tuple(host.shape), 0x5A, dtype=torch.uint8, device="cuda")
    stage_windows(host, rec, staged, padding=30, frame_base=0)
    out_s, status_s = preprocess_crops(staged, rec, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    torch.cuda.synchronize()
    bad = []
    for i in range(n):
        try:
            ok, crop = resample.square_crop(frames[fids[i]], tuple(boxes[i]), 128, 30)
            want = 1 if ok else 0
        except ZeroDivisionError:
            crop, want = None, -2
        if status[i] != want:
            bad.append((i, "status", int(status[i]), want))
        elif want == 1 and not np.array_equal(out[i], crop):
            bad.append((i, "bytes"))
    assert not bad, bad[:10]
    assert np.array_equal(status_v.cpu().numpy(), status) and np.array_equal(out_v.cpu().numpy(), out)
    assert np.array_equal(status_s.cpu().numpy(), status) and np.array_equal(out_s.cpu().numpy(), out)


def test_device_boxes_bit_equal_to_reference_and_host(torch_cuda, golden_dir):
    """`pa_boxes_from_log` (fp64 camera geometry + np.round + int() truncation on the device) against
    the 601 boxes recorded from the reference's Fighter(...).crop (tests/golden/bbox.npz) and against the vectorised host
    path on 60 000 synthetic ult_logger records: boxes bit-equal, crop records identical."""
    import json

    from playaid_core_b200.fighter import boxes_from_records, boxes_from_records_device, log_record_array, yolo_pixels_batch
    from workloads import synthetic

    g = np.load(os.path.join(golden_dir, "bbox.npz"))
    recs = json.loads(str(g["records"]))
    boxes, crops = boxes_from_records_device(log_record_array(recs), 1920, 1080)
    assert np.array_equal(boxes.cpu().numpy(), g["boxes"])
    assert np.array_equal(crops.cpu().numpy()[:, 1:5], g["yolo_pixels"])
    assert np.array_equal(crops.cpu().numpy()[:, 0], np.arange(len(recs)))
    n = 30000
    log = synthetic.synth_log_records(n, 2, seed=77, stage_id=95)      # fov 30 stage
    log += synthetic.synth_log_records(n // 2, 2, seed=78, stage_id=0, pos_x_range=(-80.0, 80.0), pos_y_range=(0.0, 60.0))
    flat = [r for f in log for r in f]
    want = boxes_from_records(flat)
    frame_idx = np.repeat(np.arange(len(log)), 2)
    got, rec = boxes_from_records_device(log_record_array(flat, frame_idx), 1920, 1080)
    assert np.array_equal(got.cpu().numpy(), want)
    rec = rec.cpu().numpy()
    assert np.array_equal(rec[:, 1:5], yolo_pixels_batch(want, 1920, 1080)) and np.array_equal(rec[:, 0], frame_idx) and not rec[:, 5:].any()

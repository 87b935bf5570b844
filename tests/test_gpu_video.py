"""NVDEC frame source (reference transforms/video.py:23-45) against cv2.VideoCapture on the SAME file.

The decoded Y / U / V samples of a conforming stream are normative; the BGR frames differ in the last step: cv2 converts with
swscale (BT.601 limited range, its own chroma up-sampling and rounding), libippvideo with the integer BT.601 matrix and nearest
chroma.  Tolerance (stated, measured on these clips): every byte within 6 levels on smooth content, mean absolute difference
below 1.5 levels.  The tests are skipped where libnvcuvid is missing or the GPU does not decode the codec."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")


def _clip(path, fourcc, n=14, w=320, h=240):
    """smooth moving content (gradients + a soft disc): chroma edges are where the two conversions differ most"""
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), 25.0, (w, h))
    if not wr.isOpened():
        return False
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        img = np.stack([60 + 120 * xx / w, 40 + 150 * yy / h, 200 - 100 * (xx + yy) / (w + h)], axis=-1)
        cx, cy = 60 + 12 * i, 80 + 5 * i
        disc = np.clip(1.5 - np.hypot(xx - cx, yy - cy) / 30.0, 0, 1)[..., None]
        img = img * (1 - disc) + disc * np.array([40, 220, 230], np.float32)
        wr.write(np.clip(img, 0, 255).astype(np.uint8))
    wr.release()
    return path.stat().st_size > 0


def _host_frames(path):
    cap = cv2.VideoCapture(str(path))
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    cap.release()
    return np.stack(out)


@pytest.mark.parametrize("fourcc,ext,codec", [("VP90", "webm", 10), ("mp4v", "mp4", 2)])
def test_nvdec_frames_match_host_decode(tmp_path, fourcc, ext, codec):
    from image_processor_pipeline_b200 import nvdec
    if not nvdec.supported(codec):
        pytest.skip(f"NVDEC path unavailable here: {nvdec._err()}")
    path = tmp_path / f"clip.{ext}"
    if not _clip(path, fourcc):
        pytest.skip(f"this cv2 build cannot write {fourcc}")
    ref = _host_frames(path)
    got = torch.cat(list(nvdec.frame_batches(path, batch=5))).cpu().numpy()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 6, diff.max()
    assert diff.mean() < 1.5, diff.mean()


def test_nvdec_frames_feed_the_chain(tmp_path):
    """the decoded batches are what stage A consumes: one fused chain over NVDEC frames runs and yields images"""
    import random
    from image_processor_pipeline_b200 import chain, nvdec, synth
    if not nvdec.supported(10):
        pytest.skip(f"NVDEC path unavailable here: {nvdec._err()}")
    path = tmp_path / "clip.webm"
    if not _clip(path, "VP90", n=6):
        pytest.skip("this cv2 build cannot write VP90")
    frames = next(iter(nvdec.frame_batches(path, batch=6)))
    bgs = torch.from_numpy(np.stack([synth.background(3, i, 240, 320) for i in range(2)])).to(frames.device)
    random.seed(5)
    res = chain.run_chain(frames, bgs, chain.ChainParams(color_ranges_to_exclude_hsv=[(0, 0, 0, 180, 255, 150)], num_rotations=1))
    assert res.images.shape[0] == len(res.labels) and res.images.shape[0] >= 1

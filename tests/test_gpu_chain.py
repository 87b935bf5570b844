"""Parity of the fused device chain against the oracle chain (the reference's steps
run one after the other with the reference's own cv2 / Pillow calls).  All integer:
images bit-exact, label files string-identical, same draw order from ``random``."""
import random

import numpy as np
import pytest
import torch

from oracle import chain as OC

pytestmark = pytest.mark.gpu

chain = pytest.importorskip("image_processor_pipeline_b200.chain")
from image_processor_pipeline_b200 import synth  # noqa: E402

DEV = "cuda:0"


def _run_both(frames, bgs, seed, run_kw=None, **kw):
    random.seed(seed)
    ref_imgs, ref_labels, ref_recs = OC.run_chain(frames, bgs, **kw)
    state_after_ref = random.getstate()
    random.seed(seed)
    res = chain.run_chain(torch.from_numpy(np.stack(frames)).to(DEV), torch.from_numpy(np.stack(bgs)).to(DEV),
                          chain.ChainParams(**kw), **(run_kw or {}))
    assert random.getstate() == state_after_ref, "the chain must consume the RNG exactly like the reference"
    return res, ref_imgs, ref_labels, ref_recs


def _compare(res, ref_imgs, ref_labels, ref_recs):
    assert len(res.labels) == len(ref_labels) == res.images.shape[0]
    out = res.images.cpu().numpy()
    for i, (img, lab, rec) in enumerate(zip(ref_imgs, ref_labels, ref_recs)):
        mine = res.records[i]
        assert (mine["frame"], mine["rot_index"], mine["sym"], mine["bg"]) == (rec["frame"], rec["rot_index"], rec["sym"], rec["bg"])
        assert tuple(mine["new"]) == tuple(rec["new"]) and tuple(mine["pos"]) == tuple(rec["pos"]) and tuple(mine["sq"]) == tuple(rec["sq"])
        assert res.labels[i] == lab
        assert np.array_equal(out[i], img), f"image {i} differs ({(out[i] != img).sum()} bytes)"


@pytest.mark.parametrize("cfg", [
    dict(num_rotations=2, include_original=True, pool=["o", "h", "v", "hv"], choose_random=2),
    dict(num_rotations=1, include_original=False, pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False),
    dict(num_rotations=3, include_original=False, angle_min=80.0, angle_max=100.0, pool=["hv"], sym_include_original=True,
         scale_min=0.4, scale_max=0.8, yolo_class_id=3),
])
@pytest.mark.parametrize("run_kw", [dict(planner="native"), dict(planner="python"), dict(planner="native", device_tables=False),
                                    dict(planner="native", ambiguity_eps=0.2)])
def test_chain_small(cfg, run_kw):
    H, W = 270, 480
    frames = [synth.frame(5, i, H, W) for i in range(4)]
    bgs = [synth.background(5, i, H, W) for i in range(3)]
    kw = dict(crop_margins=(0.05, 4, 0.02, 6), color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **cfg)
    res, *ref = _run_both(frames, bgs, 3, run_kw=run_kw, **kw)
    _compare(res, *ref)
    if run_kw.get("ambiguity_eps"):
        assert res.tables_flagged_rows > 0          # the exact-row verification path really ran


def test_chain_zones_gimp_and_portrait_background():
    H, W = 200, 320
    frames = [synth.frame(9, i, H, W, speck_density=0.03) for i in range(3)]
    bgs = [synth.background(9, i, 300, 220) for i in range(2)]          # portrait: square crop is 220 wide
    kw = dict(crop_margins=(0, 0, 0, 0), color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS,
              zones=[(0, 0, 0, 0), (3, 4, 5, 6), None, (10, 0, 0, 20)], num_rotations=2, include_original=False,
              pool=["h", "v"], choose_random=1, sym_include_original=False)
    _compare(*_run_both(frames, bgs, 11, **kw))


def test_chain_dropped_items_and_noise_frames():
    # frame 1 is keyed out entirely (no component) and disappears; frame 2 is pure noise (worst-case CCL)
    H, W = 160, 200
    frames = [synth.frame(13, 0, H, W), np.zeros((H, W, 3), np.uint8), synth.noise(13, 2, (H, W, 3)), synth.frame(13, 3, H, W)]
    bgs = [synth.background(13, i, 240, 260) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=True,
              pool=["o", "hv"], choose_random=1, sym_include_original=False)
    res, ref_imgs, ref_labels, ref_recs = _run_both(frames, bgs, 17, **kw)
    assert all(r["frame"] != 1 for r in res.records)
    _compare(res, ref_imgs, ref_labels, ref_recs)


def test_chain_1080p():
    H, W = 1080, 1920
    frames = [synth.frame(21, i, H, W) for i in range(2)]
    bgs = [synth.background(21, i, H, W) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    _compare(*_run_both(frames, bgs, 23, **kw))


def test_chain_1080p_uniform_noise_frames():
    """SURVEY.md §8d's generic generator: uniform-noise frames.  The key then keeps ~3/4 of the pixels as one giant
    component (plus thousands of small ones), the sprite is nearly the whole frame: the worst case of every stage."""
    H, W = 1080, 1920
    frames = [synth.noise(91, i, (H, W, 3)) for i in range(2)]
    bgs = [synth.background(91, i, H, W) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    _compare(*_run_both(frames, bgs, 24, **kw))


def test_chain_symmetry_pool_with_repeated_entries():
    """symmetry.py:97-122 takes any pool: repeated entries are positions random.sample draws among (a variant drawn twice is
    written twice under the same name).  Such pools are planned by the Python planner automatically."""
    H, W = 160, 240
    frames = [synth.frame(97, i, H, W) for i in range(3)]
    bgs = [synth.background(97, i, 200, 260) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=2, include_original=False,
              pool=["h", "h", "v", "hv", "o", "v"], choose_random=3, sym_include_original=False)
    _compare(*_run_both(frames, bgs, 27, **kw))


def test_chain_with_enhance_step():
    """The chain followed by tranfo.py's enhance_image under the 'sample' pairing (30 % blurred, 30 % channel tables):
    images, labels and the generator state equal the oracle's — plain and software-pipelined."""
    H, W = 200, 320
    frames = [synth.frame(95, i, H, W) for i in range(4)]
    bgs = [synth.background(95, i, 240, 300) for i in range(3)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=2, include_original=True,
              pool=["o", "h", "v", "hv"], choose_random=2, enhance=True)
    res, ref_imgs, ref_labels, ref_recs = _run_both(frames, bgs, 26, **kw)
    _compare(res, ref_imgs, ref_labels, ref_recs)
    assert len(ref_imgs) >= 10
    fd, bd = torch.from_numpy(np.stack(frames)).to(DEV), torch.from_numpy(np.stack(bgs)).to(DEV)
    random.seed(26)
    pipe = chain.ChainStream(chain.ChainParams(**kw))
    outs = list(pipe.map([(fd, bd), (fd, bd), (fd, bd)]))
    random.seed(26)
    for r in outs:
        imgs, labels, _ = OC.run_chain(frames, bgs, **kw)
        assert labels == r.labels and all(np.array_equal(a, b) for a, b in zip(r.images.cpu().numpy(), imgs))


def test_chain_odd_sizes_exact_length_buffers():
    """Frames whose byte length is not a multiple of four (333 x 217 x 3 x n): the gathers fetch the aligned 32-bit word(s)
    holding a pixel, and every such word holds at least one byte of the pixel, so nothing beyond the aligned extent of the
    caller's buffer is ever read (run under compute-sanitizer without the caching allocator: profiles/r02/memcheck*)."""
    H, W = 217, 333
    frames = [synth.frame(93, i, H, W, speck_density=0.02) for i in range(3)]
    bgs = [synth.background(93, i, 301, 275) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=2, include_original=True,
              pool=["o", "h", "v", "hv"], choose_random=2)
    _compare(*_run_both(frames, bgs, 25, **kw))


def test_chain_4k_frames_config5():
    """BASELINE.json config 5: the full chain over 3840x2160 frames (a video.py frame stream) with 4K backgrounds:
    2160x2160 outputs, sprites that need more than one 64-wide super-block in the LANCZOS passes."""
    H, W = 2160, 3840
    frames = [synth.frame(71, i, H, W, speck_density=0.002) for i in range(2)]
    bgs = [synth.background(71, i, H, W) for i in range(2)]
    kw = dict(crop_margins=(0.01, 8, 0.02, 0), color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1,
              include_original=False, pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False,
              scale_min=0.05, scale_max=0.12)
    res, *ref = _run_both(frames, bgs, 29, **kw)
    _compare(res, *ref)
    assert res.images.shape[1:] == (2160, 2160, 3)
    nsb = res._arrays[0]["nkb"]
    assert ((nsb & 0xffff) > 1).any() or ((nsb >> 16) > 1).any()        # the multi-super-block path really ran


def test_device_lanczos_tables_equal_host_tables():
    """Device-computed coefficient tables (CUDA sin) vs libm tables for many size pairs; every row that
    differs must have been flagged as ambiguous (then the host fixes it) — in practice none differs."""
    import ctypes as C
    from image_processor_pipeline_b200 import _native
    lib = _native.load()
    r = np.random.default_rng(5)
    k = 200
    items = np.zeros(k, dtype=chain.ITEM_DTYPE)
    items["spr_w"] = r.integers(2, 900, k); items["spr_h"] = r.integers(2, 900, k)
    items["new_w"] = r.integers(1, 700, k); items["new_h"] = r.integers(1, 700, k)
    items["new_w"][:5] = items["spr_w"][:5]                       # some skipped passes
    tot = chain._layout(items)
    host = chain._host_tables(items, lib, tot["table_ints"])
    items_d = torch.from_numpy(items.view(np.uint8)).to(DEV)
    tab_d = torch.zeros(tot["table_ints"], dtype=torch.int32, device=DEV)
    amb_d = torch.zeros(chain.amb_ints(), dtype=torch.int32, device=DEV)
    _native.check(lib.ipp_chain_tables(items_d.data_ptr(), k, tot["max_rows"], tab_d.data_ptr(), amb_d.data_ptr(),
                                       chain.AMB_CAPACITY, chain.AMBIGUITY_EPS, torch.cuda.current_stream().cuda_stream), "tables")
    dev_tab = tab_d.cpu().numpy()
    n_amb = int(amb_d[0].item())
    flagged = set(map(tuple, amb_d[chain.AMB_REC:chain.AMB_REC * (n_amb + 1)].cpu().numpy().reshape(-1, chain.AMB_REC)[:, :3].tolist()))
    bad = np.flatnonzero(dev_tab != host)
    for idx in bad:                                               # map the flat index back to (item, axis, row)
        i = int(np.searchsorted(items["hb_off"] + (items["h_ksize"] == 0) * 0, idx, side="right") - 1)
        found = False
        for it_i in range(k):
            it = items[it_i]
            for axis, (ko, ks, n_out) in enumerate(((int(it["hk_off"]), int(it["h_ksize"]), int(it["new_w"])),
                                                    (int(it["vk_off"]), int(it["v_ksize"]), int(it["new_h"])))):
                if ks and ko <= idx < ko + ks * n_out:
                    assert (it_i, axis, (idx - ko) // ks) in flagged, "a differing row was not flagged"
                    found = True
        assert found, "bounds differ: they involve no transcendental and must be identical"
    assert len(bad) <= 2 and n_amb < 200, (len(bad), n_amb)


def test_chain_two_kernel_gather_and_hpass_still_match(monkeypatch):
    """Stage B2 has two modes: gather + h pass fused through shared memory (default whenever the sprites are narrow
    enough) and two kernels exchanging the planar sprite through global memory.  Force the second one."""
    monkeypatch.setenv("IPP_B2_UNFUSED", "1")
    H, W = 270, 480
    frames = [synth.frame(33, i, H, W) for i in range(3)]
    bgs = [synth.background(33, i, H, W) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=2, include_original=True,
              pool=["o", "h", "v", "hv"], choose_random=2, sym_include_original=False)
    for run_kw in (dict(), dict(ambiguity_eps=0.2)):
        _compare(*_run_both(frames, bgs, 6, run_kw=run_kw, **kw))


def test_chain_generic_mask_kernel_still_matches(monkeypatch):
    """Stage A has two mask kernels: the persistent table-driven one (no zones, <= 8 filters) and the
    generic one.  Force the generic one on a configuration the fast one would take."""
    monkeypatch.setenv("IPP_MASK_GENERIC", "1")
    H, W = 270, 480
    frames = [synth.frame(31, i, H, W) for i in range(3)] + [synth.noise(31, 9, (H, W, 3))]
    bgs = [synth.background(31, i, H, W) for i in range(2)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    _compare(*_run_both(frames, bgs, 5, **kw))


def test_chain_mask_fast_path_on_all_colours_and_hue_filters():
    """Every BGR colour through the table-driven mask kernel, with filters that need the hue test
    (real H ranges incl. the wrap-free red band) — against cv2 via the oracle chain's first stage."""
    from oracle import libcalls as L
    c = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], -1).astype(np.uint8).reshape(16, 1024, 1024, 3)
    filters = [(0, 0, 0, 180, 255, 40), (15, 60, 200, 35, 255, 255), (100, 30, 30, 130, 255, 255), (0, 200, 100, 8, 255, 255),
               (170, 100, 100, 180, 255, 255), (40, 0, 240, 90, 20, 255)]
    from image_processor_pipeline_b200 import _native, hostmath
    import ctypes as C
    lib = _native.load()
    geom = chain._geom_struct(1024, 1024, (0, 0, 1024, 1024), 1024, 1024, hostmath.pack_hsv_ranges(filters))
    n = 16
    ws_bytes = lib.ipp_chain_mask_workspace_bytes(C.byref(geom), n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    info = torch.empty((n, 8), dtype=torch.int32, device=DEV)
    fr = torch.from_numpy(bgr).to(DEV)
    _native.check(lib.ipp_chain_stage_a(fr.data_ptr(), 1024 * 3, 1024 * 1024 * 3, C.byref(geom), info.data_ptr(), ws.data_ptr(),
                                        ws_bytes, n, torch.cuda.current_stream().cuda_stream), "stage_a")
    n_fg = info.cpu().numpy()[:, 5]
    for i in range(n):
        keyed = L.color_masks_to_alpha(bgr[i], filters)
        assert n_fg[i] == int((keyed[..., 3] > 1).sum()), i


def test_run_chain_host_whole_and_chunked():
    """Host-buffer entry point: one chain over the batch == run_chain; the pipelined (chunked) mode ==
    the oracle run slice after slice with the same generator."""
    H, W = 200, 320
    frames = [synth.frame(41, i, H, W) for i in range(6)]
    bgs = [synth.background(41, i, 240, 300) for i in range(6)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    fh = torch.from_numpy(np.stack(frames)).pin_memory()
    bh = torch.from_numpy(np.stack(bgs)).pin_memory()
    random.seed(9)
    ref_imgs, ref_labels, _ = OC.run_chain(frames, bgs, **kw)
    random.seed(9)
    res = chain.run_chain_host(fh, bh, chain.ChainParams(**kw), DEV)
    assert res.labels == ref_labels and all(np.array_equal(a, b) for a, b in zip(res.images.numpy(), ref_imgs))
    random.seed(9)
    ref_imgs, ref_labels, ref_frames = [], [], []
    for lo in range(0, 6, 4):
        i, l, r = OC.run_chain(frames[lo:lo + 4], bgs[lo:lo + 4], **kw)
        ref_imgs += i; ref_labels += l; ref_frames += [x["frame"] + lo for x in r]
    random.seed(9)
    for windows in (True, False):
        random.seed(9)
        res = chain.run_chain_host(fh, bh, chain.ChainParams(**kw), DEV, chunk=4, background_windows=windows)
        assert res.labels == ref_labels and [r["frame"] for r in res.records] == ref_frames
        assert all(np.array_equal(a, b) for a, b in zip(res.images.numpy(), ref_imgs))
        whole = fh.numel() + bh.numel()
        assert res.h2d_bytes == (fh.numel() + len(ref_imgs) * 240 * 240 * 3 if windows else whole)


@pytest.mark.parametrize("bg_shape", [(240, 300), (300, 220), (256, 256)])
@pytest.mark.parametrize("run_kw", [dict(), dict(ambiguity_eps=0.2)])
def test_host_backgrounds_upload_only_crop_windows(bg_shape, run_kw):
    """Backgrounds left in pinned host memory: only each item's square-crop window is copied to the
    device (ipp_chain_upload_windows) — same images / labels / generator state as device backgrounds,
    and fewer bytes than the whole images."""
    H, W = 200, 320
    frames = [synth.frame(61, i, H, W) for i in range(5)]
    bgs = [synth.background(61, i, *bg_shape) for i in range(5)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    fd = torch.from_numpy(np.stack(frames)).to(DEV)
    bh = torch.from_numpy(np.stack(bgs)).pin_memory()
    random.seed(5)
    ref = chain.run_chain(fd, bh.to(DEV), chain.ChainParams(**kw), **run_kw)
    state = random.getstate()
    random.seed(5)
    res = chain.run_chain(fd, bh, chain.ChainParams(**kw), **run_kw)
    assert random.getstate() == state and res.labels == ref.labels
    assert torch.equal(res.images, ref.images)
    cs = min(bg_shape)
    assert res.h2d_background_bytes == res.images.shape[0] * cs * cs * 3
    # many items per background: whole backgrounds are fewer bytes than one window per item
    kw2 = dict(kw, num_rotations=3, choose_random=2)
    random.seed(6)
    ref = chain.run_chain(fd, bh[:1].to(DEV), chain.ChainParams(**kw2))
    random.seed(6)
    res = chain.run_chain(fd, bh[:1].contiguous().pin_memory(), chain.ChainParams(**kw2))
    assert res.labels == ref.labels and torch.equal(res.images, ref.images)
    assert res.h2d_background_bytes == bh[:1].numel()
    with pytest.raises(RuntimeError):
        chain.run_chain(fd, torch.from_numpy(np.stack(bgs)), chain.ChainParams(**kw))      # pageable host memory


@pytest.mark.parametrize("mode", ["submit_finish", "phases", "map"])
@pytest.mark.parametrize("run_kw", [dict(), dict(ambiguity_eps=0.2)])
def test_chain_stream_pipelining_is_transparent(mode, run_kw):
    """ChainStream (the four phases of consecutive batches interleaved on one stream, small read-backs through
    pinned memory + events, table rows confirmed before stage B2 reads them) gives the same images, labels and
    generator state as plain run_chain calls — also when rows ARE flagged and re-derived."""
    H, W = 200, 320
    NB = 6
    batches = [([synth.frame(50 + b, i, H, W) for i in range(3)], [synth.background(50 + b, i, 240, 300) for i in range(2)]) for b in range(NB)]
    kw = dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    dev_batches = [(torch.from_numpy(np.stack(f)).to(DEV), torch.from_numpy(np.stack(b)).to(DEV)) for f, b in batches]
    torch.cuda.synchronize()
    random.seed(77)
    plain = [chain.run_chain(f, b, chain.ChainParams(**kw), **run_kw) for f, b in dev_batches]
    flagged = [r.tables_flagged_rows for r in plain]
    plain = [(r.images.cpu().numpy(), r.labels) for r in plain]
    state = random.getstate()
    random.seed(77)
    pipe = chain.ChainStream(chain.ChainParams(**kw), **run_kw)
    if mode == "map":
        results = list(pipe.map(iter(dev_batches)))
    elif mode == "submit_finish":                 # one batch ahead; finish() runs the missing phases itself
        ticket = pipe.submit(*dev_batches[0])
        results = []
        for k in range(NB):
            nxt = pipe.submit(*dev_batches[k + 1]) if k + 1 < NB else None
            results.append(pipe.finish(ticket))
            ticket = nxt
    else:                                         # hand-written interleaving, draws in batch order
        results, q = [], []
        for k in range(NB + 2):
            if len(q) == 3 or (k >= NB and q):
                results.append(pipe.finish(q.pop(0)))
            if len(q) >= 2:
                pipe.prepare(q[-2])
            if q:
                pipe.plan(q[-1])
            if k < NB:
                q.append(pipe.submit(*dev_batches[k]))
        while q:
            results.append(pipe.finish(q.pop(0)))
    assert random.getstate() == state
    assert [r.tables_flagged_rows for r in results] == flagged
    if run_kw:
        assert sum(flagged) > 0
    assert len(results) == NB
    for (a, la), r in zip(plain, results):
        assert la == r.labels and np.array_equal(a, r.images.cpu().numpy())


@pytest.mark.parametrize("sizes", [((37, 53), (37, 20)),        # width unchanged: identity band in the horizontal pass
                                   ((64, 40), (23, 40)),        # height unchanged: identity band in the vertical pass
                                   ((40, 30), (95, 77)),        # up-scaling (3-tap support, scale < 1)
                                   ((700, 520), (41, 33)),      # 17x down: several 64-wide super-blocks per tile
                                   ((301, 7), (9, 5)), ((5, 300), (4, 11)), ((1, 1), (3, 2)), ((129, 130), (128, 129))])
def test_tensor_core_lanczos_passes_vs_pillow(sizes, monkeypatch):
    """The two LANCZOS kernels alone (stage B2 in its two-kernel mode, without the gather): planar premultiplied
    sprites written by the test, device tables + exact fix-up, result against PIL.Image.resize(LANCZOS) — bit-exact,
    for identity passes, up-scaling and extreme down-scaling (paths the chain's own size distribution rarely reaches)."""
    import ctypes as C
    monkeypatch.setenv("IPP_B2_UNFUSED", "1")
    from PIL import Image
    from image_processor_pipeline_b200 import _native, hostmath
    lib = _native.load()
    (sw, sh), (nw, nh) = sizes
    r = np.random.default_rng(sw * 1000 + nw)
    n = 3
    rgba = r.integers(0, 256, (n, sh, sw, 4), dtype=np.uint8)
    rgba[..., 3] = np.where(r.random((n, sh, sw)) < 0.7, 255, 0)          # alpha as the chain has it: 0 or 255
    rgba[rgba[..., 3] == 0] = 0
    items = np.zeros(n, dtype=chain.ITEM_DTYPE)
    items["spr_w"], items["spr_h"], items["new_w"], items["new_h"] = sw, sh, nw, nh
    tot = chain._layout(items)
    ws = torch.zeros(tot["ws"], dtype=torch.uint8, device=DEV)
    Wp = (sw + 15) & ~15
    for i in range(n):                                                     # [4][sh][Wp] byte planes at rot_off
        planes = np.zeros((4, sh, Wp), np.uint8)
        planes[:, :, :sw] = np.moveaxis(rgba[i], 2, 0)
        off = int(items[i]["rot_off"])
        ws[off:off + planes.size] = torch.from_numpy(planes.reshape(-1)).to(DEV)
    geom = chain._geom_struct(64, 64, (0, 0, 64, 64), 64, 64, hostmath.pack_hsv_ranges(synth.REFERENCE_FILTERS))
    items_d = torch.from_numpy(items.view(np.uint8)).to(DEV)
    tab_d = torch.zeros(max(1, tot["table_ints"]), dtype=torch.int32, device=DEV)
    amb_d = torch.zeros(chain.amb_ints(), dtype=torch.int32, device=DEV)
    band_d = torch.zeros(max(tot["band_bytes"], 256), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    if tot["max_rows"]:
        _native.check(lib.ipp_chain_tables(items_d.data_ptr(), n, tot["max_rows"], tab_d.data_ptr(), amb_d.data_ptr(),
                                           chain.AMB_CAPACITY, chain.AMBIGUITY_EPS, st), "tables")
        host = chain._host_tables(items, lib, tot["table_ints"])
        assert np.array_equal(tab_d.cpu().numpy()[:len(host)], host) or int(amb_d[0]) > 0
        tab_d[:len(host)] = torch.from_numpy(host).to(DEV)                 # exact tables: this test is about the passes
    dummy = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)           # frames / mask workspace are not touched (no gather)
    _native.check(lib.ipp_chain_stage_b2(dummy.data_ptr(), 64 * 3, 64 * 64 * 3, C.byref(geom), items_d.data_ptr(), n, 0, tot["ht"], tot["vt"],
                                         tab_d.data_ptr(), band_d.data_ptr(), tot["max_band_warps"], 0, 0, dummy.data_ptr(), ws.data_ptr(), st), "b2")
    out = ws.cpu().numpy()
    for i in range(n):
        off = int(items[i]["spr_off"])
        mine = out[off:off + nw * nh * 4].reshape(nh, nw, 4)
        ref = np.asarray(Image.fromarray(rgba[i], "RGBA").resize((nw, nh), Image.Resampling.LANCZOS))
        assert np.array_equal(mine, ref), (i, int((mine != ref).sum()))


@pytest.mark.parametrize("flips", [(0, 0), (1, 1)])
@pytest.mark.parametrize("sizes", [((37, 53), (37, 20)), ((64, 40), (23, 40)), ((40, 30), (95, 77)), ((700, 520), (41, 33)),
                                   ((301, 7), (9, 5)), ((5, 300), (4, 11)), ((1, 1), (3, 2)), ((129, 130), (128, 129)),
                                   ((783, 756), (495, 468)), ((457, 451), (191, 185)), ((95, 49), (500, 17))])
def test_fused_gather_and_staged_lanczos_vs_pillow(sizes, flips):
    """Stage B2 in its shared-memory mode (fused gather + horizontal pass, CTA-staged vertical pass, 32-position
    fragment blocks): frames and mask bits written by the test, identity map (+ both flips), result against
    PIL.Image.resize(LANCZOS) of the (flipped) sprite — bit-exact, over identity passes, up-scaling, 17x down-scaling
    and the largest / smallest sprites of the bench workload."""
    import ctypes as C
    from PIL import Image
    from image_processor_pipeline_b200 import _native, hostmath
    lib = _native.load()
    (sw, sh), (nw, nh) = sizes
    fh, fv = flips
    r = np.random.default_rng(sw * 1000 + nw + 7 * fh)
    n = 3
    rgba = r.integers(0, 256, (n, sh, sw, 4), dtype=np.uint8)
    alpha = r.random((n, sh, sw)) < 0.7
    rgba[..., 3] = np.where(alpha, 255, 0)
    frames = np.ascontiguousarray(rgba[..., [2, 1, 0]])                     # frames are BGR; pixels under a clear mask bit are ignored
    items = np.zeros(n, dtype=chain.ITEM_DTYPE)
    items["frame"] = np.arange(n)
    items["kind"] = chain.KIND["copy"]
    items["cc_w"], items["cc_h"], items["spr_w"], items["spr_h"] = sw, sh, sw, sh
    items["flip_h"], items["flip_v"] = fh, fv
    items["new_w"], items["new_h"] = nw, nh
    tot = chain._layout(items)
    assert tot["fused_smem"] > 0 and tot["vstaged_smem"] > 0 and tot["ht"] == 0
    wpr = (sw + 31) // 32
    bits = np.zeros((n, sh, wpr * 32), dtype=np.uint8)
    bits[:, :, :sw] = alpha
    words = np.packbits(bits.reshape(n, sh, wpr, 32), axis=-1, bitorder="little").view("<u4").reshape(n, sh, wpr)
    bits_d = torch.from_numpy(np.ascontiguousarray(words).view(np.uint8).reshape(-1)).to(DEV)
    frames_d = torch.from_numpy(frames.reshape(-1)).to(DEV)                  # exact length: see the access guarantee in ippgpu.h
    ws = torch.zeros(tot["ws"], dtype=torch.uint8, device=DEV)
    geom = chain._geom_struct(sw, sh, (0, 0, sw, sh), max(nw, 8), max(nh, 8), hostmath.pack_hsv_ranges(synth.REFERENCE_FILTERS))
    items_d = torch.from_numpy(items.view(np.uint8)).to(DEV)
    tab_d = torch.zeros(max(1, tot["table_ints"]), dtype=torch.int32, device=DEV)
    band_d = torch.zeros(max(tot["band_bytes"], 256), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    if tot["max_rows"]:
        host = chain._host_tables(items, lib, tot["table_ints"])
        tab_d[:len(host)] = torch.from_numpy(host).to(DEV)                 # exact tables: this test is about the passes
    _native.check(lib.ipp_chain_stage_b2(frames_d.data_ptr(), sw * 3, sh * sw * 3, C.byref(geom), items_d.data_ptr(), n, tot["gt"], 0, tot["vt"],
                                         tab_d.data_ptr(), band_d.data_ptr(), tot["max_band_warps"], tot["fused_smem"], tot["vstaged_smem"],
                                         bits_d.data_ptr(), ws.data_ptr(), st), "b2")
    out = ws.cpu().numpy()
    for i in range(n):
        off = int(items[i]["spr_off"])
        mine = out[off:off + nw * nh * 4].reshape(nh, nw, 4)
        src = rgba[i].copy()
        src[src[..., 3] == 0] = 0
        if fh:
            src = src[:, ::-1]
        if fv:
            src = src[::-1]
        ref = np.asarray(Image.fromarray(np.ascontiguousarray(src), "RGBA").resize((nw, nh), Image.Resampling.LANCZOS))
        assert np.array_equal(mine, ref), (i, int((mine != ref).sum()))

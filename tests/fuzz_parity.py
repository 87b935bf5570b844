"""Randomised parity sweep (not collected by pytest: run it by hand on a GPU box):

    python tests/fuzz_parity.py [cases] [seed] [--quiet]  > profiles/r02/fuzz_r02.txt

Random frame / background sizes, border margins, filter lists (with zones / GIMP scale), rotation counts and angle ranges,
symmetry pools, scale ranges, the enhance step on or off — the fused chain against the oracle chain (images bit-exact, labels
string-equal, generator state equal), plus random batches through the batched operators (rotation nearest, overlay, key + keep
largest, enhance) against the reference's cv2 / Pillow calls.  Prints one line per case (unless --quiet) and a summary; goes on after a mismatch (up to ten) and
exits non-zero if there was one."""
import random
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from image_processor_pipeline_b200 import batch, chain, synth  # noqa: E402
from oracle import chain as OC, libcalls as L  # noqa: E402

DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def chain_case(r: np.random.Generator, seed: int):
    H, W = int(r.integers(40, 320)), int(r.integers(48, 420))
    BH, BW = int(r.integers(64, 360)), int(r.integers(64, 360))
    n, m = int(r.integers(1, 5)), int(r.integers(1, 4))
    kind = r.integers(0, 3)
    if kind == 0:
        frames = [synth.frame(seed, i, H, W, speck_density=float(r.uniform(0.0, 0.05))) for i in range(n)]
    elif kind == 1:
        frames = [synth.noise(seed, i, (H, W, 3)) for i in range(n)]
    else:
        frames = [np.where(r.random((H, W, 1)) < 0.5, synth.frame(seed, i, H, W), synth.noise(seed, i, (H, W, 3))).astype(np.uint8) for i in range(n)]
    bgs = [synth.background(seed, i, BH, BW) for i in range(m)]
    filters = list(synth.REFERENCE_FILTERS)
    if r.random() < 0.3:
        filters = filters[:int(r.integers(1, 5))]
    kw = dict(color_ranges_to_exclude_hsv=filters)
    if r.random() < 0.3:
        kw["zones"] = [None if r.random() < 0.5 else tuple(int(v) for v in r.integers(0, min(H, W) // 4, 4)) for _ in filters]
    if r.random() < 0.4:
        kw["crop_margins"] = (float(r.uniform(0, 0.1)), int(r.integers(0, 6)), float(r.uniform(0, 0.1)), int(r.integers(0, 6)))
    pool = [["o", "h", "v", "hv"], ["h", "v"], ["hv"], ["o", "h", "h", "v"], ["v", "hv", "o"]][int(r.integers(0, 5))]
    kw.update(num_rotations=int(r.integers(0, 4)), include_original=bool(r.random() < 0.5), pool=pool,
              choose_random=int(r.integers(0, len(pool) + 1)), sym_include_original=bool(r.random() < 0.5),
              yolo_class_id=int(r.integers(0, 80)), enhance=bool(r.random() < 0.35))
    if r.random() < 0.5:
        a = float(r.uniform(0, 300))
        kw.update(angle_min=a, angle_max=a + float(r.uniform(0, 59)))
    if r.random() < 0.5:
        s = float(r.uniform(0.05, 0.5))
        kw.update(scale_min=s, scale_max=s + float(r.uniform(0, 0.4)))
    random.seed(seed)
    ref_imgs, ref_labels, _ = OC.run_chain(frames, bgs, **kw)
    state = random.getstate()
    random.seed(seed)
    res = chain.run_chain(dev(np.stack(frames)), dev(np.stack(bgs)), chain.ChainParams(**kw))
    assert random.getstate() == state, "generator state differs"
    assert res.labels == ref_labels, "labels differ"
    out = res.images.cpu().numpy()
    assert len(ref_imgs) == out.shape[0]
    for a, b in zip(out, ref_imgs):
        assert np.array_equal(a, b), "image differs"
    return f"chain {H}x{W} bg {BH}x{BW} n={n} kind={kind} rot={kw['num_rotations']} pool={pool} enh={kw['enhance']} -> {out.shape[0]} images"


def ops_case(r: np.random.Generator, seed: int):
    which = int(r.integers(0, 4))
    if which == 0:
        n, H, W, c = int(r.integers(1, 6)), int(r.integers(8, 200)), int(r.integers(8, 260)), int(r.choice([3, 4]))
        imgs = r.integers(0, 256, (n, H, W, c), dtype=np.uint8)
        if c == 4:
            imgs[..., 3] = np.where(r.random((n, H, W)) < 0.4, 0, imgs[..., 3])
        random.seed(seed)
        res = batch.rotate_batch(dev(imgs), random, 0.0, 360.0)
        random.seed(seed)
        for i in range(n):
            ang = random.uniform(0.0, 360.0)
            assert np.array_equal(res.canvas(i).cpu().numpy(), L.rotate_expand(L.to_rgba(imgs[i]), ang)), "rotation differs"
        return f"rotate_batch n={n} {H}x{W}x{c}"
    if which == 1:
        S, sh, sw = int(r.integers(1, 4)), int(r.integers(6, 120)), int(r.integers(6, 160))
        M, BH, BW, n = int(r.integers(1, 4)), int(r.integers(32, 300)), int(r.integers(32, 300)), int(r.integers(1, 8))
        spr = r.integers(0, 256, (S, sh, sw, 4), dtype=np.uint8)
        bgs = r.integers(0, 256, (M, BH, BW, 3), dtype=np.uint8)
        random.seed(seed)
        res = batch.overlay_batch(dev(spr), dev(bgs), n, random)
        random.seed(seed)
        k = 0
        for p in range(n):
            try:
                comp, label, _ = L.overlay_onto_background(spr[p % S], bgs[p % M])
            except ValueError:
                continue
            assert res.labels[k] == label and np.array_equal(res.images[k].cpu().numpy(), comp), "overlay differs"
            k += 1
        assert k == len(res.labels)
        return f"overlay_batch S={S} {sh}x{sw} bg {BH}x{BW} pairs={n} -> {k}"
    if which == 2:
        n, H, W = int(r.integers(1, 6)), int(r.integers(8, 260)), int(r.integers(8, 300))
        frames = np.stack([synth.frame(seed, i, H, W, speck_density=float(r.uniform(0, 0.08))) if r.random() < 0.7 else synth.noise(seed, i, (H, W, 3)) for i in range(n)])
        out, info = batch.key_and_keep_largest(dev(frames), synth.REFERENCE_FILTERS)
        out = out.cpu().numpy()
        for i in range(n):
            keyed = L.color_masks_to_alpha(frames[i], list(synth.REFERENCE_FILTERS))
            x, y, w, h = (int(v) for v in info[i, :4])
            if not (keyed[..., 3] > 1).any():
                assert w == 0
                continue
            assert np.array_equal(out[i, y:y + h, x:x + w], L.keep_largest_component(keyed)), "keep-largest differs"
        return f"key_and_keep_largest n={n} {H}x{W}"
    n, H, W = int(r.integers(1, 7)), int(r.integers(4, 220)), int(r.integers(4, 260))
    imgs = r.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    random.seed(seed)
    flags = batch.sample_flags(n, random)
    if n > 1:
        flags[0][0] = flags[1][1] = True
    out = batch.enhance_batch(dev(imgs), flags[0], flags[1], random).cpu().numpy()
    random.seed(seed)
    L.sample_pairing_flags(n)
    for i in range(n):
        assert np.array_equal(out[i], L.enhance_image(imgs[i], bool(flags[0][i]), bool(flags[1][i]))), "enhance differs"
    return f"enhance_batch n={n} {H}x{W}"


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    quiet = "--quiet" in sys.argv
    t0 = time.time()
    done = {"chain": 0, "ops": 0}
    bad = []
    for c in range(cases):
        seed = seed0 * 100_000 + c
        r = np.random.default_rng(seed)
        fn, key = (chain_case, "chain") if c % 2 == 0 else (ops_case, "ops")
        try:
            line = fn(r, seed)
        except AssertionError as e:
            print(f"case {c} seed {seed}: MISMATCH: {e}", flush=True)
            bad.append(seed)
            if len(bad) >= 10:
                break
            continue
        done[key] += 1
        if not quiet:
            print(f"case {c:4d} ok  {line}")
    print(f"SUMMARY: {done['chain']} random chain cases and {done['ops']} random batched-operator cases bit-exact against the oracle in {time.time() - t0:.0f} s"
          + (f"; MISMATCHES at seeds {bad}" if bad else ""))
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()

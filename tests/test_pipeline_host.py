"""The orchestrator contract (reference pipeline.py) with plain host process functions. CPU only."""
import json
import random
from pathlib import Path

import pytest

from image_processor_pipeline_b200.pipeline import MODES, PathJSONEncoder, ProcessingPipeline, ProcessingStep
from image_processor_pipeline_b200.utils.utils import _validate_dirs, check_path


def _touch(d: Path, names):
    d.mkdir(parents=True, exist_ok=True)
    for n in names:
        (d / n).write_text(n)


def copy_fn(path: Path, output_dirs, suffix="", **options):
    out = output_dirs[0] / f"{path.stem}{suffix}{path.suffix}"
    out.write_text(path.read_text())
    return out


def pair_fn(a: Path, b: Path, output_dirs, **options):
    out = output_dirs[0] / f"{a.stem}+{b.stem}.txt"
    out.write_text("x")
    return [out]


def flaky_fn(path: Path, output_dirs, **options):
    if path.stem == "b":
        raise ValueError("boom")
    if path.stem == "c":
        return None
    if path.stem == "d":
        return "not a path"
    return copy_fn(path, output_dirs)


def sample_fn(path: Path, do_blur: bool, do_rgb: bool, output_dirs, **options):
    out = output_dirs[0] / f"{path.stem}_{int(do_blur)}{int(do_rgb)}.txt"
    out.write_text("x")
    return out


def test_validate_dirs_and_check_path(tmp_path):
    assert _validate_dirs([tmp_path], 1) == tmp_path
    assert _validate_dirs([tmp_path, "b", "c"], 2) == (tmp_path, Path("b"), Path("c"))
    with pytest.raises(IndexError):
        _validate_dirs([tmp_path], 2)
    assert check_path("x", tmp_path) == tmp_path / "x" and check_path(tmp_path / "y", "/zzz") == tmp_path / "y"


def test_step_validation():
    with pytest.raises(ValueError):
        ProcessingStep("s", copy_fn, input_dirs="a")                       # no output_dirs
    with pytest.raises(ValueError):
        ProcessingStep("s", copy_fn, "a", "b", pairing_method="nope")
    with pytest.raises(ValueError):
        ProcessingStep("s", copy_fn, "a", "b", pairing_method="custom")
    assert MODES == ("one_input", "zip", "modulo", "sample", "custom")
    s = ProcessingStep("s", copy_fn, "a", ["b", "c"], root_dir="/r", workers=None)
    assert s.input_paths == [Path("/r/a")] and s.output_paths == [Path("/r/b"), Path("/r/c")] and s.parallels_workers == 1


def test_one_input_statuses_and_json_log(tmp_path):
    _touch(tmp_path / "in", ["a.txt", "b.txt", "c.txt", "d.txt", "e.txt"])
    step = ProcessingStep("Étape 1", flaky_fn, tmp_path / "in", tmp_path / "out" / "imgs", save_log=True)
    with pytest.warns(UserWarning):
        step.run()
    st = {Path(l["inputs"][0]).stem: l["status"] for l in step.process_logs}
    assert st == {"a": "Success", "b": "Error", "c": "no_output", "d": "Type Error", "e": "Success"}
    log = json.loads((tmp_path / "out" / "Étape 1.json").read_text())
    assert len(log) == 5 and log[0]["outputs"] == [str(tmp_path / "out" / "imgs" / "a.txt")]
    assert json.dumps({"p": Path("x"), "t": (1, 2)}, cls=PathJSONEncoder) == '{"p": "x", "t": [1, 2]}'


def test_pairing_modes_and_their_draws(tmp_path):
    _touch(tmp_path / "a", [f"a{i}.txt" for i in range(5)])
    _touch(tmp_path / "b", [f"b{i}.txt" for i in range(2)])
    z = ProcessingStep("zip", pair_fn, [tmp_path / "a", tmp_path / "b"], tmp_path / "oz", pairing_method="zip")
    z.run()
    assert sorted(p.name for p in (tmp_path / "oz").iterdir()) == ["a0+b0.txt", "a1+b1.txt"]

    random.seed(4)
    m = ProcessingStep("mod", pair_fn, [tmp_path / "a", tmp_path / "b"], tmp_path / "om", pairing_method="modulo")
    m.run()
    random.seed(4)
    second = sorted((tmp_path / "b").iterdir())
    random.shuffle(second)
    expect = sorted(f"a{i}+{second[i % 2].stem}.txt" for i in range(5))
    assert sorted(p.name for p in (tmp_path / "om").iterdir()) == expect

    random.seed(5)
    s = ProcessingStep("smp", sample_fn, tmp_path / "a", tmp_path / "os", pairing_method="sample")
    s.run()
    random.seed(5)
    files = sorted((tmp_path / "a").iterdir())
    blur = set(random.sample(files, 1)); rgb = set(random.sample(files, 1))
    assert sorted(p.name for p in (tmp_path / "os").iterdir()) == sorted(f"{f.stem}_{int(f in blur)}{int(f in rgb)}.txt" for f in files)

    random.seed(6)
    k = ProcessingStep("k", copy_fn, tmp_path / "a", tmp_path / "ok", sample_k=2)
    k.run()
    random.seed(6)
    ids = random.sample(range(5), 2)
    assert sorted(p.name for p in (tmp_path / "ok").iterdir()) == sorted(f"a{i}.txt" for i in ids)

    c = ProcessingStep("cus", pair_fn, [tmp_path / "a", tmp_path / "b"], tmp_path / "oc", pairing_method="custom",
                       pairing_function=lambda lists: ((lists[0][0], b) for b in lists[1]))
    c.run()
    assert sorted(p.name for p in (tmp_path / "oc").iterdir()) == ["a0+b0.txt", "a0+b1.txt"]


def test_pipeline_chaining_insert_and_resume(tmp_path):
    _touch(tmp_path / "raw", ["x.txt", "y.txt"])
    pipe = ProcessingPipeline(root_dir=tmp_path)
    with pytest.raises(ValueError):
        pipe.add_step(ProcessingStep("first", copy_fn, output_dirs="o1"))
    s1 = ProcessingStep("one", copy_fn, "raw", "o1", options={"suffix": "_1"})
    s3 = ProcessingStep("three", copy_fn, output_dirs="o3", options={"suffix": "_3"})
    pipe.add_step(s1)
    pipe.add_step(s3)
    assert s1.input_paths == [tmp_path / "raw"] and s3.input_paths == [tmp_path / "o1"]
    s2 = ProcessingStep("two", copy_fn, output_dirs="o2", options={"suffix": "_2"})
    pipe.add_step(s2, position=1)
    assert s2.input_paths == [tmp_path / "o1"] and s3.input_paths == [tmp_path / "o2"]
    with pytest.raises(IndexError):
        pipe.run(from_step_index=3)
    pipe.run()
    assert sorted(p.name for p in (tmp_path / "o3").iterdir()) == ["x_1_2_3.txt", "y_1_2_3.txt"]
    (tmp_path / "o3" / "x_1_2_3.txt").unlink()
    pipe.run(from_step_index=2, only_one=True)
    assert (tmp_path / "o3" / "x_1_2_3.txt").exists()


def test_parallel_workers_process_pool_for_host_functions(tmp_path):
    _touch(tmp_path / "in", [f"f{i}.txt" for i in range(6)])
    step = ProcessingStep("par", flaky_fn, tmp_path / "in", tmp_path / "out", workers=2)
    step.run()
    assert sorted(l["status"] for l in step.process_logs) == ["Success"] * 6


def test_copy_helpers_follow_the_reference(tmp_path):
    """transforms/copy.py:6-67 — pure file plumbing, kept for drop-in imports."""
    from image_processor_pipeline_b200.transforms import copy as C
    src = tmp_path / "src"; a = tmp_path / "a"; b = tmp_path / "b"
    for d in (src, a, b):
        d.mkdir()
    (src / "img_0001.png").write_bytes(b"png")
    (src / "img_0001.txt").write_text("0 0.5 0.5 0.1 0.1\n")
    out = C.copy_img_with_labels(src / "img_0001.png", src / "img_0001.txt", [a, b])
    assert out == [a / "img_0001.png", b / "img_0001.txt"] and (b / "img_0001.txt").read_text().startswith("0 0.5")
    with pytest.raises(ValueError):
        C.copy_img_with_labels(src / "img_0001.png", src / "img_0001.txt", [a])
    assert C.copy_img_with_labels(src / "absent.png", src / "img_0001.txt", [a, b]) is None
    assert C.copy_files(src / "img_0001.png", [a], suffix="v2") == a / "img_0001_v2.png"
    assert C.copy_files(src / "img_0001.png", [a], replace_params=("img", "pic")) == a / "pic_0001.png"
    assert C.copy_files(src / "img_0001.txt", [b]) == b / "img_0001.txt"
    with pytest.raises(ValueError):
        C.copy_files(src / "img_0001.png", [a], suffix="x", replace_params=("a", "b"))
    with pytest.raises(IndexError):
        C.copy_files(src / "img_0001.png", [])

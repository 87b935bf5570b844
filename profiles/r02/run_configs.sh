# every BASELINE config through the driver-style bench (1 GPU): b200 arm + reference arm
set -x
mkdir -p gpurun_out/cfg
for c in c1 c2 c3 c4 c5 chain; do
  python bench.py --config $c > gpurun_out/cfg/bench_$c.json 2> gpurun_out/cfg/bench_$c.err || tail -5 gpurun_out/cfg/bench_$c.err
  python bench.py --config $c --impl reference --steps 3 --warmup 1 > gpurun_out/cfg/ref_$c.json 2> gpurun_out/cfg/ref_$c.err || tail -5 gpurun_out/cfg/ref_$c.err
  python - <<PY
import json
for kind in ("bench", "ref"):
    try:
        d = json.load(open("gpurun_out/cfg/%s_$c.json" % kind))
        print(kind, "$c", round(d["value"], 1), "img/s", "ms/step", round(d["ms_per_step"], 3), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1),
              "roofline", d.get("roofline") and (d["roofline"]["kernel"], round(d["roofline"]["frac"], 3)), "frac_of_peak", d.get("frac_of_peak") and round(d["frac_of_peak"], 3),
              {k: v for k, v in d.items() if k.startswith("value_")}, "cpu1", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1))
        if kind == "bench":
            print("   kernels", {k: v["avg_ms"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(kind, "$c", "FAILED", e)
PY
done

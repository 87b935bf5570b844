for v in "" seg512 seg256; do
  for st in caller side; do
    if [ -n "$v" ]; then export IPPGPU_LIB=$PWD/image_processor_pipeline_b200/_variants/libippgpu_$v.so; else unset IPPGPU_LIB; fi
    python bench.py --no-e2e --steps 12 --warmup 4 --stage-a-stream $st > gpurun_out/var_${v:-base}_$st.json 2> gpurun_out/var_${v:-base}_$st.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${v:-base}_$st.json"))
    print("${v:-base}", "$st", round(d["value"]), round(d["ms_per_step"],3), {k:v["avg_ms"] for k,v in d["kernels"].items() if v["avg_ms"]>0.2})
except Exception as e:
    print("${v:-base}", "$st", "FAILED", e)
PY
  done
done

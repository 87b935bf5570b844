# multi-GPU lines of one box: PCIe ceiling of the end-to-end mix, the headline chain and BASELINE configs[4] (4K stream)
N=$1
mkdir -p gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then TR="python"; fi
$TR profiles/probes/pcie_probe_multi.py > gpurun_out/multi/pcie_n$N.json 2> gpurun_out/multi/pcie_n$N.err; cat gpurun_out/multi/pcie_n$N.json
$TR bench.py --gpus $N --steps 40 --warmup 5 --no-noise > gpurun_out/multi/chain_n$N.json 2> gpurun_out/multi/chain_n$N.err
$TR bench.py --gpus $N --config c5 --steps 30 --warmup 5 > gpurun_out/multi/c5_n$N.json 2> gpurun_out/multi/c5_n$N.err
python - <<PY
import json
for c in ("chain", "c5"):
    try:
        d = json.load(open("gpurun_out/multi/%s_n$N.json" % c))
        print(c, "N=$N", round(d["value"]), "img/s", "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(c, "N=$N FAILED", e)
PY

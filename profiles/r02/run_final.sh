# final lines of round 2 after the window gather / enhance changes (one GPU)
mkdir -p gpurun_out/fin
python bench.py > gpurun_out/fin/bench_chain.json 2> gpurun_out/fin/bench_chain.err
python bench.py --impl reference > gpurun_out/fin/bench_ref_chain.json 2> gpurun_out/fin/bench_ref_chain.err
python bench.py --enhance --no-noise > gpurun_out/fin/bench_chain_enhance.json 2> gpurun_out/fin/bench_chain_enhance.err
B="python bench.py --no-e2e --no-noise"
$B --enhance --batch 64 --steps 1 --warmup 1 > gpurun_out/fin/plain_enh.log 2>&1 && \
ncu --set full --clock-control none -k regex:'^enh_' -s 6 -c 6 -o gpurun_out/fin/prof_enh $B --enhance --batch 64 --steps 1 --warmup 1 > gpurun_out/fin/ncu_enh.log 2>&1
python profiles/summarize.py kernel gpurun_out/fin/prof_enh.ncu-rep > gpurun_out/fin/kernels_enh_r02.txt
rm -f gpurun_out/fin/*.ncu-rep
python tests/fuzz_parity.py 6000 91 --quiet > gpurun_out/fin/fuzz_r02.txt 2>&1; tail -1 gpurun_out/fin/fuzz_r02.txt

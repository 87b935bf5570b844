# ncu evidence of round 2 (one gpurun call; every ncu run directly follows the same command run plainly with &&)
set -x
B="python bench.py --no-e2e --no-noise"
$B --steps 4 --warmup 3 > gpurun_out/ncu_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(chain_|ccl_|hsv_|bbox_)' -c 400 --csv --log-file gpurun_out/launches_r02.csv $B --steps 4 --warmup 3 > gpurun_out/ncu1.log 2>&1
$B --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:'^(chain_|ccl_)' -s 15 -c 16 -o gpurun_out/prof_chain_r02 $B --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
$B --config c2 --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none -k regex:'rotate_batch' -s 1 -c 2 -o gpurun_out/prof_c2_r02 $B --config c2 --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
$B --config c4 --batch 512 --steps 1 --warmup 1 > gpurun_out/ncu_plain4.log 2>&1 && \
ncu --set full --clock-control none -k regex:'bgra_from_keep|chain_mask_seg' -s 2 -c 2 -o gpurun_out/prof_c4_r02 $B --config c4 --batch 512 --steps 1 --warmup 1 > gpurun_out/ncu4.log 2>&1
$B --config c1 --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu_plain5.log 2>&1 && \
ncu --set full --clock-control none -k regex:'rows_kernel' -s 5 -c 2 -o gpurun_out/prof_c1_r02 $B --config c1 --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu5.log 2>&1
$B --config c3 --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu_plain6.log 2>&1 && \
ncu --set full --clock-control none -k regex:'sprite_planes|chain_hpass|chain_vs|chain_paste' -s 4 -c 4 -o gpurun_out/prof_c3_r02 $B --config c3 --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu6.log 2>&1
$B --enhance --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_plain7.log 2>&1 && \
ncu --set full --clock-control none -k regex:'^enh_' -s 5 -c 5 -o gpurun_out/prof_enh_r02 $B --enhance --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu7.log 2>&1
# summaries are made on the box (the reports together exceed what gpurun brings back); the reports are then dropped
mkdir -p gpurun_out/ncu_r02
python profiles/summarize.py launches gpurun_out/launches_r02.csv > gpurun_out/ncu_r02/launches_r02.txt
for n in chain c2 c4 c1 c3 enh; do python profiles/summarize.py kernel gpurun_out/prof_${n}_r02.ncu-rep > gpurun_out/ncu_r02/kernels_${n}_r02.txt; done
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
head -20 gpurun_out/ncu_r02/launches_r02.txt

# rotate_batch_kernel tuning sweep (round 2): build variants with
#   python -m image_processor_pipeline_b200.build --variant NAME -DIPP_ROT_TILE_H=.. -DIPP_ROT_UNROLL=.. -DIPP_ROT_MINB=.. -DIPP_ROT_MINB_BIL=..
# then run this on a GPU box.  Results (frac = nearest kernel vs the copy peak; bilinear images/s):
#   tile 64x32 (round-1 shape) 0.46 | x64 0.45 | x128 0.51 | x256 0.53 | x512 0.52        (unroll 2, no register cap)
#   tile 64x256: unroll 1 / 2 / 4 with 8 CTAs per SM (32 registers) 0.60 / 0.60 / 0.60; 6 CTAs 0.57; 5 CTAs (46 registers) 0.53
#   bilinear, min CTAs 2 / 3 / 4 / 5 (86 / 80 / 64 / 48 registers): 39.1 / 45.3 / 47.0 / 45.0 k images/s
for l in image_processor_pipeline_b200/_variants/*.so; do IPPGPU_LIB=$l python bench.py --config c2 --no-e2e --cpu-sample 0 --steps 20 --warmup 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d.get('value_bilinear'))"; done

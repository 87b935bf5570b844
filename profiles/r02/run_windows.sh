# crop windows of host backgrounds: copy engines (cudaMemcpy2DAsync per item) vs the SM-driven gather from pinned memory
for k in chain c5; do for cfg in dma:32 sm:16 sm:32 sm:64; do m=${cfg%%:*}; c=${cfg##*:}
IPP_WINDOWS_COPY=$m IPP_WINDOWS_CTAS=$c python bench.py --config $k --no-noise --cpu-sample 0 --steps 20 --warmup 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$k $m', $c, 'e2e', round(d['e2e']['value'],1), 'resident', round(d['value']))"
done; done

# compute-sanitizer memcheck over the chain / batch parity tests, torch's caching allocator off (every tensor is its own
# cudaMalloc of the exact size: out-of-bounds accesses are seen as such)
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_chain.py tests/test_gpu_batch.py -m gpu -x -q \
  -k "odd_sizes or fused_gather or chain_small or rotate_batch_nearest or crop_batch or overlay_batch or key_and_keep" > gpurun_out/memcheck_r02.txt 2>&1
tail -15 gpurun_out/memcheck_r02.txt

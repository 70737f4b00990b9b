#!/bin/sh
# round 2, call w: quick check of a kernel change -- parity (kernel tests), then k_encode_chunks time on the default content and on noise
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.txt 2>&1 || { tail -8 gpurun_out/r2w_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2w_smoke.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2w_pytest_kernels.txt
{
for rep in 1 2; do
  echo "natural: $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1)  grey: $(timeout 120 python tools/time_kernel.py 300 2 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2w_times.txt

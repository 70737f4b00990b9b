#!/bin/sh
# round 2, call o: tail restructure (header words direct, long-block recode out of line, 32-bit record index, predicated scan)
# -- parity, kernel time, one ncu --set full capture with source
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.txt 2>&1 || { tail -8 gpurun_out/r2o_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2o_smoke.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2o_pytest_kernels.txt
{
for rep in 1 2; do
  echo "natural: $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2o_times.txt
B="python bench.py --steps 2 --warmup 3 --frames 40 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/r2o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/r2o_prof $B > gpurun_out/r2o_ncu_full.log 2>&1
tail -2 gpurun_out/r2o_ncu_full.log

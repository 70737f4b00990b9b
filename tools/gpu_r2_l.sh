#!/bin/sh
# round 2, call l: L2 cache-policy hints (pixels evict-first, chunk records evict-last) -- parity, then A/B of kernel times
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.txt 2>&1 || { tail -8 gpurun_out/r2l_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2l_smoke.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -4
{
for v in nohints hints nohints hints; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v: $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2l_l2_hints_ab.txt

#!/bin/sh
# A/B of prebuilt variants, then parity + phase split with the integer-luma build as the product library.
sh tools/gpu_variants.sh base new intluma tab_cpasync pf_next intluma_cp_pf 2>&1 | tee gpurun_out/variants2.log
for v in new intluma; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== phases $v"; sh tools/gpu_phases.sh
done 2>&1 | tee gpurun_out/phases2.log
cp build_variants/libm1cu_intluma.so ec504_imageencoder_b200/libm1cu.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_intluma.log

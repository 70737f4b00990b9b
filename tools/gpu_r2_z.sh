#!/bin/sh
# round 2, call z: the randomized geometry / quality / content test, then the whole GPU suite once more
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "random_geometries" 2>&1 | tail -6 | tee gpurun_out/r2z_random.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2z_pytest_1gpu.txt

#!/bin/sh
# round 2, call g: the warp-per-chunk kernel (k_encode_groups) -- parity first, then A/B against the CTA-per-chunk kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.txt 2>&1 || { tail -8 gpurun_out/r2g_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2g_smoke.txt
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "both_encode or kernel_variant" 2>&1 | tail -12 > gpurun_out/r2g_pytest_kernels.txt; cat gpurun_out/r2g_pytest_kernels.txt
grep -q failed gpurun_out/r2g_pytest_kernels.txt && { echo KERNEL_TESTS_FAILED; }
python tools/content_sweep.py r2g 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(r['kernel'][:16], r['content'], 'q', r['quality'], 'fps', round(r['frames_per_s']), 'enc_ms', round(r['encode_kernel_ms'], 3), 'frac', round(r['encode_kernel_roofline_frac'], 4))
" | tee gpurun_out/r2g_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2g_pytest.txt; cat gpurun_out/r2g_pytest.txt

#!/bin/sh
# round 2, call m: do the CTAs of one SM run their colour / block phases in lockstep?  Timeline trace + first-wave stagger A/B
mkdir -p gpurun_out
cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
{
echo "== trace (launch 6 of time_kernel.py, 300 frames)"
M1_TRACE=6 M1_TRACE_FILE=gpurun_out/r2m_trace.bin timeout 200 python tools/time_kernel.py 300 0 2>&1 | tail -1
python tools/trace_phases.py gpurun_out/r2m_trace.bin
for rep in 1 2; do
for ns in 0 600 1200 2400; do
  echo "== stagger $ns ns: $(M1_STAGGER_NS=$ns timeout 200 python tools/time_kernel.py 300 0 2>&1 | tail -1)"
done
done
echo "== trace with stagger 1200"
M1_STAGGER_NS=1200 M1_TRACE=6 M1_TRACE_FILE=gpurun_out/r2m_trace_st.bin timeout 200 python tools/time_kernel.py 300 0 2>&1 | tail -1
python tools/trace_phases.py gpurun_out/r2m_trace_st.bin
} 2>&1 | tee gpurun_out/r2m_lockstep.txt
rm -f gpurun_out/r2m_trace_st.bin

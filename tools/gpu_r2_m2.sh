cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
{
M1_TRACE=6 M1_TRACE_FILE=gpurun_out/r2m_trace.bin timeout 200 python tools/time_kernel.py 300 0 2>&1 | tail -1
python tools/trace_phases.py gpurun_out/r2m_trace.bin
M1_STAGGER_NS=1200 M1_TRACE=6 M1_TRACE_FILE=gpurun_out/r2m_trace_st.bin timeout 200 python tools/time_kernel.py 300 0 2>&1 | tail -1
python tools/trace_phases.py gpurun_out/r2m_trace_st.bin
} 2>&1 | tee gpurun_out/r2m_lockstep_trace.txt

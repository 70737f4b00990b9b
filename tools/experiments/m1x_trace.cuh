// m1x_trace.cuh -- PROFILING build only (nvcc -DM1_EXPERIMENTS, tools/build_experiments.sh exp): per-CTA phase
// timeline of k_encode_chunks and a first-wave start stagger.  Included by csrc/m1cu_kernels.cu under
// #ifdef M1_EXPERIMENTS; the product library never sees this file.
//   M1_TRACE=n         dump the timeline of the n-th encode launch to $M1_TRACE_FILE (default gpurun_out/trace.bin):
//                      per CTA four 64-bit words = (smid << 48 | start), end of colour phase, end of block phase, end
//                      (globaltimer ns, low 48 bits); read with tools/trace_phases.py
//   M1_STAGGER_NS=d    the k-th CTA to start on an SM waits k * d ns (k < M1_STAGGER_CTAS, default the resident count)
#pragma once
#include <vector>
#include <stdio.h>
#include "m1x_env.h"
__device__ unsigned long long *m1x_trace;       // [CTA][4]: (smid << 48 | start), end of colour, end of blocks, end   (globaltimer ns)
__device__ unsigned int m1x_sm_arrivals[1024];  // CTAs that have started on each SM in this launch
__device__ int m1x_stagger_ns, m1x_stagger_ctas;
__device__ __forceinline__ unsigned long long m1x_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned m1x_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#define M1X_MARK(slot) do { if (m1x_trace && threadIdx.x == 0) m1x_trace[4 * m1x_lin + (slot)] = (m1x_now() & 0xffffffffffffull) | ((slot) == 0 ? (unsigned long long)m1x_smid() << 48 : 0ull); } while (0)

static int m1x_launch_no = 0;
static unsigned long long *m1x_d_trace = nullptr;
static inline void m1x_before_launch(dim3 grid, cudaStream_t st)
{
    static const int stagger = m1x_env_int("M1_STAGGER_NS");
    static const int stagger_ctas = m1x_env_int("M1_STAGGER_CTAS") ? m1x_env_int("M1_STAGGER_CTAS") : 7;
    static const int trace_launch = m1x_env_int("M1_TRACE");
    ++m1x_launch_no;
    const size_t n_ctas = (size_t)grid.x * grid.y * grid.z;
    if (stagger) {
        void *arr; cudaGetSymbolAddress(&arr, m1x_sm_arrivals);
        cudaMemsetAsync(arr, 0, sizeof(unsigned int) * 1024, st);
        if (m1x_launch_no == 1) { cudaMemcpyToSymbol(m1x_stagger_ns, &stagger, sizeof(int)); cudaMemcpyToSymbol(m1x_stagger_ctas, &stagger_ctas, sizeof(int)); }
    }
    if (trace_launch && m1x_launch_no == trace_launch) {
        cudaMalloc(&m1x_d_trace, n_ctas * 32); cudaMemset(m1x_d_trace, 0, n_ctas * 32);
        cudaMemcpyToSymbol(m1x_trace, &m1x_d_trace, sizeof(m1x_d_trace));
    }
}
static inline void m1x_after_launch(dim3 grid, cudaStream_t st)
{
    static const int trace_launch = m1x_env_int("M1_TRACE");
    if (!(trace_launch && m1x_launch_no == trace_launch)) return;
    const size_t n_ctas = (size_t)grid.x * grid.y * grid.z;
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h(n_ctas * 4);
    cudaMemcpy(h.data(), m1x_d_trace, n_ctas * 32, cudaMemcpyDeviceToHost);
    unsigned long long *nul = nullptr; cudaMemcpyToSymbol(m1x_trace, &nul, sizeof(nul));
    const char *fn = getenv("M1_TRACE_FILE");
    if (FILE *f = fopen(fn ? fn : "gpurun_out/trace.bin", "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
}

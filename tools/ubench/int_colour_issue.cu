// int_colour_issue.cu -- micro-benchmark for the integer colour path:
// what do the instructions of m1cu_colour.cuh cost per sub-partition, alone and mixed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_colour_issue int_colour_issue.cu && ./int_colour_issue
// Same harness as fp64_issue.cu: CH independent chains per class and warp, 148 CTAs, 1..8 warps per sub-partition;
// the table gives cycles per GROUP (one instruction of each listed class).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8

enum { DPU = 1, DPS = 2, DP4 = 4, WID = 8, WIM = 16, MN3 = 32, IMA = 64, ALU = 128, MOV = 256, SHR = 512, PRM = 1024, I2P = 2048, FFM = 4096, LEA_ = 8192 };

template <int M>
__global__ void k(long long *cycles, int *sinki, int iseed)
{
    int a[CH], b[CH], c[CH], m[CH], l[CH], s[CH], p[CH], e[CH];
    unsigned n3[CH];
    unsigned long long w[CH], wi[CH];
    float f[CH], g[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = iseed + i * 3 + threadIdx.x; b[i] = iseed * 5 + i; c[i] = threadIdx.x + i; m[i] = iseed * 7 + i + threadIdx.x * 5;
        l[i] = iseed + i; s[i] = 0x7fffffff - threadIdx.x - i; p[i] = threadIdx.x * 3 + i; n3[i] = 0xffffffffu - i; e[i] = i;
        w[i] = i + threadIdx.x; wi[i] = i * 5 + threadIdx.x; f[i] = (float)(threadIdx.x + i); g[i] = 0.f;
    }
    const int x = iseed + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (M & DPU) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(it));
            if (M & DPS) asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(x), "r"(it));
            if (M & DP4) asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(x), "r"(it));
            if (M & WID) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(it));
            if (M & WIM) asm volatile("mul.wide.u32 %0, %1, 137439;" : "=l"(wi[i]) : "r"((unsigned)wi[i] + (unsigned)(wi[i] >> 32)));
            if (M & MN3) asm volatile("{\n\t.reg .u32 t;\n\tmin.u32 t, %1, %2;\n\tmin.u32 %0, %0, t;\n\t}" : "+r"(n3[i]) : "r"(x + i), "r"(it));
            if (M & IMA) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(m[i]) : "r"(x), "r"(it));
            if (M & ALU) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(x), "r"(it));
            if (M & SHR) asm volatile("shr.s32 %0, %0, 2;" : "+r"(s[i]));
            if (M & PRM) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(p[i]) : "r"(x));
            if (M & I2P) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(g[i]) : "r"(e[i] + __float_as_int(g[i])));
            if (M & FFM) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(f[i]) : "f"(g[0]));
            if (M & LEA_) asm volatile("{\n\t.reg .u32 t;\n\tshl.b32 t, %1, 3;\n\tadd.u32 %0, %0, t;\n\t}" : "+r"(e[i]) : "r"(x));
        }
    }
    const long long t1 = clock64();
    int si = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i)
        si += a[i] + b[i] + c[i] + m[i] + l[i] + s[i] + p[i] + e[i] + (int)n3[i] + (int)w[i] + (int)(w[i] >> 32) + (int)wi[i] + (int)(wi[i] >> 32) + (int)f[i] + (int)g[i];
    sinki[blockIdx.x * blockDim.x + threadIdx.x] = si;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int M>
void row(const char *name)
{
    int nops = 0;
    for (int b = 1; b <= LEA_; b <<= 1) nops += (M & b) ? 1 : 0;
    printf("%-34s", name);
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int threads = 128 * wps, blocks = 148;
        long long *cyc; int *si;
        cudaMalloc(&cyc, blocks * sizeof(long long)); cudaMalloc(&si, blocks * threads * sizeof(int));
        k<M><<<blocks, threads>>>(cyc, si, 3);
        k<M><<<blocks, threads>>>(cyc, si, 3);
        cudaDeviceSynchronize();
        long long hcyc[148]; cudaMemcpy(hcyc, cyc, sizeof hcyc, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < blocks; ++i) avg += hcyc[i]; avg /= blocks;
        cudaFree(cyc); cudaFree(si);
        printf(" %7.3f", avg / ((double)wps * ITERS * CH));
    }
    printf("   (%d instr / group)\n", nops);
}

int main()
{
    printf("cycles per group of instructions (one per listed class) per sub-partition; 148 CTAs, one per SM\n");
    printf("%-34s %7s %7s %7s %7s\n", "group", "1 w/sp", "2 w/sp", "4 w/sp", "8 w/sp");
    row<DPU>("IDP.2A.U16.U8");
    row<DPS>("IDP.2A.S16.U8 (hi)");
    row<DP4>("IDP.4A.S8.U8");
    row<WID>("IMAD.WIDE.U32 (3 reg)");
    row<WIM>("IMAD.WIDE.U32 (imm, +IADD)");
    row<MN3>("VIMNMX3.U32");
    row<IMA>("IMAD");
    row<ALU>("LOP3");
    row<SHR>("SHF.R.S32");
    row<PRM>("PRMT");
    row<I2P>("I2FP.F32.U32 (+IADD)");
    row<FFM>("FFMA (imm)");
    row<LEA_>("LEA");
    row<DPS | DPU>("IDP.2A x2");
    row<DPS | DPU | WID>("IDP.2A x2 + IMAD.WIDE");
    row<DPS | DPU | WID | MN3>("IDP.2A x2 + IMAD.WIDE + VIMNMX3");
    row<DPS | DPU | WID | MN3 | ALU>("... + LOP3");
    row<DPS | ALU>("IDP.2A + LOP3");
    row<DPS | FFM>("IDP.2A + FFMA");
    row<WID | ALU>("IMAD.WIDE + LOP3");
    row<IMA | ALU>("IMAD + LOP3");
    row<IMA | ALU | FFM>("IMAD + LOP3 + FFMA");
    row<ALU | FFM>("LOP3 + FFMA");
    row<ALU | PRM>("LOP3 + PRMT");
    return 0;
}

// fp64_issue.cu -- micro-benchmark behind DESIGN.md section 7: how do FP64, ALU and IMAD instructions
// share a sub-partition's issue port on this GPU, and what do I2F.F64 / IDP.2A / IMAD.WIDE / IMAD.HI cost?
// Every warp runs CH independent, thread-dependent dependency chains per instruction class; the table
// gives cycles per warp instruction per sub-partition for 1, 2, 4 and 8 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8

enum { F64 = 1, ALU = 2, IMA = 4, I2F = 8, IDP = 16, WID = 32, MHI = 64, I2B = 128, MAG = 256, FAD = 512, FFI = 1024, FFR = 2048 };

template <int M>
__global__ void k(long long *cycles, double *sinkd, int *sinki, double seed, int iseed)
{
    double d[CH]; int a[CH], m[CH], c[CH], p[CH], h[CH]; float fa[CH], fi[CH], fr[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        d[i] = seed + i + threadIdx.x; a[i] = iseed + i * 3 + threadIdx.x; m[i] = iseed * 7 + i + threadIdx.x * 5;
        fa[i] = (float)(threadIdx.x + i); fi[i] = fa[i] * 0.5f; fr[i] = fa[i] * 0.25f;
        c[i] = threadIdx.x + i; p[i] = threadIdx.x * 3 + i; h[i] = 0x7fffffff - threadIdx.x - i; w[i] = i + threadIdx.x;
    }
    const double c1 = seed * 0.5;
    const int x = iseed + threadIdx.x;
    const float fx = (float)seed + 0.001f * threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (M & F64) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c1));
            if (M & ALU) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(x), "r"(it));
            if (M & IMA) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(m[i]) : "r"(x), "r"(it));
            if (M & I2F) { double t; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t) : "r"(c[i])); c[i] = __double2loint(t) | 1; }
            if (M & IDP) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(p[i]) : "r"(x), "r"(it));
            if (M & WID) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(it));
            if (M & MHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(h[i]) : "r"(0xfffffff1u));
            if (M & FAD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa[i]) : "f"(fx));
            if (M & FFI) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(fi[i]) : "f"(fx));      // immediate multiplier
            if (M & FFR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fr[i]) : "f"(fx), "f"(fa[0]));  // three registers
            if (M & I2B) {      // the kernel's conversion: I2F.F64.U8 with a byte selector (byte 1, 2 or 3 of the word)
                double t; const unsigned sh = (unsigned)c[i] >> (8 * (1 + i % 3));
                asm volatile("cvt.rn.f64.u8 %0, %1;" : "=d"(t) : "r"(sh)); c[i] = (__double2loint(t) + c[i]) | 0x01010101;
            }
            if (M & MAG) {      // the 2^52 alternative: byte extract + DADD
                const double t = __hiloint2double(0x43300000, (c[i] >> (8 * (1 + i % 3))) & 0xff) - 4503599627370496.0;
                c[i] = (__double2loint(t) + c[i]) | 0x01010101;
            }
        }
    }
    const long long t1 = clock64();
    double sd = 0; int si = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { sd += d[i] + fa[i] + fi[i] + fr[i]; si += a[i] + m[i] + c[i] + p[i] + h[i] + (int)w[i] + (int)(w[i] >> 32); }
    sinkd[blockIdx.x * blockDim.x + threadIdx.x] = sd;
    sinki[blockIdx.x * blockDim.x + threadIdx.x] = si;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int M>
void row(const char *name)
{
    int nops = 0;
    for (int b = 1; b <= FFR; b <<= 1) nops += (M & b) ? 1 : 0;
    printf("%-24s", name);
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int threads = 128 * wps, blocks = 148;
        long long *cyc; double *sd; int *si;
        cudaMalloc(&cyc, blocks * sizeof(long long)); cudaMalloc(&sd, blocks * threads * sizeof(double)); cudaMalloc(&si, blocks * threads * sizeof(int));
        k<M><<<blocks, threads>>>(cyc, sd, si, 1.0000001, 3);
        k<M><<<blocks, threads>>>(cyc, sd, si, 1.0000001, 3);
        cudaDeviceSynchronize();
        long long hcyc[148]; cudaMemcpy(hcyc, cyc, sizeof hcyc, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < blocks; ++i) avg += hcyc[i]; avg /= blocks;
        cudaFree(cyc); cudaFree(sd); cudaFree(si);
        printf(" %7.3f", avg / ((double)wps * ITERS * CH));      // cycles per GROUP (one of each class in the mode)
    }
    printf("   (%d instr / group)\n", nops);
}

int main()
{
    printf("cycles per group of instructions (one per listed class) per sub-partition; 148 CTAs, one per SM\n");
    printf("%-24s %7s %7s %7s %7s\n", "group", "1 w/sp", "2 w/sp", "4 w/sp", "8 w/sp");
    row<F64>("DADD");
    row<ALU>("LOP3");
    row<IMA>("IMAD");
    row<F64 | ALU>("DADD + LOP3");
    row<F64 | IMA>("DADD + IMAD");
    row<ALU | IMA>("LOP3 + IMAD");
    row<F64 | ALU | IMA>("DADD + LOP3 + IMAD");
    row<I2F>("I2F.F64.U32");
    row<I2F | F64>("I2F.F64.U32 + DADD");
    row<I2B>("I2F.F64.U8.Bn (+IADD+LOP)");
    row<I2B | F64>("I2F.F64.U8.Bn + DADD");
    row<MAG>("PRMT+DADD magic (+2)");
    row<FAD>("FADD");
    row<FFI>("FFMA (immediate)");
    row<FFR>("FFMA (3 registers)");
    row<FAD | IMA>("FADD + IMAD");
    row<FAD | ALU>("FADD + LOP3");
    row<FAD | ALU | IMA>("FADD + LOP3 + IMAD");
    row<FAD | F64>("FADD + DADD");
    row<IDP>("IDP.2A");
    row<WID>("IMAD.WIDE");
    row<MHI>("IMAD.HI");
    row<IDP | IMA>("IDP.2A + IMAD");
    row<WID | IMA>("IMAD.WIDE + IMAD");
    return 0;
}

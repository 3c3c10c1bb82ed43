// pipe_rates.cu — register-only issue-rate microbenchmarks of the integer instructions the stencil kernels are made of
// (sm_100a).  Prints warp-instructions per clock per SM for every op and for a few two-pipe mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_rates tools/pipe_rates.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__device__ __forceinline__ void op(unsigned& a, unsigned& b, unsigned c) {
    if (OP == 0) a = (a & b) ^ c;                                 // LOP3
    else if (OP == 1) a = __vabsdiffu4(a, c) , b ^= a;            // VABSDIFF4 (+ LOP3 to keep both live)
    else if (OP == 2) a = __vmaxs2(a, c);                         // VIMNMX.S16x2
    else if (OP == 3) a = __byte_perm(a, c, 0x4140);              // PRMT
    else if (OP == 4) a = a * c + b;                              // IMAD
    else if (OP == 5) a = __dp4a(a, c, b);                        // IDP.4A
    else if (OP == 6) a = __funnelshift_r(a, c, 8);               // SHF
    else if (OP == 7) a = __vimin3_s16x2(a, b, c);                // VIMNMX3.S16x2
    else if (OP == 8) a = max((int)a, (int)c);                    // VIMNMX (s32)
    else if (OP == 9) a = a + c + b;                              // IADD3
    else if (OP == 10) a = __popc(a) + c;                         // POPC (+ add)
    else if (OP == 11) { a = (a & b) ^ c; b = b * c + a; }        // LOP3 + IMAD pair
    else if (OP == 12) { a = __vmaxs2(a, c); b = b * c + a; }     // VIMNMX.S16x2 + IMAD pair
    else if (OP == 13) { a = __vabsdiffu4(a, c); b = b * c + a; } // VABSDIFF4 + IMAD pair
    else if (OP == 14) { a = __vmaxs2(a, c); b = (b & a) ^ c; }   // VIMNMX.S16x2 + LOP3 pair
    else if (OP == 15) a = __vabsdiffu4(a, c);                    // VABSDIFF4 alone (dependent chain per register)
    else if (OP == 16) a = __dp2a_lo(a, c, b);                    // IDP.2A
    else if (OP == 17) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); x = __hmax2(x, y); a = *reinterpret_cast<unsigned*>(&x); b += c; }   // HMNMX2 (+ IADD keeps the operand changing)
    else if (OP == 18) { a = __vmaxs2(a, b); b += c; }            // VIMNMX.S16x2 (+ IADD), same shape as 17
    else if (OP == 19) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); x = __hmax2(x, y); a = *reinterpret_cast<unsigned*>(&x); b = __vmaxs2(b, c) ; }   // HMNMX2 + VIMNMX.S16x2 pair
    else if (OP == 20) { a = __vmaxs2(a, b); b = __vmins2(b, c); }   // two VIMNMX.S16x2
    else if (OP == 21) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b), z = *reinterpret_cast<const __half2*>(&c); x = __hmax2(x, y); y = __hmin2(y, z); a = *reinterpret_cast<unsigned*>(&x); b = *reinterpret_cast<unsigned*>(&y); }   // two HMNMX2
    else if (OP == 22) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); x = __hmax2(x, y); a = *reinterpret_cast<unsigned*>(&x); b = b * c + a; }   // HMNMX2 + IMAD pair
}

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(unsigned* out, unsigned seed, long long* clk) {
    unsigned a[UNROLL], b[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed * (threadIdx.x + 1) + i; b[i] = seed ^ (threadIdx.x * 77 + i); }
    const unsigned c = seed | 0x01010101u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) op<OP>(a[i], b[i], c + it);
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per) {
    unsigned* d; long long* dc;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4;   // 4 x 256 threads = 32 warps per SM, 8 per scheduler
    cudaMalloc(&d, blocks * 256 * 4); cudaMalloc(&dc, blocks * 8);
    rate_kernel<OP><<<blocks, 256>>>(d, 12345u, dc);
    rate_kernel<OP><<<blocks, 256>>>(d, 54321u, dc);
    cudaDeviceSynchronize();
    long long* h = new long long[blocks];
    cudaMemcpy(h, dc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // warp instructions per SM = 32 warps x ITERS x UNROLL x ops_per (the c + it add is hoisted / counted separately)
    const double wi = 32.0 * ITERS * UNROLL * ops_per;
    printf("%-28s %6.3f warp-inst/clk/SM (%d counted per iteration, %.0f clk)\n", name, wi / avg, ops_per, avg);
    cudaFree(d); cudaFree(dc); delete[] h;
}

int main() {
    run<0>("LOP3", 1); run<15>("VABSDIFF4", 1); run<1>("VABSDIFF4+LOP3", 2); run<2>("VIMNMX.S16x2", 1); run<7>("VIMNMX3.S16x2", 1);
    run<8>("VIMNMX.S32", 1); run<3>("PRMT", 1); run<6>("SHF.R.W", 1); run<9>("IADD3", 1); run<4>("IMAD", 1); run<5>("IDP.4A", 1);
    run<16>("IDP.2A", 1); run<10>("POPC+IADD", 2); run<11>("LOP3+IMAD", 2); run<12>("VIMNMX.S16x2+IMAD", 2);
    run<13>("VABSDIFF4+IMAD", 2); run<14>("VIMNMX.S16x2+LOP3", 2);
    run<17>("HMNMX2+IADD", 2); run<18>("VIMNMX.S16x2+IADD", 2); run<19>("HMNMX2+VIMNMX.S16x2", 2); run<20>("2x VIMNMX.S16x2", 2); run<21>("2x HMNMX2", 2);
    run<22>("HMNMX2+IMAD", 2);
    return 0;
}

// permbench.cu -- development micro-benchmark: Keccak-f[1600] register-resident, no memory traffic.
// Compares instruction-mix variants of the permutation on one GPU (perms/s), to decide what the product
// kernels use.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o permbench permbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__constant__ uint64_t RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
__constant__ uint32_t POW2[32]; // filled at run time so ptxas cannot strength-reduce the multiplies back to shifts

enum { ROT_SHF = 0, ROT_FMA4 = 1, ROT_FMA3 = 2 };

template <int N>
__device__ __forceinline__ uint64_t rol_shf(uint64_t x)
{
    if constexpr (N == 0) return x;
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t rlo, rhi;
    if constexpr (N == 32) { rlo = hi; rhi = lo; }
    else if constexpr (N < 32) { rlo = __funnelshift_l(hi, lo, N); rhi = __funnelshift_l(lo, hi, N); }
    else { rlo = __funnelshift_l(lo, hi, N - 32); rhi = __funnelshift_l(hi, lo, N - 32); }
    return ((uint64_t)rhi << 32) | rlo;
}
// rotate on the FMA pipe: (a << r) + (b >> (32 - r)) as multiply-high + multiply-add by 2^r
template <int N, int MODE>
__device__ __forceinline__ uint64_t rol_fma(uint64_t x)
{
    if constexpr (N == 0) return x;
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if constexpr (N == 32) return ((uint64_t)lo << 32) | hi;
    constexpr int R = N < 32 ? N : N - 32;
    const uint32_t a = N < 32 ? lo : hi, b = N < 32 ? hi : lo; // rotating (b:a) left by R < 32
    const uint32_t m = POW2[R];
    uint32_t nlo, nhi;
    if constexpr (MODE == ROT_FMA4) {
        uint32_t t0, t1;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(t0) : "r"(b), "r"(m));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(nlo) : "r"(a), "r"(m), "r"(t0));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(t1) : "r"(a), "r"(m));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(nhi) : "r"(b), "r"(m), "r"(t1));
    } else {
        // c = { a >> (32-R), a << R } (swapped halves of a * 2^R); b * 2^R + c = { nhi, nlo }
        uint32_t c_lo, c_hi;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(c_lo) : "r"(a), "r"(m));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(c_hi) : "r"(a), "r"(m));
        uint64_t c = ((uint64_t)c_hi << 32) | c_lo, d;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(b), "r"(m), "l"(c));
        nhi = (uint32_t)d; nlo = (uint32_t)(d >> 32);
    }
    return ((uint64_t)nhi << 32) | nlo;
}
template <int N, int MODE>
__device__ __forceinline__ uint64_t rol(uint64_t x)
{
    if constexpr (MODE == ROT_SHF) return rol_shf<N>(x); else return rol_fma<N, MODE>(x);
}

// MODE_RHO: how the 24 rho rotations are done; MODE_TH: how the 5 theta rol-1 are done;
// NFMA: only the first NFMA rho lanes use MODE_RHO, the rest SHF (pipe balancing); FOLD: fold D into theta-apply
template <int MODE_RHO, int MODE_TH, int NFMA, bool FOLD>
__device__ __forceinline__ void keccak_round(uint64_t (&a)[25], uint64_t rc)
{
    uint64_t c[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    uint64_t r[5];
#pragma unroll
    for (int x = 0; x < 5; ++x) r[x] = rol<1, MODE_TH>(c[x]);
#define TH(I) (FOLD ? (a[I] ^ c[((I) % 5 + 4) % 5] ^ r[((I) % 5 + 1) % 5]) : (a[I] ^ d[(I) % 5]))
    uint64_t d[5];
    if constexpr (!FOLD) {
#pragma unroll
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ r[(x + 1) % 5];
    }
#define RP(I, J, R, K) b[J] = rol<R, ((K) < NFMA ? MODE_RHO : ROT_SHF)>(TH(I));
    RP(0, 0, 0, 99)   RP(1, 10, 1, 0)   RP(2, 20, 62, 1)  RP(3, 5, 28, 2)   RP(4, 15, 27, 3)
    RP(5, 16, 36, 4)  RP(6, 1, 44, 5)   RP(7, 11, 6, 6)   RP(8, 21, 55, 7)  RP(9, 6, 20, 8)
    RP(10, 7, 3, 9)   RP(11, 17, 10, 10) RP(12, 2, 43, 11) RP(13, 12, 25, 12) RP(14, 22, 39, 13)
    RP(15, 23, 41, 14) RP(16, 8, 45, 15) RP(17, 18, 15, 16) RP(18, 3, 21, 17) RP(19, 13, 8, 18)
    RP(20, 14, 18, 19) RP(21, 24, 2, 20) RP(22, 9, 61, 21) RP(23, 19, 56, 22) RP(24, 4, 14, 23)
#undef RP
#undef TH
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
        for (int x = 0; x < 5; ++x) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
    }
    a[0] ^= rc;
}

template <int UNROLL, int MODE_RHO, int MODE_TH, int NFMA, bool FOLD>
__global__ void __launch_bounds__(128) perm_kernel(uint64_t* out, int iters)
{
    uint64_t a[25];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 25; ++i) a[i] = tid * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int r = 0; r < 24; r += UNROLL) {
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) keccak_round<MODE_RHO, MODE_TH, NFMA, FOLD>(a, RC[r + k]);
        }
    }
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 25; ++i) x ^= a[i];
    out[tid] = x;
}

template <int UNROLL, int MODE_RHO, int MODE_TH, int NFMA, bool FOLD>
void run(const char* name, int blocks_per_sm, uint64_t* d_out, uint64_t* h_ref)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, threads = 128, iters = 2000;
    perm_kernel<UNROLL, MODE_RHO, MODE_TH, NFMA, FOLD><<<blocks, threads>>>(d_out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    perm_kernel<UNROLL, MODE_RHO, MODE_TH, NFMA, FOLD><<<blocks, threads>>>(d_out, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    uint64_t h[4];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    const double perms = (double)blocks * threads * iters;
    const bool same = h_ref[0] == 0 ? true : (h[0] == h_ref[0] && h[3] == h_ref[3]);
    if (h_ref[0] == 0) { for (int i = 0; i < 4; ++i) h_ref[i] = h[i]; }
    printf("{\"variant\": \"%s\", \"blocks_per_sm\": %d, \"gperm_s\": %.3f, \"ms\": %.3f, \"matches_baseline\": %s, \"err\": \"%s\"}\n", name,
           blocks_per_sm, perms / ms / 1e6, ms, same ? "true" : "false", cudaGetErrorString(e));
    fflush(stdout);
}

int main()
{
    uint32_t p[32];
    for (int i = 0; i < 32; ++i) p[i] = 1u << i;
    cudaMemcpyToSymbol(POW2, p, sizeof p);
    uint64_t* d_out;
    cudaMalloc(&d_out, 8ull * 148 * 16 * 128 * 2);
    for (int bps : {3, 4, 6, 8}) {
        uint64_t ref[4] = {0, 0, 0, 0};
        run<2, ROT_SHF, ROT_SHF, 0, false>("shf_u2", bps, d_out, ref);
        run<2, ROT_SHF, ROT_SHF, 0, true>("shf_u2_fold", bps, d_out, ref);
        run<1, ROT_SHF, ROT_SHF, 0, true>("shf_u1_fold", bps, d_out, ref);
        run<4, ROT_SHF, ROT_SHF, 0, true>("shf_u4_fold", bps, d_out, ref);
        run<24, ROT_SHF, ROT_SHF, 0, true>("shf_u24_fold", bps, d_out, ref);
        run<2, ROT_FMA4, ROT_SHF, 24, true>("fma4_all_rho", bps, d_out, ref);
        run<2, ROT_FMA3, ROT_SHF, 24, true>("fma3_all_rho", bps, d_out, ref);
        run<2, ROT_FMA4, ROT_FMA4, 24, true>("fma4_all_rho_th", bps, d_out, ref);
        run<2, ROT_FMA3, ROT_FMA3, 24, true>("fma3_all_rho_th", bps, d_out, ref);
        run<2, ROT_FMA4, ROT_SHF, 12, true>("fma4_12", bps, d_out, ref);
        run<2, ROT_FMA3, ROT_SHF, 12, true>("fma3_12", bps, d_out, ref);
        run<2, ROT_FMA4, ROT_SHF, 16, true>("fma4_16", bps, d_out, ref);
        run<2, ROT_FMA3, ROT_SHF, 16, true>("fma3_16", bps, d_out, ref);
        run<2, ROT_FMA3, ROT_SHF, 20, true>("fma3_20", bps, d_out, ref);
        run<2, ROT_FMA4, ROT_SHF, 8, true>("fma4_8", bps, d_out, ref);
    }
    return 0;
}

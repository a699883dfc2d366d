// smem_ops.cu — micro-benchmark behind the design of spf_quad_kernel's relaxation step:
// cost of shared-memory operations with random (bank-conflicting) addresses, as a function
// of the number of active lanes: LDS, STS.32, STS.U8, ATOMS.MIN, ATOMS.OR (fire-and-forget).
// One CTA per SM x 3 (as the kernel runs), 384 threads; every thread executes ITER operations
// on pseudo-random words of a 14k-word array; reported: SM cycles per warp instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/smem_ops_bench scripts/ubench/smem_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int T = 384, ITER = 4096, N = 13856;

template <int OP>
__global__ void __launch_bounds__(T, 3) k(uint32_t active_mod, unsigned long long *cyc, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < N; i += T) sm[i] = 0x7fffffffu;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const bool act = (threadIdx.x % active_mod) == 0;
    uint32_t acc = 0;
    uint8_t *sm8 = reinterpret_cast<uint8_t *>(sm);
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITER; ++it) {
        x = x * 1664525u + 1013904223u;
        const uint32_t idx = (x >> 8) % N;
        if (act) {
            if (OP == 0) acc += sm[idx];
            if (OP == 1) sm[idx] = x;
            if (OP == 2) sm8[idx] = (uint8_t)x;
            if (OP == 3) atomicMin(&sm[idx], x >> 4);
            if (OP == 4) atomicOr(&sm[idx >> 5], 1u << (idx & 31));
            if (OP == 5) { atomicMin(&sm[idx], x >> 4); atomicOr(&sm[(idx >> 5)], 1u << (idx & 31)); }
            if (OP == 6) { atomicMin(&sm[idx], x >> 4); sm8[idx] = 1; }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345) sink[0] = acc;
}

template <int OP>
void run(const char *name, int sms) {
    unsigned long long *d; uint32_t *s;
    cudaMalloc(&d, sizeof(*d) * sms * 3); cudaMalloc(&s, 4);
    cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, N * 4);
    for (uint32_t mod : {1u, 2u, 4u, 8u, 16u, 32u}) {
        k<OP><<<sms * 3, T, N * 4>>>(mod, d, s);
        cudaDeviceSynchronize();
        unsigned long long h[3 * 256];
        cudaMemcpy(h, d, sizeof(*d) * sms * 3, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms * 3; ++i) avg += h[i];
        avg /= sms * 3;
        // per SM: 3 CTAs x 12 warps x ITER warp-instructions in `avg` cycles
        printf("%-14s active lanes/warp %2u: %7.2f SM-cycles per warp-instruction (loop overhead included)\n", name, 32 / mod,
               avg / (3.0 * 12 * ITER));
    }
    cudaFree(d); cudaFree(s);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<0>("LDS", sms); run<1>("STS.32", sms); run<2>("STS.U8", sms); run<3>("ATOMS.MIN", sms);
    run<4>("ATOMS.OR", sms); run<5>("ATOMS.MIN+OR", sms); run<6>("ATOMS.MIN+STS8", sms);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

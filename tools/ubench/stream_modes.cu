// Micro-benchmark: how fast can a low-occupancy persistent kernel (the mesher's shape: ~16 warps/SM) stream 128 KB
// chunks, by staging mechanism? Modes: 0 = LDG.128 to registers (K per lane per batch), 1 = cp.async 16B ring,
// 2 = cp.async.bulk (TMA 1D) + mbarrier ring. Each "consumes" the data with a cheap XOR reduction.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int CH = 32768; // u32 per chunk

template <int K>
__global__ void __launch_bounds__(256) mode_ldg(const uint4 *__restrict__ d, size_t n_chunks, unsigned *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned acc = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint4 *base = d + c * (CH / 4);
        // 8192 uint4 per chunk = 256 warp-loads; warp w takes loads w, w+nw, ...
        for (int g0 = warp; g0 < 256; g0 += nw * K) {
            uint4 v[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                int g = g0 + k * nw;
                if (g < 256) v[k] = __ldg(base + g * 32 + lane);
                else v[k] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < K; k++) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
        }
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ void cp_async16(void *s, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// warp-private ring, DEPTH units of 3 warp-loads ahead
template <int DEPTH>
__global__ void __launch_bounds__(256) mode_cpasync(const uint4 *__restrict__ d, size_t n_chunks, unsigned *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int SLOTS = DEPTH + 2;
    uint4 *ring = reinterpret_cast<uint4 *>(smem) + warp * SLOTS * 96;
    unsigned acc = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint4 *base = d + c * (CH / 4);
        const int n_units = (256 / 3 + nw - 1 - warp) / nw; // units u*nw+warp, 3 loads each (last partial ignored: 255 loads)
        auto issue = [&](int u) {
            if (u < n_units) {
                const int g = (u * nw + warp) * 3;
#pragma unroll
                for (int j = 0; j < 3; j++) cp_async16(ring + (u % SLOTS) * 96 + j * 32 + lane, base + (g + j) * 32 + lane);
            }
            cp_commit();
        };
        for (int u = 0; u < DEPTH; u++) issue(u);
        for (int u = 0; u < n_units; u++) {
            __syncwarp();
            issue(u + DEPTH);
            cp_wait<DEPTH>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 3; j++) {
                uint4 v = ring[(u % SLOTS) * 96 + j * 32 + lane];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
        cp_wait<0>();
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// TMA 1D bulk copies: each warp owns SLOTS slots of 4 KB (one y-layer) with one mbarrier each
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            (uint32_t)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}

template <int SLOTS>
__global__ void __launch_bounds__(256) mode_bulk(const uint4 *__restrict__ d, size_t n_chunks, unsigned *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint4 *ring = reinterpret_cast<uint4 *>(smem) + warp * SLOTS * 256;              // SLOTS x 4 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * SLOTS * 4096) + warp * SLOTS;
    if (lane == 0)
        for (int i = 0; i < SLOTS; i++) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    uint32_t issued = 0, consumed = 0; // running unit counters of this warp (slot = n % SLOTS, parity = (n / SLOTS) & 1)
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint4 *base = d + c * (CH / 4);
        const int n_units = (32 + nw - 1 - warp) / nw; // layers warp, warp+nw, ...
        int iu = 0;
        auto issue = [&]() {
            if (iu < n_units) {
                if (lane == 0) {
                    uint64_t *b = bars + issued % SLOTS;
                    mbar_expect(b, 4096);
                    bulk_copy(ring + (issued % SLOTS) * 256, base + (size_t)(iu * nw + warp) * 256, 4096, b);
                }
                issued++;
                iu++;
            }
        };
        for (int k = 0; k < SLOTS - 1; k++) issue();
        for (int u = 0; u < n_units; u++) {
            __syncwarp(); // all lanes done with the slot about to be refilled
            issue();
            mbar_wait(bars + consumed % SLOTS, (consumed / SLOTS) & 1);
            const uint4 *src = ring + (consumed % SLOTS) * 256;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint4 v = src[j * 32 + lane];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
            consumed++;
        }
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main(int argc, char **argv) {
    size_t n_chunks = 65536; // 8 GiB
    uint4 *d;
    unsigned *out;
    CK(cudaMalloc(&d, n_chunks * CH * 4));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(d, 0, n_chunks * CH * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto report = [&](const char *name, float ms) { printf("%-40s %8.3f ms  %7.1f GB/s\n", name, ms, n_chunks * CH * 4.0 / ms / 1e6); };
#define RUN(name, launch)                       \
    {                                           \
        launch;                                 \
        CK(cudaDeviceSynchronize());            \
        cudaEventRecord(e0);                    \
        for (int r = 0; r < 3; r++) { launch; } \
        cudaEventRecord(e1);                    \
        CK(cudaEventSynchronize(e1));           \
        float ms;                               \
        cudaEventElapsedTime(&ms, e0, e1);      \
        report(name, ms / 3);                   \
    }
    for (int grid : {296, 592}) {
        printf("grid %d x 256 threads\n", grid);
        RUN("ldg K=4", (mode_ldg<4><<<grid, 256>>>(d, n_chunks, out)));
        RUN("ldg K=8", (mode_ldg<8><<<grid, 256>>>(d, n_chunks, out)));
        RUN("ldg K=16", (mode_ldg<16><<<grid, 256>>>(d, n_chunks, out)));
        {
            CK(cudaFuncSetAttribute(mode_cpasync<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 96 * 16));
            RUN("cp.async depth 2 (3 KB/warp in flight)", (mode_cpasync<2><<<grid, 256, 8 * 4 * 96 * 16>>>(d, n_chunks, out)));
            CK(cudaFuncSetAttribute(mode_cpasync<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 6 * 96 * 16));
            RUN("cp.async depth 4 (6 KB/warp in flight)", (mode_cpasync<4><<<grid, 256, 8 * 6 * 96 * 16>>>(d, n_chunks, out)));
        }
        {
            int sm2 = 8 * 2 * 4096 + 8 * 2 * 8, sm3 = 8 * 3 * 4096 + 8 * 3 * 8;
            CK(cudaFuncSetAttribute(mode_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
            RUN("bulk TMA 2 slots (4 KB/warp in flight)", (mode_bulk<2><<<grid, 256, sm2>>>(d, n_chunks, out)));
            CK(cudaFuncSetAttribute(mode_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm3));
            RUN("bulk TMA 3 slots (8 KB/warp in flight)", (mode_bulk<3><<<grid, 256, sm3>>>(d, n_chunks, out)));
        }
    }
    return 0;
}

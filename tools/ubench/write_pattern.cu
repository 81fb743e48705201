// write_pattern.cu — how fast can a B200 absorb emit_kernel's output pattern? Every warp writes "batches" of BATCH bytes in
// pieces of 1280 bytes (one vertex pass), batch b of the grid-stride loop at byte offset b * BATCH.
//   mode 0: st.global.cs.v4 from registers      mode 1: same through shared memory (LDS.128 + STG.128)
//   mode 2: cp.async.bulk shared -> global      mode 3: st.global.v4 (default policy)
// BATCH 6400 = 128-byte aligned pieces; 6000 / 6160 = 16-byte aligned pieces that cut 32-byte sectors and 128-byte lines.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 4) wr(uint4 *out, unsigned long long n_batches, unsigned batch_bytes) {
    __shared__ __align__(16) uint4 stage[8][2][82];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long step = (unsigned long long)gridDim.x * 8;
    unsigned np = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * 8 + warp; b < n_batches; b += step) {
        char *base = reinterpret_cast<char *>(out) + b * batch_bytes;
        for (unsigned off = 0; off < batch_bytes; off += 1280) {
            const unsigned bytes = min(1280u, batch_bytes - off);
            const unsigned units = bytes / 16;
            uint4 *dst = reinterpret_cast<uint4 *>(base + off);
            const uint4 v = make_uint4(lane, off, (unsigned)b, 7u);
            if (MODE == 0 || MODE == 3) {
                for (unsigned k = lane; k < units; k += 32) {
                    if (MODE == 0) __stcs(dst + k, v);
                    else dst[k] = v;
                }
            } else {
                uint4 *buf = stage[warp][np & 1];
                np++;
                if (MODE == 2) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                for (unsigned k = lane; k < units; k += 32) buf[k] = v;
                if (MODE == 2) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                     "r"((unsigned)__cvta_generic_to_shared(buf)), "r"(units * 16u)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    __syncwarp();
                    for (unsigned k = lane; k < units; k += 32) __stcs(dst + k, buf[k]);
                    __syncwarp();
                }
            }
        }
    }
    if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const size_t total = 40ull << 30;
    uint4 *out;
    if (cudaMalloc(&out, total + 65536) != cudaSuccess) return 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const unsigned sizes[3] = {6400, 6000, 6160};
    for (int mode = 0; mode < 4; mode++)
        for (int si = 0; si < 3; si++) {
            const unsigned long long nb = total / sizes[si];
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                if (mode == 0) wr<0><<<592, 256>>>(out, nb, sizes[si]);
                if (mode == 1) wr<1><<<592, 256>>>(out, nb, sizes[si]);
                if (mode == 2) wr<2><<<592, 256>>>(out, nb, sizes[si]);
                if (mode == 3) wr<3><<<592, 256>>>(out, nb, sizes[si]);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            printf("mode %d batch %u: %.3f ms  %.0f GB/s  (%s)\n", mode, sizes[si], best, nb * (double)sizes[si] / best / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

// store_pattern.cu -- memory-system probe: how fast can TMA tensor stores stream a (N, A) fp64 matrix when each
// warp walks along the columns of RB rows at a time (box = RB rows x CB cols)?  No arithmetic; isolates the
// write path (TMA -> L2 -> HBM) from the kernel's compute.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/store_pattern tools/store_pattern.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BUFS>
__global__ void __launch_bounds__(128) store_kernel(const __grid_constant__ CUtensorMap map, int rb, int cb, int n_col_chunks,
                                                    long long n_row_groups, int delay) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = smem + warp * BUFS * 4096;
    for (int i = lane; i < BUFS * 4096 / 8; i += 32) reinterpret_cast<double*>(stage)[i] = 1.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long g = (long long)blockIdx.x * 4 + warp;
    if (g >= n_row_groups) return;
    double acc = lane;
    for (int c = 0; c < n_col_chunks; ++c) {
        for (int d = 0; d < delay; ++d) acc = fma(acc, 1.0000001, 1e-9);   // stand-in for the per-chunk compute time
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map),
                         "r"(c * cb), "r"((int)(g * rb)), "r"(smem_u32(stage + (c % BUFS) * 4096))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BUFS - 1) : "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (acc == 1234.5) stage[0] = 1;
}

// mode B: per iteration lane 0 issues `m` back-to-back 32x16 boxes for ADJACENT column blocks (one commit group)
__global__ void __launch_bounds__(128) store_multi_kernel(const __grid_constant__ CUtensorMap map, int m, int n_col_chunks,
                                                          long long n_row_groups, int wait_all) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = smem + warp * m * 4096;
    for (int i = lane; i < m * 4096 / 8; i += 32) reinterpret_cast<double*>(stage)[i] = 1.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long g = (long long)blockIdx.x * 4 + warp;
    if (g >= n_row_groups) return;
    for (int c = 0; c < n_col_chunks; c += m) {
        if (lane == 0) {
            for (int q = 0; q < m && c + q < n_col_chunks; ++q) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map),
                             "r"((c + q) * 16), "r"((int)(g * 32)), "r"(smem_u32(stage + q * 4096))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (wait_all) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// mode C: every lane streams ITS OWN row piece (cb columns) with a 1-D bulk copy from a padded shared-memory row;
// 32 rows per warp (the thread-per-sample mapping), `bufs` tile buffers per warp
__global__ void __launch_bounds__(128) store_rows_kernel(double* out, int A, int cb, int pitch, int n_col_chunks,
                                                         long long n_rows, int bufs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* stage = reinterpret_cast<double*>(smem) + (size_t)warp * bufs * 32 * pitch;
    for (int i = lane; i < bufs * 32 * pitch; i += 32) stage[i] = 1.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long row = ((long long)blockIdx.x * 4 + warp) * 32 + lane;
    if (row - lane >= n_rows) return;
    for (int c = 0; c < n_col_chunks; ++c) {
        const int cols = min(cb, A - c * cb);
        if (row < n_rows) {
            double* g = out + row * A + (long long)c * cb;
            const uint32_t src = smem_u32(stage + ((c % bufs) * 32 + lane) * pitch);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(src), "r"(cols * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (bufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// mode D: 3-D tensor view (16 cols, A/16 column blocks, N rows): ONE op writes `cbs` adjacent 128-byte column blocks of
// `rows` rows from shared-memory sub-tiles laid out [column block][row][128 B] (each sub-tile 128B-swizzle friendly).
// A warp owns 32 rows (thread-per-sample mapping) and buffers `cbs` chunks of 16 columns before storing.
__global__ void __launch_bounds__(128) store_3d_kernel(const __grid_constant__ CUtensorMap map, int rows, int cbs, int n_colblocks,
                                                       long long n_row_groups, int bufs, int order) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_bytes = 32 * cbs * 128;                  // 32 rows x cbs column blocks
    unsigned char* stage = smem + (size_t)warp * bufs * tile_bytes;
    for (int i = lane; i < bufs * tile_bytes / 8; i += 32) reinterpret_cast<double*>(stage)[i] = 1.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long g = (long long)blockIdx.x * 4 + warp;   // 32-row group
    if (g >= n_row_groups) return;
    int it = 0;
    for (int cb0 = 0; cb0 + cbs <= n_colblocks; cb0 += cbs, ++it) {
        if (lane == 0) {
            unsigned char* buf = stage + (it % bufs) * tile_bytes;
            for (int rg = 0; rg < 32 / rows; ++rg) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map),
                             "r"(0), "r"(order ? (int)(g * 32 + rg * rows) : cb0), "r"(order ? cb0 : (int)(g * 32 + rg * rows)),
                             "r"(smem_u32(buf + rg * rows * cbs * 128))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (bufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const long long N = 1000000;
    const int A = argc > 1 ? atoi(argv[1]) : 256;
    double* out;
    cudaMalloc(&out, (size_t)N * A * 8);
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(store_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int shapes[][2] = {{32, 16}, {16, 32}, {8, 64}, {4, 128}, {2, 256}};
    for (int delay : {0, 400}) {
        for (auto& sh : shapes) {
            const int rb = sh[0], cb = sh[1];
            if (cb > A) continue;
            CUtensorMap map;
            const cuuint64_t dims[2] = {(cuuint64_t)A, (cuuint64_t)N};
            const cuuint64_t strides[1] = {(cuuint64_t)A * 8};
            const cuuint32_t box[2] = {(cuuint32_t)cb, (cuuint32_t)rb};
            const cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d for box %dx%d\n", (int)r, rb, cb); continue; }
            const long long groups = (N + rb - 1) / rb;
            const int chunks = (A + cb - 1) / cb;
            const unsigned blocks = (unsigned)((groups + 3) / 4);
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                store_kernel<2><<<blocks, 128, 4 * 2 * 4096 + 1024>>>(map, rb, cb, chunks, groups, delay);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            cudaError_t err = cudaGetLastError();
            printf("A=%d delay=%d box %2d rows x %3d cols: %.3f ms  %.0f GB/s  (%s)\n", A, delay, rb, cb, best,
                   (double)N * A * 8 / best / 1e6, cudaGetErrorString(err));
        }
    }
    {
        cudaFuncSetAttribute(store_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        CUtensorMap map;
        const cuuint64_t dims[2] = {(cuuint64_t)A, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)A * 8};
        const cuuint32_t box[2] = {16, 32};
        const cuuint32_t es[2] = {1, 1};
        enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const long long groups = (N + 31) / 32;
        const int chunks = (A + 15) / 16;
        for (int wait_all : {1, 0}) for (int m : {1, 2, 4, 8}) {
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                store_multi_kernel<<<(unsigned)((groups + 3) / 4), 128, 4 * m * 4096 + 1024>>>(map, m, chunks, groups, wait_all);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("A=%d multi-issue m=%d x (32 rows x 16 cols) swizzle128, wait_%s: %.3f ms  %.0f GB/s (%s)\n", A, m,
                   wait_all ? "all" : "prev", best, (double)N * A * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {
        cudaFuncSetAttribute(store_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int bufs : {1, 2}) for (int cb : {16, 32, 64, 128}) {
            const int pitch = cb + 2;
            const size_t smem_bytes = (size_t)4 * bufs * 32 * pitch * 8;
            if (smem_bytes > 200 * 1024) continue;
            const int chunks = (A + cb - 1) / cb;
            const unsigned blocks = (unsigned)((N + 127) / 128);
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                store_rows_kernel<<<blocks, 128, smem_bytes>>>(out, A, cb, pitch, chunks, N, bufs);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("A=%d per-lane 1-D bulk rows: 32 rows x %3d cols per warp, bufs=%d (%zu KB smem/block): %.3f ms  %.0f GB/s (%s)\n", A,
                   cb, bufs, smem_bytes / 1024, best, (double)N * A * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {
        cudaFuncSetAttribute(store_3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        const int ncb = A / 16;
        const int cfgs[][2] = {{32, 1}, {16, 2}, {8, 4}, {32, 2}, {16, 4}, {32, 4}, {8, 8}};   // {rows per op, column blocks per op}
        for (int order : {0, 1}) for (int bufs : {1, 2}) for (auto& c : cfgs) {
            const int rows = c[0], cbs = c[1];
            const size_t smem_bytes = (size_t)4 * bufs * 32 * cbs * 128 + 1024;
            if (smem_bytes > 200 * 1024) continue;
            CUtensorMap map;
            // order == 0: dims (cols16, column blocks, rows): smem [row][block][128 B], a row's pieces are emitted consecutively
            // order == 1: dims (cols16, rows, column blocks): smem [block][row][128 B] (swizzle-friendly for thread-per-row writers)
            const cuuint64_t dims[3] = {16, order ? (cuuint64_t)N : (cuuint64_t)ncb, order ? (cuuint64_t)ncb : (cuuint64_t)N};
            const cuuint64_t strides[2] = {order ? (cuuint64_t)A * 8 : 128, order ? 128 : (cuuint64_t)A * 8};
            const cuuint32_t box[3] = {16, order ? (cuuint32_t)rows : (cuuint32_t)cbs, order ? (cuuint32_t)cbs : (cuuint32_t)rows};
            const cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("3d encode failed %d\n", (int)r); continue; }
            const long long groups = (N + 31) / 32;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                store_3d_kernel<<<(unsigned)((groups + 3) / 4), 128, smem_bytes>>>(map, rows, cbs, ncb, groups, bufs, order);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const int used_cb = ncb / cbs * cbs;
            printf("A=%d 3-D[order %d] ops: %2d rows x %d col-blocks (%4d B pieces) per op, warp tile 32 rows x %3d cols, bufs=%d (%3zu KB/block): %.3f ms  %.0f GB/s (%s)\n",
                   A, order, rows, cbs, cbs * 128, cbs * 16, bufs, smem_bytes / 1024, best, (double)N * used_cb * 128 / best / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}

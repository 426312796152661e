// CUDA programming guide TMA example (2D tile load + store), reduced: known-good recipe to test the platform.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 32;

__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    out[threadIdx.x] = smem_buffer[0][threadIdx.x % SW];
}

int main()
{
    std::vector<int> h((size_t)GW * GH);
    for (size_t i = 0; i < h.size(); i++) h[i] = (int)i;
    int *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 128 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))sym;
    CUtensorMap m{};
    cuuint64_t size[2] = {GW, GH};
    cuuint64_t stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SW, SH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d query %d\n", (int)r, (int)q);
    k<<<1, 128>>>(m, 64, 3, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    int res[4];
    cudaMemcpy(res, o, 16, cudaMemcpyDeviceToHost);
    printf("got %d %d %d %d expect %d..\n", res[0], res[1], res[2], res[3], 3 * GW + 64);
    return 0;
}

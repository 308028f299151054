// stand-alone probe of the TMA box load used by the tile kernels (mc_tma.cuh): loads an 18 x 18 x 9 box of
// doubles at a (possibly negative) origin and checks it against the tensor.
#include <cstdio>
#include <vector>
#include "../../moosecad_b200/csrc/mc_tma.cuh"
using namespace mc;
constexpr int H = 18, HZ = 9, HX = 20;  // inner extent 20: the box starts at an even x (16-byte aligned)
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, double* out) {
    __shared__ __align__(128) double box[HZ * H * HX];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, HZ * H * HX * 8);
        tma_load_3d(box, &tmap, x, y, z, &bar);
    }
    mbar_wait(&bar, 0u);
    for (int i = threadIdx.x; i < HZ * H * HX; i += blockDim.x) out[i] = box[i];
}
int main() {
    const int NX = 33, PX = 48, NY = 20, NZ = 12;
    std::vector<double> h((size_t)PX * NY * NZ);
    for (int z = 0; z < NZ; ++z) for (int y = 0; y < NY; ++y) for (int x = 0; x < PX; ++x) h[((size_t)z * NY + y) * PX + x] = 1 + x + 100 * y + 10000 * z;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, HZ * H * HX * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, fn);
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {NX, NY, NZ}, strides[2] = {PX * 8, (cuuint64_t)PX * NY * 8};
    cuuint32_t box[3] = {HX, H, HZ}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    const int origins[6][3] = {{16, 3, 4}, {14, 3, 4}, {-2, 3, 4}, {2, -1, 4}, {2, 3, -1}, {-2, -1, -1}};
    for (int t = 0; t < 6; ++t) {
        const int x0 = origins[t][0], y0 = origins[t][1], z0 = origins[t][2];
        k<<<1, 256>>>(tm, x0, y0, z0, o);
        e = cudaDeviceSynchronize();
        printf("kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("origin (%d,%d,%d) FAILED\n", x0, y0, z0); return 1; }
        std::vector<double> got(HZ * H * HX);
        cudaMemcpy(got.data(), o, got.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int z = 0; z < HZ; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < HX; ++x) {
            const int X = x0 + x, Y = y0 + y, Z = z0 + z;
            const double want = (X >= 0 && Y >= 0 && Z >= 0 && X < NX && Y < NY && Z < NZ) ? 1 + X + 100 * Y + 10000 * Z : 0.0;
            if (got[(z * H + y) * HX + x] != want) ++bad;
        }
        printf("origin (%d,%d,%d): %d mismatches\n", x0, y0, z0, bad);
    }
    return 0;
}

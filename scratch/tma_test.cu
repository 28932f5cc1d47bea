// Isolated TMA 3-D box-load probe.  usage: tma_test <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
struct alignas(64) Blob { unsigned char b[128]; };
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p);}  
template<int ROWS>
__global__ void k(const __grid_constant__ Blob tmap, const Blob* gmap, int use_global, int c0, int c1, int c2, float* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float* buf = (float*)sm;
    uint64_t* bar = (uint64_t*)(sm + ROWS*64*4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(ROWS*64*4) : "memory");
        const void* d = use_global ? (const void*)gmap : (const void*)&tmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            :: "r"(s32(buf)), "l"(d), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(s32(bar)) : "memory");
    }
    for (int i = threadIdx.x; i < ROWS*64; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*ENC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int U = 96, A = 40, V = 64;
    std::vector<float> h((size_t)U*A*V);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float* d; cudaMalloc(&d, h.size()*4); cudaMemcpy(d, h.data(), h.size()*4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    ENC enc = (ENC)p;
    Blob map;
    int rows = (variant & 1) ? 16 : 30;
    cuuint64_t dims[3] = {U, A, V};
    cuuint64_t strides[2] = {(cuuint64_t)U*4, (cuuint64_t)U*4*A};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)rows};
    cuuint32_t es[3] = {1,1,1};
    CUresult r = enc((CUtensorMap*)&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     (variant & 8) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode=%d rows=%d\n", variant, (int)r, rows);
    Blob* gmap; cudaMalloc(&gmap, sizeof(Blob)); cudaMemcpy(gmap, &map, sizeof(Blob), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 30*64*4);
    int c0 = argc > 2 ? atoi(argv[2]) : 0, c1 = 3, c2 = argc > 3 ? atoi(argv[3]) : 4;
    printf("  c0=%d c2=%d\n", c0, c2);
    int ug = (variant & 4) ? 1 : 0;
    if (rows == 16) k<16><<<1, 128, 16*64*4 + 64>>>(map, gmap, ug, c0, c1, c2, out);
    else k<30><<<1, 128, 30*64*4 + 64>>>(map, gmap, ug, c0, c1, c2, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  sync: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(rows*64); cudaMemcpy(o.data(), out, o.size()*4, cudaMemcpyDeviceToHost);
        // expected: element (r, c) = proj[(c2+r), c1, c0+c] or 0 when out of bounds
        int bad = 0;
        for (int rr = 0; rr < rows; ++rr) for (int c = 0; c < 64; ++c) {
            int u = c0 + c, v = c2 + rr; float ex = (u >= 0 && u < U && v >= 0 && v < V) ? (float)(((size_t)v*A + c1)*U + u) : 0.f;
            if (o[rr*64+c] != ex) ++bad;
        }
        printf("  mismatches: %d   o[0]=%g o[65]=%g\n", bad, o[0], o[65]);
    }
    return 0;
}

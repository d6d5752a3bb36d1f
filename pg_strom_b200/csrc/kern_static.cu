/*
 * kern_static.cu - device code that does not depend on a query and is
 * therefore compiled ahead of time by nvcc for sm_100a (everything query
 * specific goes through NVRTC, see cuda_layer.cpp).
 */
#include <cuda_runtime.h>
#include <stdint.h>

/* streaming fill used to evict L2 between timed iterations: 128-bit stores,
 * grid-stride, one CTA wave per SM */
__global__ void
pgs_static_fill_u128(uint4 *dst, size_t nvec, uint32_t value)
{
    uint4 v = make_uint4(value, value, value, value);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
         i < nvec;
         i += (size_t)gridDim.x * blockDim.x)
        dst[i] = v;
}

/* streaming read checksum: used by the bench to measure what a plain
 * 128-bit coalesced read of the same bytes achieves on this device */
__global__ void
pgs_static_read_u128(const uint4 *src, size_t nvec, unsigned long long *out)
{
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
         i < nvec;
         i += (size_t)gridDim.x * blockDim.x)
    {
        uint4 v = __ldg(src + i);
        acc += (unsigned long long)v.x + v.y + v.z + v.w;
    }
    for (int d = 16; d > 0; d >>= 1)
        acc += __shfl_xor_sync(0xffffffffU, acc, d);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(out, acc);
}

extern "C" int
pgs_static_launch_fill(void *dst, size_t bytes, uint32_t value, int sm_count, void *stream)
{
    pgs_static_fill_u128<<<sm_count * 8, 256, 0, (cudaStream_t)stream>>>(
        (uint4 *)dst, bytes / 16, value);
    return (int)cudaGetLastError();
}

extern "C" int
pgs_static_launch_read(const void *src, size_t bytes, unsigned long long *out,
                       int sm_count, void *stream)
{
    pgs_static_read_u128<<<sm_count * 8, 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)src, bytes / 16, out);
    return (int)cudaGetLastError();
}

// String interning on the device: the hashing half of the reference's family join.
//
// The reference keys its join on STRINGS: DataSet::getGeneFamilies (DataSet.cpp:188-202) looks every gene ID up in
// the blast table's map<string, string> (GeneFamily.cpp:12-16, 38-48) and GHM::buildMatrixFast hashes the family
// strings again for every list pair (GHM.cpp:34-50).  Here strings are replaced once, before the dataset is
// created, by dense integers: equal strings get equal ids, ids number the distinct strings in order of first
// occurrence (deterministic, independent of thread timing).  K1 then joins on integers.
//
//   insert   one thread per string: FNV-1a hash, linear probing in a table of representative indices; a slot is
//            claimed with one CAS, a claimed slot is compared byte by byte (exact: hash collisions only cost a probe)
//   classes  smallest index of every class (atomicMin), "is the first of its class" flags
//   scan     exclusive prefix sum of the flags (block sums, one block over the sums, add)
//   ids      id = rank of the class's first member
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ bool same_string(const char* __restrict__ blob, const int64_t* __restrict__ off, int64_t a, int64_t b)
{
    const int64_t a0 = off[a], b0 = off[b];
    const int64_t len = off[a + 1] - a0;
    if (off[b + 1] - b0 != len) return false;
    for (int64_t k = 0; k < len; k++)
        if (blob[a0 + k] != blob[b0 + k]) return false;
    return true;
}

__global__ void __launch_bounds__(256) intern_insert_kernel(const char* __restrict__ blob, const int64_t* __restrict__ off, int64_t n,
                                                            int32_t* table, uint32_t mask, int32_t* __restrict__ rep)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t h = 1469598103934665603ull;   // FNV-1a
    for (int64_t k = off[i]; k < off[i + 1]; k++) h = (h ^ (uint8_t)blob[k]) * 1099511628211ull;
    uint32_t slot = (uint32_t)(h ^ (h >> 32)) & mask;
    while (true) {
        int32_t cur = table[slot];
        if (cur < 0) {
            cur = atomicCAS(&table[slot], -1, (int32_t)i);
            if (cur < 0) {   // this string represents its class (until the smallest index is taken below)
                rep[i] = (int32_t)i;
                return;
            }
        }
        if (same_string(blob, off, cur, i)) {
            rep[i] = cur;
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) intern_min_kernel(const int32_t* __restrict__ rep, int64_t n, int32_t* cmin)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMin(&cmin[rep[i]], (int32_t)i);
}

// flags -> per-block sums (1,024 strings per block); flag[i] = string i is the first of its class
__global__ void __launch_bounds__(1024) intern_flag_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ cmin, int64_t n,
                                                           int32_t* __restrict__ flag_scan, int32_t* __restrict__ block_sum)
{
    __shared__ int32_t warp_sum[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t f = (i < n && cmin[rep[i]] == (int32_t)i) ? 1 : 0;
    int32_t inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int32_t v = warp_sum[lane], s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        warp_sum[lane] = s - v;
        if (lane == 31) block_sum[blockIdx.x] = s;
    }
    __syncthreads();
    if (i < n) flag_scan[i] = warp_sum[w] + inc - f;   // exclusive, inside the block
}

// exclusive scan of the block sums by one block (n_blocks is small: n / 1,024)
__global__ void __launch_bounds__(1024) intern_block_scan_kernel(int32_t* block_sum, int n_blocks, int32_t* total)
{
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int32_t v = i < n_blocks ? block_sum[i] : 0;
        int32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int32_t x = warp_sum[lane], s = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t t = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += t;
            }
            warp_sum[lane] = s - x;
        }
        __syncthreads();
        const int32_t excl = carry + warp_sum[w] + inc - v;
        if (i < n_blocks) block_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) intern_ids_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ cmin,
                                                         const int32_t* __restrict__ flag_scan, const int32_t* __restrict__ block_sum,
                                                         int64_t n, int32_t* __restrict__ ids)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t first = cmin[rep[i]];
    ids[i] = block_sum[first >> 10] + flag_scan[first];
}

}  // namespace

#ifndef IADH_HOST_EMU   // (tests/emu compiles the kernels above for the host)
// device pointers throughout; table = table_size int32 (-1), cmin = n int32 (INT_MAX); both initialised by the caller
void launch_intern(const char* blob, const int64_t* off, int64_t n, int32_t* table, uint32_t table_size, int32_t* rep, int32_t* cmin,
                   int32_t* flag_scan, int32_t* block_sum, int32_t* total, int32_t* ids, cudaStream_t st)
{
    if (n == 0) return;
    const unsigned g256 = (unsigned)((n + 255) / 256), g1024 = (unsigned)((n + 1023) / 1024);
    intern_insert_kernel<<<g256, 256, 0, st>>>(blob, off, n, table, table_size - 1, rep);
    intern_min_kernel<<<g256, 256, 0, st>>>(rep, n, cmin);
    intern_flag_kernel<<<g1024, 1024, 0, st>>>(rep, cmin, n, flag_scan, block_sum);
    intern_block_scan_kernel<<<1, 1024, 0, st>>>(block_sum, (int)g1024, total);
    intern_ids_kernel<<<g256, 256, 0, st>>>(rep, cmin, flag_scan, block_sum, n, ids);
}
#endif  // IADH_HOST_EMU

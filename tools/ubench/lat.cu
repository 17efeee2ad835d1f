// Latency microbenchmarks that set the cost model of the clustering kernel (one busy thread):
// dependent global loads (L1 hit), read-after-write through L1, shared memory through generic
// and LDS paths, FP64 add / div / sqrt chains, __syncthreads with idle warps.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(int* buf, double* dbuf, long long* out, int iters)
{
    extern __shared__ int sm[];
    const int tid = threadIdx.x;
    long long t0, t1;
    // 0: dependent global load chain (pointer chase inside 4 KB: L1 hits)
    if (tid == 0) {
        for (int i = 0; i < 1024; i++) buf[i] = (i + 17) & 1023;
    }
    __syncthreads();
    if (tid == 0) {
        int p = 0;
        for (int i = 0; i < 1024; i++) p = buf[p];  // warm
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = buf[p];
        t1 = clock64();
        out[0] = (t1 - t0) / iters + (p & 0);
        // 1: store then dependent load of the same address (does the store keep the L1 line?)
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            buf[p] = (p + 17) & 1023;
            p = buf[p];
        }
        t1 = clock64();
        out[1] = (t1 - t0) / iters + (p & 0);
        // 2: volatile (bypass L1) chain = L2 latency
        volatile int* vb = buf;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = vb[p];
        t1 = clock64();
        out[2] = (t1 - t0) / iters + (p & 0);
        // 3: shared memory chain (LDS)
        for (int i = 0; i < 1024; i++) sm[i] = (i + 17) & 1023;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = sm[p];
        t1 = clock64();
        out[3] = (t1 - t0) / iters + (p & 0);
        // 4: shared memory through a generic pointer
        int* gp = (int*)((void*)sm);
        int* volatile hide = gp;
        int* q = hide;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = q[p];
        t1 = clock64();
        out[4] = (t1 - t0) / iters + (p & 0);
        // 5: dependent FP64 add chain
        double a = dbuf[0], b = dbuf[1];
        t0 = clock64();
        for (int i = 0; i < iters; i++) a = a + b;
        t1 = clock64();
        out[5] = (t1 - t0) / iters;
        dbuf[2] = a;
        // 6: dependent FP64 div chain
        a = dbuf[0];
        t0 = clock64();
        for (int i = 0; i < iters; i++) a = b / a;
        t1 = clock64();
        out[6] = (t1 - t0) / iters;
        dbuf[3] = a;
        // 7: dependent sqrt chain
        a = dbuf[0] + 5.0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) a = sqrt(a) + b;
        t1 = clock64();
        out[7] = (t1 - t0) / iters;
        dbuf[4] = a;
        // 8: dependent FP64 mul+add (no fma) chain
        a = dbuf[0];
        t0 = clock64();
        for (int i = 0; i < iters; i++) a = __dadd_rn(__dmul_rn(a, b), b);
        t1 = clock64();
        out[8] = (t1 - t0) / iters;
        dbuf[5] = a;
    }
    __syncthreads();
    // 9: __syncthreads round trip with all warps arriving at once
    t0 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0) / iters;
    // 10: another thread's store, then thread 0 reads (L1 coherence inside the CTA)
    __syncthreads();
    long long acc = 0;
    for (int i = 0; i < 64; i++) {
        if (tid == 32) buf[i] = i + 1;
        __syncthreads();
        if (tid == 0) {
            t0 = clock64();
            int v = buf[i];
            t1 = clock64() + (v & 0);
            acc += t1 - t0;
        }
        __syncthreads();
    }
    if (tid == 0) out[10] = acc / 64;
    // 11: same, but the line was read (cached) by thread 0 before the other thread's store
    acc = 0;
    for (int i = 0; i < 64; i++) {
        int pre = 0;
        if (tid == 0) pre = buf[512 + i * 0 + (i & 7)];
        __syncthreads();
        if (tid == 32) buf[512 + (i & 7)] = i + pre;
        __syncthreads();
        if (tid == 0) {
            t0 = clock64();
            int v = buf[512 + (i & 7)];
            t1 = clock64() + (v & 0);
            acc += t1 - t0;
        }
        __syncthreads();
    }
    if (tid == 0) out[11] = acc / 64;
}

int main()
{
    int* buf;
    double* dbuf;
    long long* out;
    cudaMalloc(&buf, 1 << 16);
    cudaMalloc(&dbuf, 64);
    cudaMallocManaged(&out, 16 * 8);
    double h[2] = {1.2345, 1.0000001};
    cudaMemcpy(dbuf, h, 16, cudaMemcpyHostToDevice);
    const char* names[] = {"LDG chain (L1 hit)", "ST+LD same addr", "LDG volatile (L2)", "LDS chain", "generic->smem chain", "DADD chain",
                           "DDIV chain", "DSQRT+DADD chain", "DMUL+DADD chain", "__syncthreads (NT threads)", "read after other-thread store (cold line)",
                           "read after other-thread store (line was in L1)"};
    for (int nt : {32, 128, 256}) {
        k<<<1, nt, 8192>>>(buf, dbuf, out, 2000);
        cudaError_t e = cudaDeviceSynchronize();
        printf("NT=%d (%s)\n", nt, cudaGetErrorString(e));
        for (int i = 0; i < 12; i++) printf("  %-48s %lld cycles\n", names[i], out[i]);
    }
    return 0;
}

// Isolated cost of pq.pop() for both heap layouts, one warp, by heap size.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o heap_pop heap_pop.cu
#include <cstdio>
#include <vector>
#include "../../p-dmpc_b200/csrc/pdmpc_kernels.cuh"
using namespace pdmpc;

constexpr int HS = 4096;
struct Sm {
    double hf[HS + 2];
    unsigned long long hw[HS];
    HEnt heap[HS];
};

__device__ double rnd(unsigned &s) {
    s = s * 1664525u + 1013904223u;
    return (double)(s >> 8) * (1.0 / 16777216.0);
}

template <int MODE>
__global__ void bench(int n0, int extra_warps_busy, long long *out, HEnt *gl) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sm &sm = *reinterpret_cast<Sm *>(raw);
    const int lane = threadIdx.x % 32;
    if (threadIdx.x >= 32) {   // optional neighbours hammering shared memory + FP64
        double acc = 0;
        volatile double *p = sm.hf;
        for (int it = 0; it < extra_warps_busy; ++it) acc = acc * 1.000001 + p[(it * 7 + lane) & 1023];
        if (acc == 123.456) out[100] = 1;
        return;
    }
    unsigned seed = 12345u;
    Tile<32> t;
    t.shift = 0; t.lane = lane; t.mask = 0xffffffffu;
    HeapSplit hs;
    hs.sf = shared_base_once(sm.hf); hs.sw = shared_base_once(sm.hw); hs.gl = gl; hs.hs = HS; hs.len = 0;
    Heap<HS, 32> ho;
    ho.sm = sm.heap; ho.gl = gl; ho.len = 0;
    for (int i = 0; i < n0; ++i) {
        HEnt e;
        e.f = rnd(seed); e.w = i;
        if (MODE == 0) hs.push_many(e, 1, lane); else ho.push_many(e, 1, t);
    }
    __syncwarp();
    long long tot = 0;
    unsigned long long chk = 0;
    const int npop = n0 / 2;
    // steady state: pop one, push one (keeps the size), like the search does
    for (int i = 0; i < npop; ++i) {
        long long t0 = clock64();
        HEnt top = (MODE == 0) ? hs.pop(lane) : ho.pop(t);
        long long t1 = clock64();
        tot += t1 - t0;
        chk += top.w;
        HEnt e;
        e.f = top.f + rnd(seed) * 0.3; e.w = i;
        if (MODE == 0) hs.push_many(e, 1, lane); else ho.push_many(e, 1, t);
    }
    if (lane == 0) { out[0] = tot / npop; out[1] = (long long)chk; }
}

int main() {
    long long *out;
    HEnt *gl;
    cudaMalloc(&out, 1024 * 8);
    cudaMalloc(&gl, sizeof(HEnt) * 65536);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
    for (int busy : {0, 200000}) {
        for (int n : {16, 64, 256, 1024, 4000}) {
            long long h[2][2];
            for (int mode = 0; mode < 2; ++mode) {
                int threads = busy ? 512 : 32;
                if (mode == 0) bench<0><<<1, threads, sizeof(Sm)>>>(n, busy, out, gl);
                else bench<1><<<1, threads, sizeof(Sm)>>>(n, busy, out, gl);
                cudaDeviceSynchronize();
                cudaMemcpy(h[mode], out, 16, cudaMemcpyDeviceToHost);
            }
            printf("heap %5d entries, %s: split walk %5lld cyc/pop   cooperative look-ahead %5lld cyc/pop   (checksums %lld %lld)\n", n,
                   busy ? "15 busy neighbour warps" : "alone", h[0][0], h[1][0], h[0][1], h[1][1]);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Speed-of-light probe for the step kernels' traffic pattern (EXPERIMENT, not product code):
// per env read 4 B state + 4 B action, write 4 x 4 B (next_state, obs, reward, flags); no compute.
// Tells how much of the measured copy peak this 1:2 read:write mix can reach at the bench's
// batch sizes, with and without programmatic dependent launch (PDL).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe scripts/exp_stream_probe.cu && /tmp/probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <bool PDL>
__global__ void __launch_bounds__(512, 2) probe(const int4* __restrict__ s, const int4* __restrict__ a, int4* __restrict__ o0,
                                                int4* __restrict__ o1, int4* __restrict__ o2, int4* __restrict__ o3, int64_t ng) {
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    const int64_t nt = (int64_t)gridDim.x * blockDim.x;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int4 cs = make_int4(0, 0, 0, 0), ca = cs;
    if (g < ng) { cs = __ldcs(s + g); ca = __ldcs(a + g); }
    while (g < ng) {
        const int64_t gn = g + nt;
        int4 ns = cs, na = ca;
        if (gn < ng) { ns = __ldcs(s + gn); na = __ldcs(a + gn); }
        int4 v = make_int4(cs.x ^ ca.x, cs.y ^ ca.y, cs.z ^ ca.z, cs.w ^ ca.w);
        __stcs(o0 + g, v); __stcs(o1 + g, ca); __stcs(o2 + g, cs); __stcs(o3 + g, v);
        cs = ns; ca = na; g = gn;
    }
}

// reset pattern: nothing read, two streams written (state + obs: 8 B per env), 256-thread CTAs like pomdp_reset_kernel
__global__ void __launch_bounds__(256) probe_reset(int4* __restrict__ o0, int4* __restrict__ o1, int64_t ng) {
    const int64_t nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += nt) {
        const int4 v = make_int4((int)g, (int)g + 1, (int)g + 2, (int)g + 3);
        __stcs(o0 + g, v); __stcs(o1 + g, v);
    }
}
static float run_reset(int64_t n, int sets, int iters, int4** bufs, cudaStream_t st, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int64_t ng = n / 4;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_reset, 256, 0);
    const int64_t need = (ng + 255) / 256, cap = (int64_t)sms * per_sm, trips = (need + cap - 1) / cap;
    const int grid = (int)(need <= cap ? need : (need + trips - 1) / trips);
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < iters; ++i) {
        int4** b = bufs + 6 * (i % sets);
        probe_reset<<<grid, 256, 0, st>>>(b[2], b[3], ng);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters * 1e3f;
}

template <bool PDL>
static float run(int64_t n, int sets, int iters, int4** bufs, cudaStream_t st, int grid, bool graph = false) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int64_t ng = n / 4;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    for (int rep = 0; rep < 2; ++rep) {
        if (rep == 1 && !graph) cudaEventRecord(e0, st);
        if (graph) {
            if (rep == 1) break;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        }
        for (int i = 0; i < iters; ++i) {
            int4** b = bufs + 6 * (i % sets);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
            cudaLaunchKernelEx(&cfg, probe<PDL>, (const int4*)b[0], (const int4*)b[1], b[2], b[3], b[4], b[5], ng);
        }
    }
    if (graph) {
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1; }
        e = cudaGraphInstantiate(&ge, g, 0);
        if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
        size_t ne = 0;
        cudaGraphGetEdges(g, nullptr, nullptr, &ne);
        cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        cudaGraphLaunch(ge, st);
    }
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters * 1e3f;
}

int main() {
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int64_t sizes[] = {1 << 20, 1 << 22, 1 << 24, 1 << 26};
    for (int64_t n : sizes) {
        const int sets = n == (1 << 20) ? 24 : n == (1 << 22) ? 6 : 2;      // rotating sets: > 126 MB of L2 in every case
        std::vector<int4*> bufs(6 * sets);
        for (auto& p : bufs) { CK(cudaMalloc(&p, n * 4)); CK(cudaMemset(p, 1, n * 4)); }
        // the product's grid: at most 2 CTAs of 512 per SM, balanced so that every thread runs the same number of trips
        const int64_t need = (n / 4 + 511) / 512, cap = (int64_t)sms * 2;
        const int64_t trips = (need + cap - 1) / cap;
        const int grid = (int)(need <= cap ? need : (need + trips - 1) / trips);
        const int iters = n <= (1 << 22) ? 2000 : 400;
        const float t0 = run<false>(n, sets, iters, bufs.data(), st, grid);
        const float t1 = run<true>(n, sets, iters, bufs.data(), st, grid);
        const float t2 = run<false>(n, sets, iters, bufs.data(), st, grid, true);
        const float t3 = run<true>(n, sets, iters, bufs.data(), st, grid, true);
        printf("n=2^%d  24 B/env = %.1f MB/launch: plain %.2f us (%.0f GB/s)   PDL %.2f us (%.0f GB/s)   graph %.2f us   graph+PDL %.2f us\n",
               (int)__builtin_ctzll(n), n * 24 / 1e6, t0, n * 24.0 / t0 / 1e3, t1, n * 24.0 / t1 / 1e3, t2, t3);
        const float t4 = run_reset(n, sets, iters, bufs.data(), st, sms);
        printf("n=2^%d  reset pattern, 8 B/env written = %.1f MB/launch: graph %.2f us (%.0f GB/s)\n", (int)__builtin_ctzll(n), n * 8 / 1e6, t4,
               n * 8.0 / t4 / 1e3);
        for (auto& p : bufs) cudaFree(p);
    }
    return 0;
}

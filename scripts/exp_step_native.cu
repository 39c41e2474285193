// EXPERIMENT (not product code): time pomdp_rock_step through the C ABI without Python/torch --
// eager stream launches vs a natively captured CUDA graph, with PDL on/off (POMDP_B200_NO_PDL=1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/stepn scripts/exp_step_native.cu && /tmp/stepn
#include "../gym_pomdp_b200/csrc/pomdp_kernels.cu"

#include <stdio.h>
#include <vector>

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : (1 << 22);
    const int sets = 6, iters = 2000;
    PomdpRockParams q = {11, 11, 0, 0, 0.8};
    std::vector<char> tbl(pomdp_rock_table_bytes(&q));
    pomdp_rock_build_table(&q, tbl.data());
    void* d_tbl;
    cudaMalloc(&d_tbl, tbl.size());
    cudaMemcpy(d_tbl, tbl.data(), tbl.size(), cudaMemcpyHostToDevice);
    std::vector<int32_t*> b(6 * sets);
    std::vector<int32_t> h(n);
    for (int i = 0; i < 6 * sets; ++i) {
        cudaMalloc(&b[i], n * 4);
        for (int64_t j = 0; j < n; ++j) {
            uint32_t r = (uint32_t)(j * 2654435761u + i * 40503u);
            if (i % 6 == 0) h[j] = (int32_t)((r % 11) | (((r >> 8) % 11) << 4) | ((r >> 3) & 0x15555500));   // x, y, statuses in {0,1}
            else h[j] = (int32_t)((r >> 5) % 16);
        }
        cudaMemcpy(b[i], h.data(), n * 4, cudaMemcpyHostToDevice);
    }
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto step = [&](int i) {
        int32_t** s = b.data() + 6 * (i % sets);
        int rc = pomdp_rock_step(&q, d_tbl, s[0], s[1], s[2], s[3], (float*)s[4], s[5], n, 0, 0x5EED, (uint32_t)i, st);
        if (rc) { printf("step rc=%d %s\n", rc, pomdp_last_error()); exit(1); }
    };
    for (int i = 0; i < 20; ++i) step(i);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) step(i);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("n=%lld eager: %.2f us/launch\n", (long long)n, ms / iters * 1e3);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < iters; ++i) step(i);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess) { printf("capture: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return 1; }
    cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("n=%lld graph: %.2f us/launch (%.0f GB/s algorithmic)\n", (long long)n, ms / iters * 1e3, n * 24.0 / (ms / iters * 1e-3) / 1e9);
    return 0;
}

// kbench_len: times the long-sequence kernels (L = 256, 512) in their two FIFO placements --
// shared memory (sw_kernel) and global memory / L2 (sw_kernel_gfifo, persistent grid) -- over
// block shapes.  Development tool, not part of the product path.  Every variant's scores are
// compared word for word with the shared-memory kernel's (which the GPU parity tests pin to
// the oracle).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I../smith-waterman-simd_b200/csrc \
//        -o kbench_len kbench_len.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "sw_kernel.cuh"
#include "sw_params.h"

using namespace swb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static uint8_t *d1, *d2; static int32_t *dsc, *dref;
static uint32_t* dscratch; static size_t scratch_bytes;
static int n_sm;

static uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

template <int L>
static uint64_t n_pairs() { return 454656; }   // = 148 SMs x 3072: WHOLE waves of every shape compared here -- 16 of the 3-warp, 8 of the 6-warp, 6 of the
                                               // 8-warp and 4 of the 12-warp configurations (resident pairs = SMs x warps x 32 x 2).  A pair is one indivisible item, so
                                               // a partly filled last wave would be charged to the shape, not the batch: the first comparison (262144 pairs,
                                               // profiles/r01/kbench_len_*.jsonl) gave the shapes 9.2 / 4.6 / 3.5 / 2.3 waves and so under-read the 12-warp one most.

static bool same(uint64_t n)
{
    std::vector<int32_t> a(n), b(n);
    CK(cudaMemcpy(a.data(), dsc, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), dref, n * 4, cudaMemcpyDeviceToHost));
    return memcmp(a.data(), b.data(), n * 4) == 0;
}

template <class F>
static float time_launches(F&& launch)
{
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <bool FAST, int L, int NT, int MINB>
static void run_smem(const SwParams& prm, bool is_ref)
{
    const uint64_t N = n_pairs<L>();
    auto kern = sw_kernel<FAST, L, NT, MINB>;
    const size_t smem = sw_smem_bytes<L, NT>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    const unsigned grid = (unsigned)(((N + 1) / 2 + NT - 1) / NT);
    int32_t* out = is_ref ? dref : dsc;
    CK(cudaMemset(out, 0xff, N * 4));
    const float ms = time_launches([&] { kern<<<grid, NT, smem>>>(d1, d2, out, N, prm, (unsigned)L); });
    printf("{\"fifo\": \"smem\", \"fast\": %d, \"L\": %d, \"nt\": %d, \"minb\": %d, \"regs\": %d, \"occ_blocks\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"gcups\": %.1f, \"ok\": %s}\n",
           (int)FAST, L, NT, MINB, fa.numRegs, occ, occ * NT / 32, ms, N * (double)L * L / (ms * 1e-3) / 1e9, is_ref ? "\"ref\"" : (same(N) ? "true" : "false"));
    fflush(stdout);
}

template <bool FAST, int L, int NT, int MINB, int AHEAD = 1>
static void run_gfifo(const SwParams& prm)
{
    const uint64_t N = n_pairs<L>();
    auto kern = sw_kernel_gfifo<FAST, L, NT, MINB, SW_DEFAULT_VARIANT, AHEAD>;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, 0));
    if (occ > MINB) occ = MINB;
    const unsigned grid = (unsigned)(n_sm * occ);
    const size_t need = (size_t)grid * L * NT * 4;
    if (need > scratch_bytes) { printf("{\"fifo\": \"global\", \"L\": %d, \"nt\": %d, \"minb\": %d, \"skipped\": \"scratch\"}\n", L, NT, MINB); return; }
    CK(cudaMemset(dsc, 0xff, N * 4));
    const float ms = time_launches([&] { kern<<<grid, NT>>>(d1, d2, dsc, N, prm, (unsigned)L, dscratch); });
    printf("{\"fifo\": \"global\", \"ahead\": %d, \"fast\": %d, \"L\": %d, \"nt\": %d, \"minb\": %d, \"regs\": %d, \"occ_blocks\": %d, \"warps_per_sm\": %d, \"scratch_mb\": %.1f, \"ms\": %.4f, \"gcups\": %.1f, \"ok\": %s}\n",
           AHEAD, (int)FAST, L, NT, MINB, fa.numRegs, occ, occ * NT / 32, need / 1048576.0, ms, N * (double)L * L / (ms * 1e-3) / 1e9, same(N) ? "true" : "false");
    fflush(stdout);
}

template <int L>
static void fill_inputs()
{
    const uint64_t N = n_pairs<L>();
    std::vector<uint8_t> a(N * L), b(N * L);
    uint64_t s = 12345 + L;
    // related pairs (b = a with 10 % substitutions) so that scores are long alignments, not noise
    for (uint64_t i = 0; i < N * L; i += 16) {
        uint64_t r = splitmix(s), m = splitmix(s);
        for (int k = 0; k < 16; ++k) {
            a[i + k] = (r >> (2 * k)) & 3;
            const bool mut = ((m >> (4 * k)) & 15) < 2;
            b[i + k] = mut ? ((r >> (32 + 2 * k)) & 3) : a[i + k];
        }
    }
    CK(cudaMemcpy(d1, a.data(), N * L, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, b.data(), N * L, cudaMemcpyHostToDevice));
}

static bool g_quick = false;   // `kbench_len ncu`: the L = 512 pair of kernels only, for a profiler capture

int main(int argc, char** argv)
{
    g_quick = argc > 1;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    n_sm = prop.multiProcessorCount;
    const uint64_t maxN = n_pairs<512>();
    CK(cudaMalloc(&d1, maxN * 512)); CK(cudaMalloc(&d2, maxN * 512));
    CK(cudaMalloc(&dsc, maxN * 4)); CK(cudaMalloc(&dref, maxN * 4));
    scratch_bytes = (size_t)n_sm * 12 * 512 * 32 * 4 * 2;   // up to 24 warps/SM at L = 512
    CK(cudaMalloc(&dscratch, scratch_bytes));
    const int8_t sm[16] = {10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10};

    {
        const SwParams fast = sw_make_params(sm, 15, 0, 512), gen = sw_make_params(sm, 15, 1, 512);
        fill_inputs<512>();
        run_smem<true, 512, 32, 3>(fast, true);
        if (g_quick) { run_gfifo<true, 512, 64, 4>(fast); return 0; }
        run_gfifo<true, 512, 32, 4>(fast);
        run_gfifo<true, 512, 64, 4>(fast);
        run_gfifo<true, 512, 64, 4, 2>(fast);
        run_gfifo<true, 512, 128, 2, 2>(fast);
        run_gfifo<true, 512, 64, 5>(fast);
        run_gfifo<true, 512, 64, 5, 2>(fast);
        run_gfifo<true, 512, 64, 6>(fast);
        run_gfifo<true, 512, 64, 6, 2>(fast);
        run_smem<false, 512, 32, 3>(gen, false);
        run_gfifo<false, 512, 64, 4>(gen);
        run_gfifo<false, 512, 64, 4, 2>(gen);
        run_gfifo<false, 512, 64, 6>(gen);
    }
    {
        const SwParams fast = sw_make_params(sm, 15, 0, 256), gen = sw_make_params(sm, 15, 1, 256);
        fill_inputs<256>();
        run_smem<true, 256, 64, 3>(fast, true);
        run_gfifo<true, 256, 64, 4>(fast);
        run_gfifo<true, 256, 64, 4, 2>(fast);
        run_gfifo<true, 256, 64, 6>(fast);
        run_gfifo<true, 256, 64, 6, 2>(fast);
    }
    return 0;
}

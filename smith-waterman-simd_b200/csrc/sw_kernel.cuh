// sw_kernel.cuh -- the sm_100a kernel around sw_core.cuh.
//
// Work decomposition: thread t of the grid scores pairs 2t (low halves) and 2t+1 (high
// halves) completely; there is no inter-thread communication.  A block of NT threads owns
// NT*128 words of shared memory: thread-interleaved FIFOs (word c of thread x at
// [c*NT + x], so a warp's access is one conflict-free 128-byte wavefront) that carry each
// strip's bottom row to the next strip, plus the 4-word profile table.
//
// HBM layout read by the kernel: seq1 and seq2 as the reference passes them
// (std::array<uint8_t,128>, source.cpp:36-37) laid end to end: [n][128] bytes, codes 0..3.
// Each thread pulls 16 bases of each target per 16 steps with two 8-byte loads, one
// iteration ahead; 260 bytes of HBM traffic per pair against 16 384 cell updates.
#pragma once
#include <cuda_runtime.h>
#include "sw_core.cuh"

namespace swb {

constexpr int SW_DEFAULT_VARIANT = SW_V_BEST_FMA;   // the variant bits (sw_core.cuh) the shipped kernels use

template <int NT>
struct SmemFifo {
    static constexpr int kPrefetch = 0;        // 29-cycle LDS: read at the step of use
    uint32_t* f;   // &fifo[0*NT + threadIdx.x]
    __device__ __forceinline__ uint32_t pop(int c) const { return f[c * NT]; }
    __device__ __forceinline__ void push(int c, uint32_t v) { f[c * NT] = v; }
};

// The same FIFO in GLOBAL memory, for sequence lengths whose L-word FIFO would leave too few
// resident threads if it lived in shared memory (L = 512: 3 warps per SM).  One slot of L*NT
// words per resident block of a persistent grid, thread-interleaved (a warp's access is one
// 128-byte line); the working set (resident threads x L x 4 B, 58 MB at L = 256, 116 MB at
// L = 512) lives in the 126 MB L2.  A thread only ever reads what it wrote itself, so plain
// (coherent) loads and stores are ordered correctly without fences; .cg keeps them out of L1.
template <int NT, int AHEAD = 1>
struct GlobalFifo {
    static constexpr int kPrefetch = AHEAD;    // L2 latency: read 8..15 (1) or 16..31 (2) steps ahead (sw_core.cuh)
    uint32_t* f;   // &slot[0*NT + threadIdx.x]
    __device__ __forceinline__ uint32_t pop(int c) const { return __ldcg(f + c * NT); }
    __device__ __forceinline__ void push(int c, uint32_t v) { __stcg(f + c * NT, v); }
};

struct SmemTable {
    const uint32_t* t;
    __device__ __forceinline__ uint32_t operator()(uint32_t byte_off) const
    {
        return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(t) + byte_off);
    }
};

template <int L, int NT>
constexpr size_t sw_smem_bytes() { return (size_t)(L * NT + 4) * sizeof(uint32_t); }

template <int NT>
constexpr size_t sw128_smem_bytes() { return sw_smem_bytes<SW_L, NT>(); }

// Sequence length L is a template parameter (128 = the reference's shape; 256 and 512 for the
// length sweep).  The per-thread FIFO is L words of shared memory, so NT shrinks as L grows.
template <bool FAST, int L, int NT, int MINB, int V = SW_DEFAULT_VARIANT>
__global__ void __launch_bounds__(NT, MINB)
sw_kernel(const uint8_t* __restrict__ seq1, const uint8_t* __restrict__ seq2,
             int32_t* __restrict__ scores, unsigned long long n, const SwParams prm, const unsigned seq2_stride)
{
    // seq2_stride = L: pair p's target is seq2[p][L];  0: one target shared by all pairs
    // (the batched shape of SmithWaterman_8b111x32mark1, source.cpp:1227-1234).
    extern __shared__ uint32_t smem[];
    uint32_t* t4s = smem + L * NT;
    if (threadIdx.x < 4) t4s[threadIdx.x] = prm.t4[threadIdx.x];
    __syncthreads();

    const unsigned long long p = 2ull * ((unsigned long long)blockIdx.x * NT + threadIdx.x);
    if (p >= n) return;
    const unsigned long long q = (p + 1 < n) ? p + 1 : p;   // odd tail: the high half repeats the low pair

    SmemFifo<NT> fifo{smem + threadIdx.x};
    SmemTable t4{t4s};
    int32_t lo, hi;
    sw_two_pairs<FAST, L, V>(seq1 + p * L, seq2 + p * seq2_stride, (q != p) ? (uint32_t)L : 0u, (q != p) ? seq2_stride : 0u,
                             fifo, t4, prm, lo, hi);
    if (q != p) {
        *reinterpret_cast<int2*>(scores + p) = make_int2(lo, hi);   // p is even: 8-byte aligned
    } else {
        scores[p] = lo;
    }
}

// Persistent-grid variant with the FIFO in global memory: grid = resident blocks; block b owns
// slot b of `fifo_scratch` (L*NT words) and walks the work items b, b+gridDim.x, ...
template <bool FAST, int L, int NT, int MINB, int V = SW_DEFAULT_VARIANT, int AHEAD = 1>
__global__ void __launch_bounds__(NT, MINB)
sw_kernel_gfifo(const uint8_t* __restrict__ seq1, const uint8_t* __restrict__ seq2,
                int32_t* __restrict__ scores, unsigned long long n, const SwParams prm, const unsigned seq2_stride,
                uint32_t* __restrict__ fifo_scratch)
{
    __shared__ uint32_t t4s[4];
    if (threadIdx.x < 4) t4s[threadIdx.x] = prm.t4[threadIdx.x];
    __syncthreads();
    GlobalFifo<NT, AHEAD> fifo{fifo_scratch + (size_t)blockIdx.x * L * NT + threadIdx.x};
    SmemTable t4{t4s};
    const unsigned long long n_items = (n + 1) / 2;
    for (unsigned long long item = (unsigned long long)blockIdx.x * NT + threadIdx.x; item < n_items;
         item += (unsigned long long)gridDim.x * NT) {
        const unsigned long long p = 2ull * item;
        const unsigned long long q = (p + 1 < n) ? p + 1 : p;
        int32_t lo, hi;
        sw_two_pairs<FAST, L, V>(seq1 + p * L, seq2 + p * seq2_stride, (q != p) ? (uint32_t)L : 0u, (q != p) ? seq2_stride : 0u,
                                 fifo, t4, prm, lo, hi);
        if (q != p) *reinterpret_cast<int2*>(scores + p) = make_int2(lo, hi);
        else scores[p] = lo;
    }
}

} // namespace swb

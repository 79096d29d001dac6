// swb200_api.cu -- the C ABI of include/swb200.h: context, batch packer (pinned staging,
// chunked H2D / kernel / D2H overlap), index-range sharding over the GPUs of one box with
// a plain host-side gather (no collective: pairs are independent), and the
// kernel launches.  There is no CPU scoring path in this file or anywhere in the product.
#include "../../include/swb200.h"

#include <cuda_runtime.h>

#include <sched.h>

#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "sg_kernel.cuh"
#include "sw_feed_kernel.cuh"
#include "sw_pair_kernel.cuh"
#include "sw_kernel.cuh"
#include "sw_params.h"

namespace swb {                  // hostpack.cpp
void pack2bit_host(const uint8_t* codes, uint8_t* packed, size_t n_codes);
void unpack2bit_host(const uint8_t* packed, uint8_t* codes, size_t n_codes);
}

namespace {

using namespace swb;

constexpr int kSlots = 3;        // chunks in flight per GPU (chunk pipeline, run_range)
// Chunk pipeline: a chunk is two whole waves of the L = 128 kernel's resident pairs (148 SMs x 6 blocks x 64 threads x 2
// pairs = 113 664 on a B200) and the same number of BYTES at every length, so no launch ends in a mostly empty wave
// (round 1's fixed 131 072 pairs were 1.15 waves: 17 % of every chunk's kernel time was tail).
constexpr uint64_t kChunkWaves = 2;

thread_local std::string g_init_error;
thread_local std::string g_last_error_copy;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t* d_seq1 = nullptr;       // [chunk_pairs][128]
    uint8_t* d_seq2 = nullptr;
    uint8_t* d_pk1 = nullptr;        // [chunk_pairs][32]  2-bit packed staging
    uint8_t* d_pk2 = nullptr;
    int32_t* d_scores = nullptr;     // [chunk_pairs]
    bool busy = false;
};

struct FeedState;                    // feed.inc: the persistent-kernel batch packer
struct PairSlot;                     // pairpath.inc: the per-pair call's mapped slots
struct SgPipe;                       // sg_pipe.inc: the semi-global aligner's compressed-wire host pipeline

struct Device {
    int id = -1;
    cudaDeviceProp prop{};
    Slot slots[kSlots];
    std::mutex mu;                   // one host batch at a time per GPU
    unsigned long long* d_bad = nullptr;
    // semi-global aligner: round-record scratch, and four staging slots for host batches
    struct SgScratch {
        uint4* traces = nullptr;         // round records, one row per pair of a launch + one spare row
        size_t trace_bytes = 0;
    } sg_dev;                            // scratch of the device-resident entry
    struct SgSlot {
        SgScratch scratch;               // each staging slot has its own, so the two slots' kernels are independent
        cudaStream_t stream = nullptr;
        uint8_t *d_seq1 = nullptr, *d_seq2 = nullptr, *d_ops = nullptr;
        int32_t* d_meta = nullptr;       // [4][cap]: score, end_y, end_x, n_ops
        size_t cap_pairs = 0; int len = 0; bool with_ops = false;
    } sg_slots[4];
    uint64_t chunk_pairs = 0;        // chunk pipeline: pairs of 128 bases per chunk (whole waves), set by setup_device
    FeedState* feed = nullptr;       // persistent-kernel batch packer (feed.inc)
    PairSlot* pair = nullptr;        // per-pair path (pairpath.inc), created on first use
    SgPipe* sg_pipe = nullptr;       // semi-global host pipeline (sg_pipe.inc), created on first use
};

} // namespace

struct swb200_ctx {
    std::vector<Device*> devs;
    std::string err;
    std::mutex err_mu;               // per-GPU threads, lanes and submit threads can all fail at once
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> packed_pairs{0}, raw_pairs{0};   // host batches: pairs sent 2-bit packed / as bytes
    int force_general = 0;
    int latency_path = 1;            // small batches of 128-mers take the one-warp-per-pair kernel (pairpath.inc); test hook
    int pair_doorbell = 1;           // a single pair goes to the resident server kernel through its doorbell (pairpath.inc); test hook
    int pack_threads = -1;           // per GPU; -1 = auto (host cores available to this process), 0 = off
    std::mutex tickets_mu;
    std::map<uint64_t, std::pair<std::thread, int*>> tickets;
    uint64_t next_ticket = 1;
};

namespace {

int fail(swb200_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = buf; }
    else g_init_error = buf;
    return code;
}

#define SWB_CUDA(ctx, call)                                                        \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return fail((ctx), SWB200_ERR_CUDA, #call, e_);     \
    } while (0)

// 2-bit packed -> byte codes, the reference's `unpack` (source.cpp:1580-1583):
// dest[i*4+j] = (src[i] >> 2j) & 3.  One thread expands 4 packed bytes into 16 codes.
__global__ void unpack2bit_kernel(const uint32_t* __restrict__ src, uint4* __restrict__ dst, unsigned long long n_words)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = __ldg(src + i);
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const uint32_t x = (w >> (8 * b)) & 0xffu;
        o[b] = (x & 3u) | (((x >> 2) & 3u) << 8) | (((x >> 4) & 3u) << 16) | (((x >> 6) & 3u) << 24);
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void count_bad_codes_kernel(const uint8_t* __restrict__ codes, unsigned long long n, unsigned long long* n_bad)
{
    unsigned long long local = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        local += (codes[i] > 3u);
    local = __reduce_add_sync(0xffffffffu, (unsigned)local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_bad, local);
}

// Launch geometry per sequence length.  With the FIFO in shared memory (L words per thread,
// 192 KiB of FIFO per SM in every case) the resident thread count shrinks as L grows: 12 warps
// per SM at L = 128, 6 at 256, 3 at 512 -- and 3 warps leave one of the four schedulers idle.
// L = 512 therefore keeps its FIFO in global memory (GlobalFifo, sw_kernel.cuh): a persistent
// grid of SMs x MINB blocks, one L*NT-word slot per block, resident in L2.
// L = 128: 6 blocks x 64 threads measured best of the shapes in profiles/r01/kbench_*.jsonl;
// the long shapes: profiles/r01/kbench_len.jsonl.
template <int L> struct LenCfg;
template <> struct LenCfg<128> { static constexpr int NT = 64,  MINB = 6; static constexpr bool GFIFO = false; };
template <> struct LenCfg<256> { static constexpr int NT = 64,  MINB = 3; static constexpr bool GFIFO = false; };
template <> struct LenCfg<512> { static constexpr int NT = 64,  MINB = 6; static constexpr bool GFIFO = true; };   // 8080 vs 6439 GCUPS in shared memory

template <bool FAST, int L>
constexpr auto kernel_ptr()
{
    if constexpr (LenCfg<L>::GFIFO) return sw_kernel_gfifo<FAST, L, LenCfg<L>::NT, LenCfg<L>::MINB>;
    else return sw_kernel<FAST, L, LenCfg<L>::NT, LenCfg<L>::MINB>;
}

template <int L>
constexpr size_t kernel_smem() { return LenCfg<L>::GFIFO ? 0 : sw_smem_bytes<L, LenCfg<L>::NT>(); }

template <bool FAST, int L>
cudaError_t launch_sw(const uint8_t* d1, const uint8_t* d2, int32_t* dsc, uint64_t n, const SwParams& prm, cudaStream_t st, bool shared_target)
{
    if (n == 0) return cudaSuccess;
    constexpr int NT = LenCfg<L>::NT;
    const uint64_t threads = (n + 1) / 2;
    const uint64_t blocks = (threads + NT - 1) / NT;
    const unsigned stride = shared_target ? 0u : (unsigned)L;
    if constexpr (LenCfg<L>::GFIFO) {
        // persistent grid; the FIFO slots come from the stream-ordered pool (kept warm by
        // setup_device's release threshold), so concurrent streams never share a slot
        int dev = 0, n_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        const uint64_t resident = (uint64_t)n_sm * LenCfg<L>::MINB;
        const unsigned grid = (unsigned)(blocks < resident ? blocks : resident);
        uint32_t* fifo = nullptr;
        e = cudaMallocAsync(&fifo, (size_t)grid * L * NT * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        sw_kernel_gfifo<FAST, L, NT, LenCfg<L>::MINB><<<grid, NT, 0, st>>>(d1, d2, dsc, n, prm, stride, fifo);
        e = cudaGetLastError();
        const cudaError_t e2 = cudaFreeAsync(fifo, st);
        return e != cudaSuccess ? e : e2;
    } else {
        sw_kernel<FAST, L, NT, LenCfg<L>::MINB><<<(unsigned)blocks, NT, sw_smem_bytes<L, NT>(), st>>>(d1, d2, dsc, n, prm, stride);
        return cudaGetLastError();
    }
}

cudaError_t launch_for(const SwParams& prm, int L, const uint8_t* d1, const uint8_t* d2, int32_t* dsc, uint64_t n, cudaStream_t st,
                       bool shared_target = false)
{
    const bool sh = shared_target;
    switch (L) {
    case 128: return prm.fast ? launch_sw<true, 128>(d1, d2, dsc, n, prm, st, sh) : launch_sw<false, 128>(d1, d2, dsc, n, prm, st, sh);
    case 256: return prm.fast ? launch_sw<true, 256>(d1, d2, dsc, n, prm, st, sh) : launch_sw<false, 256>(d1, d2, dsc, n, prm, st, sh);
    case 512: return prm.fast ? launch_sw<true, 512>(d1, d2, dsc, n, prm, st, sh) : launch_sw<false, 512>(d1, d2, dsc, n, prm, st, sh);
    default:  return cudaErrorInvalidValue;
    }
}

template <bool FAST, int L>
cudaError_t prepare_kernel()
{
    if constexpr (LenCfg<L>::GFIFO) return cudaSuccess;
    else {
        auto kern = kernel_ptr<FAST, L>();
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kernel_smem<L>());
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
}

template <bool FAST, int L>
cudaError_t kernel_resources(cudaFuncAttributes* fa, int* blocks, int* nt, int* smem)
{
    constexpr int NT = LenCfg<L>::NT;
    auto kern = kernel_ptr<FAST, L>();
    cudaError_t e = cudaFuncGetAttributes(fa, kern);
    if (e != cudaSuccess) return e;
    *nt = NT;
    *smem = (int)kernel_smem<L>();
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, NT, kernel_smem<L>());
    if (LenCfg<L>::GFIFO && *blocks > LenCfg<L>::MINB) *blocks = LenCfg<L>::MINB;   // the persistent grid launches MINB per SM
    return e;
}

cudaError_t launch_unpack(const uint8_t* d_packed, uint8_t* d_codes, uint64_t n_seqs, cudaStream_t st)
{
    if (n_seqs == 0) return cudaSuccess;
    const unsigned long long n_words = n_seqs * 8ull;   // 32 packed bytes = 8 words per sequence
    const unsigned grid = (unsigned)((n_words + 255) / 256);
    unpack2bit_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_packed), reinterpret_cast<uint4*>(d_codes), n_words);
    return cudaGetLastError();
}

FeedState* feed_create();        // feed.inc
cudaError_t sg_pipe_prepare_kernels(int carveout);   // sg_pipe.inc

int setup_device(swb200_ctx* ctx, Device* d)
{
    SWB_CUDA(ctx, cudaSetDevice(d->id));
    SWB_CUDA(ctx, cudaGetDeviceProperties(&d->prop, d->id));
    SWB_CUDA(ctx, (prepare_kernel<true, 128>()));  SWB_CUDA(ctx, (prepare_kernel<false, 128>()));
    SWB_CUDA(ctx, (prepare_kernel<true, 256>()));  SWB_CUDA(ctx, (prepare_kernel<false, 256>()));
    SWB_CUDA(ctx, (prepare_kernel<true, 512>()));  SWB_CUDA(ctx, (prepare_kernel<false, 512>()));
    SWB_CUDA(ctx, cudaFuncSetAttribute(sw_feed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feed_smem_bytes()));
    SWB_CUDA(ctx, cudaFuncSetAttribute(sw_feed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feed_smem_bytes()));
    SWB_CUDA(ctx, cudaFuncSetAttribute(sw_feed_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SWB_CUDA(ctx, cudaFuncSetAttribute(sw_feed_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    d->chunk_pairs = kChunkWaves * (uint64_t)d->prop.multiProcessorCount * LenCfg<128>::MINB * LenCfg<128>::NT * 2;
    d->feed = feed_create();
    SWB_CUDA(ctx, cudaMalloc(&d->d_bad, sizeof(unsigned long long)));
    SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_TB_SMEM)));
    SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_TB_SMEM)));
    SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_TB_SMEM)));
    SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_TB_SMEM)));
    // ONE shared-memory carve-out for every kernel of the semi-global aligner.  An SM changes its L1 / shared-memory split
    // only when it is empty: with the default (the forward kernel uses no shared memory, the traceback 32 KiB per block) a
    // traceback launched beside the forward kernels of other chunks waited until every SM had drained -- measured in the
    // host pipeline (sg_pipe.inc): a 3 ms traceback took 16 ms, ending with the last forward kernel.
    {
        const int carve = cudaSharedmemCarveoutMaxShared;
        SWB_CUDA(ctx, cudaFuncSetAttribute(sg2_xdrop_kernel<true, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, cudaFuncSetAttribute(sg2_xdrop_kernel<false, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, cudaFuncSetAttribute(sg2_xdrop_kernel<true, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, cudaFuncSetAttribute(sg2_xdrop_kernel<false, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<8, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)));
        SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)));
        SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<16, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)));
        SWB_CUDA(ctx, (cudaFuncSetAttribute(sg_traceback_kernel<16, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)));
        SWB_CUDA(ctx, cudaFuncSetAttribute(sg_left_align_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, cudaFuncSetAttribute(unpack2bit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        SWB_CUDA(ctx, sg_pipe_prepare_kernels(carve));
    }
    {
        // keep freed stream-ordered allocations (the L = 512 FIFO slots) in the pool across syncs
        cudaMemPool_t pool;
        SWB_CUDA(ctx, cudaDeviceGetDefaultMemPool(&pool, d->id));
        uint64_t keep = ~0ull;
        SWB_CUDA(ctx, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (Slot& s : d->slots) {
        SWB_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        SWB_CUDA(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    return SWB200_OK;
}

// Device staging is allocated on first host-batch use, so a context that only ever
// launches on caller-owned device arrays holds no large buffers.
int ensure_staging(swb200_ctx* ctx, Device* d, bool packed)
{
    const uint64_t cp = d->chunk_pairs;
    for (Slot& s : d->slots) {
        if (!s.d_seq1) {
            SWB_CUDA(ctx, cudaMalloc(&s.d_seq1, cp * SWB200_SEQ_LEN));
            SWB_CUDA(ctx, cudaMalloc(&s.d_seq2, cp * SWB200_SEQ_LEN));
            SWB_CUDA(ctx, cudaMalloc(&s.d_scores, cp * sizeof(int32_t)));
        }
        if (packed && !s.d_pk1) {
            SWB_CUDA(ctx, cudaMalloc(&s.d_pk1, cp * 32));
            SWB_CUDA(ctx, cudaMalloc(&s.d_pk2, cp * 32));
        }
    }
    return SWB200_OK;
}

// Whatever was enqueued on the chunk pipeline's slots has finished (errors ignored: this is the error path's drain, so
// that no copy into the caller's arrays is still in flight when an error code is returned).
void drain_slots(Device* d)
{
    for (Slot& s : d->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        s.busy = false;
    }
}

// Chunk pipeline -- one GPU's share [lo, hi) of a host batch the persistent-kernel packer (feed.inc) does not take:
// small batches, the length sweep's L = 256 / 512, and the one-vs-many shape.  Chunks of d->chunk_pairs * 128 / L pairs
// cycle through kSlots streams, so chunk c+1's H2D and chunk c-1's D2H run under chunk c's kernel.
int run_range_locked(swb200_ctx* ctx, Device* d, const uint8_t* seq1, const uint8_t* seq2, bool packed, int L,
                     const SwParams& prm, int32_t* scores, uint64_t lo, uint64_t hi, bool shared_target)
{
    SWB_CUDA(ctx, cudaSetDevice(d->id));
    int rc = ensure_staging(ctx, d, packed);
    if (rc != SWB200_OK) return rc;
    const size_t in_stride = packed ? 32 : (size_t)L;
    const uint64_t chunk = d->chunk_pairs * SWB200_SEQ_LEN / (uint64_t)L;    // same bytes per chunk at every length
    int si = 0;
    for (uint64_t c0 = lo; c0 < hi; c0 += chunk, si = (si + 1) % kSlots) {
        Slot& s = d->slots[si];
        const uint64_t m = (hi - c0 < chunk) ? hi - c0 : chunk;
        if (s.busy) { s.busy = false; SWB_CUDA(ctx, cudaEventSynchronize(s.done)); }
        uint8_t* in1 = packed ? s.d_pk1 : s.d_seq1;
        uint8_t* in2 = packed ? s.d_pk2 : s.d_seq2;
        SWB_CUDA(ctx, cudaMemcpyAsync(in1, seq1 + c0 * in_stride, m * in_stride, cudaMemcpyHostToDevice, s.stream));
        if (shared_target) SWB_CUDA(ctx, cudaMemcpyAsync(in2, seq2, in_stride, cudaMemcpyHostToDevice, s.stream));   // one target for all pairs
        else               SWB_CUDA(ctx, cudaMemcpyAsync(in2, seq2 + c0 * in_stride, m * in_stride, cudaMemcpyHostToDevice, s.stream));
        if (packed) {
            SWB_CUDA(ctx, launch_unpack(s.d_pk1, s.d_seq1, m, s.stream));
            SWB_CUDA(ctx, launch_unpack(s.d_pk2, s.d_seq2, m, s.stream));
            ctx->launches += 2;
        }
        SWB_CUDA(ctx, launch_for(prm, L, s.d_seq1, s.d_seq2, s.d_scores, m, s.stream, shared_target));
        ctx->launches += 1;
        SWB_CUDA(ctx, cudaMemcpyAsync(scores + c0, s.d_scores, m * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        SWB_CUDA(ctx, cudaEventRecord(s.done, s.stream));
        s.busy = true;
    }
    for (Slot& s : d->slots)
        if (s.busy) { s.busy = false; SWB_CUDA(ctx, cudaEventSynchronize(s.done)); }
    return SWB200_OK;
}

int run_range(swb200_ctx* ctx, Device* d, const uint8_t* seq1, const uint8_t* seq2, bool packed, int L,
              const SwParams& prm, int32_t* scores, uint64_t lo, uint64_t hi, bool shared_target = false)
{
    if (hi <= lo) return SWB200_OK;
    std::lock_guard<std::mutex> lock(d->mu);
    const int rc = run_range_locked(ctx, d, seq1, seq2, packed, L, prm, scores, lo, hi, shared_target);
    if (rc != SWB200_OK) drain_slots(d);      // nothing may still be writing the caller's `scores` once the error is returned
    return rc;
}

// The persistent-kernel batch packer (a textual part of this translation unit)
#include "feed.inc"

// The per-pair call's own path (a textual part of this translation unit)
#include "pairpath.inc"

// Live integer-pipe peak for the benchmark's roofline (a textual part of this translation unit)
#include "peakprobe.inc"

// Semi-global X-drop aligner: scratch, launches and the host pipeline (a textual part of this translation unit)
#include "sg_host.inc"

// One host thread per GPU of the context; body(k) returns that GPU's code.  A thread that cannot be created (or an
// allocation that fails on the way) becomes SWB200_ERR_NOMEM after the threads that did start have been joined:
// the C ABI never lets an exception out and never leaves a joinable thread behind.
template <class Body>
int run_per_device(swb200_ctx* ctx, size_t G, Body body)
{
    std::vector<int> rcs;
    std::vector<std::thread> pool;
    int spawn_rc = SWB200_OK;
    try {
        rcs.assign(G, SWB200_OK);
        pool.reserve(G);
        for (size_t k = 0; k < G; ++k) pool.emplace_back([&rcs, &body, k] { rcs[k] = body(k); });
    } catch (const std::exception& e) {
        spawn_rc = fail(ctx, SWB200_ERR_NOMEM, e.what());
    }
    for (auto& t : pool) t.join();
    if (spawn_rc != SWB200_OK) return spawn_rc;
    for (int r : rcs) if (r != SWB200_OK) return r;
    return SWB200_OK;
}

int check_args(swb200_ctx* ctx, const void* a, const void* b, const int8_t* sm, int gap, const void* out, uint64_t n)
{
    if (!ctx) return SWB200_ERR_ARG;
    if (!sm) return fail(ctx, SWB200_ERR_ARG, "score_matrix is NULL");
    if (n != 0 && (!a || !b || !out)) return fail(ctx, SWB200_ERR_ARG, "NULL array with n > 0");
    const int dom = sw_check_domain(sm, gap);
    if (dom == SW_DOMAIN_BAD_MATRIX) return fail(ctx, SWB200_ERR_DOMAIN, "score_matrix entry -128 is outside the reference's domain [-127,127]");
    if (dom == SW_DOMAIN_BAD_GAP) return fail(ctx, SWB200_ERR_DOMAIN, "gap_penalty outside [0,127]");
    return SWB200_OK;
}

int check_len(swb200_ctx* ctx, const int8_t* sm, int L)
{
    if (L != 128 && L != 256 && L != 512) return fail(ctx, SWB200_ERR_ARG, "seq_len must be 128, 256 or 512");
    if (!sw_len_supported(sm, L)) return fail(ctx, SWB200_ERR_DOMAIN, "seq_len * max(score_matrix) exceeds the packed int16 range");
    return SWB200_OK;
}

int score_host(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, bool packed,
               const int8_t* sm, int8_t gap, int32_t* scores, uint64_t n, int L = SWB200_SEQ_LEN, bool shared_target = false)
{
    int rc = check_args(ctx, seq1, seq2, sm, gap, scores, n);
    if (rc == SWB200_OK) rc = check_len(ctx, sm, L);
    if (rc != SWB200_OK || n == 0) return rc;
    // A handful of pairs: the latency kernel (one warp per pair, no staging) -- the per-pair call lives here.
    if (L == SWB200_SEQ_LEN && !packed && n <= kPairPathMax && ctx->latency_path && !ctx->force_general)
        return pair_run(ctx, ctx->devs[0], seq1, seq2, sm, gap, scores, n, shared_target);
    const SwParams prm = sw_make_params(sm, gap, ctx->force_general, L);
    const size_t G = ctx->devs.size();
    pair_dismiss(ctx->devs[0]);          // a resident per-pair server (GPU 0 only) makes room
    // Batches of the reference shape go through the persistent-kernel packer (feed.inc); byte-coded ones with the host
    // 2-bit packing lanes beside the raw copies.  Everything else (small batches, L = 256 / 512, one shared target)
    // takes the chunk pipeline.
    const bool feed = (L == SWB200_SEQ_LEN) && !shared_target && n / G >= kFeedMinPairs;
    const int n_pack = (feed && !packed) ? pack_threads_per_gpu(ctx) : 0;
    auto range = [=](Device* d, uint64_t lo, uint64_t hi) {
        return feed ? feed_run(ctx, d, seq1, seq2, packed, prm, scores, lo, hi, n_pack)
                    : run_range(ctx, d, seq1, seq2, packed, L, prm, scores, lo, hi, shared_target);
    };
    if (G == 1 || n < 2 * G) return range(ctx->devs[0], 0, n);
    // Contiguous index ranges [k*n/G, (k+1)*n/G), one host thread per GPU;
    // every GPU DMA-writes its own slice of `scores`: that is the whole gather.
    return run_per_device(ctx, G, [&](size_t k) { return range(ctx->devs[k], n * k / G, n * (k + 1) / G); });
}

} // namespace

extern "C" {

int swb200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail(nullptr, SWB200_ERR_NO_DEVICE, "cudaGetDeviceCount", e); return SWB200_ERR_NO_DEVICE; }
    int usable = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++usable;
    }
    return usable;
}

int swb200_init(swb200_ctx** out, const int* devices, int n_devices)
{
    if (!out || n_devices < 0) return fail(nullptr, SWB200_ERR_ARG, "swb200_init: bad arguments");
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
        return fail(nullptr, SWB200_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)", e);
    // n_devices == 0: every USABLE device (compute capability 10.x, what swb200_device_count() reports); a device list
    // makes no sense then.  Otherwise `devices` (or 0..n-1) is taken as given and a non-sm_100 device is an error.
    std::vector<int> ids;
    try {
        if (n_devices == 0) {
            if (devices) return fail(nullptr, SWB200_ERR_ARG, "swb200_init: a device list needs n_devices > 0");
            for (int i = 0; i < visible; ++i) {
                int major = 0;
                if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ids.push_back(i);
            }
            if (ids.empty()) return fail(nullptr, SWB200_ERR_NO_DEVICE, "no compute capability 10.x device visible (kernels are sm_100a only)");
        } else {
            if (n_devices > visible) return fail(nullptr, SWB200_ERR_NO_DEVICE, "more devices requested than visible");
            for (int k = 0; k < n_devices; ++k) ids.push_back(devices ? devices[k] : k);
        }
    } catch (const std::exception& e) {
        return fail(nullptr, SWB200_ERR_NOMEM, e.what());
    }
    n_devices = (int)ids.size();
    swb200_ctx* ctx = nullptr;
    try {
        ctx = new swb200_ctx;
        for (int k = 0; k < n_devices; ++k) {
            Device* d = new Device;
            d->id = ids[k];
            ctx->devs.push_back(d);
            int major = 0;
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d->id);
            int rc = (major == 10) ? setup_device(ctx, d)
                                   : fail(ctx, SWB200_ERR_NO_DEVICE, "device is not compute capability 10.x (kernels are sm_100a only)");
            if (rc != SWB200_OK) {
                g_init_error = ctx->err;
                swb200_shutdown(ctx);
                return rc;
            }
        }
    } catch (const std::exception& e) {          // allocation failure: a code, like every other error of this ABI
        if (ctx) swb200_shutdown(ctx);
        return fail(nullptr, SWB200_ERR_NOMEM, e.what());
    }
    *out = ctx;
    return SWB200_OK;
}

void swb200_shutdown(swb200_ctx* ctx)
{
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lock(ctx->tickets_mu);
        for (auto& kv : ctx->tickets) { if (kv.second.first.joinable()) kv.second.first.join(); delete kv.second.second; }
        ctx->tickets.clear();
    }
    for (Device* d : ctx->devs) {
        cudaSetDevice(d->id);
        feed_shutdown(d);
        pair_shutdown(d);
        for (Slot& s : d->slots) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            if (s.busy && s.done) cudaEventSynchronize(s.done);   // a packed-device call on the caller's stream may still read the staging
            cudaFree(s.d_seq1); cudaFree(s.d_seq2); cudaFree(s.d_pk1); cudaFree(s.d_pk2); cudaFree(s.d_scores);
            if (s.done) cudaEventDestroy(s.done);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        cudaFree(d->d_bad);
        sg_pipe_shutdown(d);
        cudaFree(d->sg_dev.traces);
        for (auto& g : d->sg_slots) {
            if (g.stream) { cudaStreamSynchronize(g.stream); cudaStreamDestroy(g.stream); }
            cudaFree(g.scratch.traces);
            cudaFree(g.d_seq1); cudaFree(g.d_seq2); cudaFree(g.d_ops); cudaFree(g.d_meta);
        }
        delete d;
    }
    delete ctx;
}

int swb200_n_devices(const swb200_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

const char* swb200_last_error(const swb200_ctx* ctx)
{
    if (!ctx) return g_init_error.c_str();
    // worker threads may be writing ctx->err: copy it under the lock into this thread's own buffer
    std::lock_guard<std::mutex> lk(const_cast<swb200_ctx*>(ctx)->err_mu);
    g_last_error_copy = ctx->err;
    return g_last_error_copy.c_str();
}

const char* swb200_strerror(int code)
{
    switch (code) {
    case SWB200_OK: return "ok";
    case SWB200_ERR_ARG: return "bad argument";
    case SWB200_ERR_DOMAIN: return "score matrix or gap penalty outside the reference's domain";
    case SWB200_ERR_NO_DEVICE: return "no usable sm_100 device";
    case SWB200_ERR_CUDA: return "CUDA call failed";
    case SWB200_ERR_NOMEM: return "out of memory";
    case SWB200_ERR_TICKET: return "unknown ticket";
    default: return "unknown error";
    }
}

int swb200_alloc_pinned(void** ptr, size_t bytes)
{
    if (!ptr) return SWB200_ERR_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, SWB200_ERR_NOMEM, "cudaHostAlloc", e);
    return SWB200_OK;
}

int swb200_free_pinned(void* ptr)
{
    if (!ptr) return SWB200_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? SWB200_OK : SWB200_ERR_CUDA;
}

int swb200_score_batch(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int8_t gap, int32_t* scores, uint64_t n)
{
    return score_host(ctx, seq1, seq2, false, sm, gap, scores, n);
}

int swb200_score_batch_packed(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int8_t gap, int32_t* scores, uint64_t n)
{
    return score_host(ctx, seq1, seq2, true, sm, gap, scores, n);
}

int swb200_score_one_vs_many(swb200_ctx* ctx, const uint8_t* seq1s, const uint8_t* seq2, const int8_t* sm, int8_t gap,
                             int32_t* scores, uint64_t n)
{
    return score_host(ctx, seq1s, seq2, false, sm, gap, scores, n, SWB200_SEQ_LEN, true);
}

int swb200_score_batch_111(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, int32_t* scores, uint64_t n)
{
    static const int8_t m111[16] = {1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1};   // source.cpp:1079
    return score_host(ctx, seq1, seq2, false, m111, 1, scores, n);
}

int swb200_score_pair(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int8_t gap, int32_t* score)
{
    return score_host(ctx, seq1, seq2, false, sm, gap, score, 1);
}

static int submit_impl(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, bool packed, const int8_t* sm, int8_t gap,
                       int32_t* scores, uint64_t n, swb200_ticket* ticket)
{
    if (!ticket) return ctx ? fail(ctx, SWB200_ERR_ARG, "ticket is NULL") : SWB200_ERR_ARG;
    int rc = check_args(ctx, seq1, seq2, sm, gap, scores, n);
    if (rc != SWB200_OK) return rc;
    std::array<int8_t, 16> m;
    memcpy(m.data(), sm, 16);
    std::lock_guard<std::mutex> lock(ctx->tickets_mu);
    int* result = nullptr;
    std::thread th;
    try {
        result = new int(SWB200_OK);
        auto slot = ctx->tickets.emplace(ctx->next_ticket, std::make_pair(std::thread(), result)).first;   // the map node first: it cannot fail later
        th = std::thread([=] { *result = score_host(ctx, seq1, seq2, packed, m.data(), gap, scores, n); });
        slot->second.first = std::move(th);
    } catch (const std::exception& e) {          // no thread, no memory: an error code, and no half-made ticket
        ctx->tickets.erase(ctx->next_ticket);
        delete result;
        return fail(ctx, SWB200_ERR_NOMEM, e.what());
    }
    *ticket = ctx->next_ticket++;
    return SWB200_OK;
}

int swb200_submit(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int8_t gap,
                  int32_t* scores, uint64_t n, swb200_ticket* ticket)
{
    return submit_impl(ctx, seq1, seq2, false, sm, gap, scores, n, ticket);
}

int swb200_submit_packed(swb200_ctx* ctx, const uint8_t* seq1_packed, const uint8_t* seq2_packed, const int8_t* sm, int8_t gap,
                         int32_t* scores, uint64_t n, swb200_ticket* ticket)
{
    return submit_impl(ctx, seq1_packed, seq2_packed, true, sm, gap, scores, n, ticket);
}

int swb200_wait(swb200_ctx* ctx, swb200_ticket ticket)
{
    if (!ctx) return SWB200_ERR_ARG;
    std::thread th;
    int* result = nullptr;
    {
        std::lock_guard<std::mutex> lock(ctx->tickets_mu);
        auto it = ctx->tickets.find(ticket);
        if (it == ctx->tickets.end()) return fail(ctx, SWB200_ERR_TICKET, "unknown or already-waited ticket");
        th = std::move(it->second.first);
        result = it->second.second;
        ctx->tickets.erase(it);
    }
    th.join();
    const int rc = *result;
    delete result;
    return rc;
}

static int device_args(swb200_ctx* ctx, int device_index, const void* a, const void* b, const void* out, uint64_t n)
{
    if (!ctx) return SWB200_ERR_ARG;
    if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SWB200_ERR_ARG, "device_index out of range");
    if (n && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15u)) return fail(ctx, SWB200_ERR_ARG, "device arrays must be 16-byte aligned");
    pair_dismiss(ctx->devs[device_index]);       // a resident per-pair server makes room for the caller's launch
    return SWB200_OK;
}

int swb200_score_batch_len_device(swb200_ctx* ctx, int device_index, int seq_len, const uint8_t* d_seq1, const uint8_t* d_seq2,
                                  const int8_t* sm, int8_t gap, int32_t* d_scores, uint64_t n, void* cuda_stream)
{
    int rc = check_args(ctx, d_seq1, d_seq2, sm, gap, d_scores, n);
    if (rc == SWB200_OK) rc = check_len(ctx, sm, seq_len);
    if (rc == SWB200_OK) rc = device_args(ctx, device_index, d_seq1, d_seq2, d_scores, n);
    if (rc != SWB200_OK || n == 0) return rc;
    SWB_CUDA(ctx, cudaSetDevice(ctx->devs[device_index]->id));
    const SwParams prm = sw_make_params(sm, gap, ctx->force_general, seq_len);
    SWB_CUDA(ctx, launch_for(prm, seq_len, d_seq1, d_seq2, d_scores, n, (cudaStream_t)cuda_stream));
    ctx->launches += 1;
    return SWB200_OK;
}

int swb200_score_batch_device(swb200_ctx* ctx, int device_index, const uint8_t* d_seq1, const uint8_t* d_seq2,
                              const int8_t* sm, int8_t gap, int32_t* d_scores, uint64_t n, void* cuda_stream)
{
    return swb200_score_batch_len_device(ctx, device_index, SWB200_SEQ_LEN, d_seq1, d_seq2, sm, gap, d_scores, n, cuda_stream);
}

int swb200_score_batch_len(swb200_ctx* ctx, int seq_len, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int8_t gap,
                           int32_t* scores, uint64_t n)
{
    return score_host(ctx, seq1, seq2, false, sm, gap, scores, n, seq_len);
}

int swb200_score_batch_packed_device(swb200_ctx* ctx, int device_index, const uint8_t* d_pk1, const uint8_t* d_pk2,
                                     const int8_t* sm, int8_t gap, int32_t* d_scores, uint64_t n, void* cuda_stream)
{
    int rc = check_args(ctx, d_pk1, d_pk2, sm, gap, d_scores, n);
    if (rc == SWB200_OK) rc = device_args(ctx, device_index, d_pk1, d_pk2, d_scores, n);
    if (rc != SWB200_OK || n == 0) return rc;
    Device* d = ctx->devs[device_index];
    std::lock_guard<std::mutex> lock(d->mu);
    SWB_CUDA(ctx, cudaSetDevice(d->id));
    const SwParams prm = sw_make_params(sm, gap, ctx->force_general);
    // The persistent consumer kernel with every tile flag already set: a block expands its own 128 pairs into the
    // context's byte staging (L2-resident) and scores them -- one launch per 2 M pairs, stream-ordered on the caller's stream.
    return feed_packed_resident(ctx, d, d_pk1, d_pk2, prm, d_scores, n, (cudaStream_t)cuda_stream);
}

int swb200_validate_codes_device(swb200_ctx* ctx, int device_index, const uint8_t* d_codes, uint64_t n_bytes,
                                 uint64_t* n_bad, void* cuda_stream)
{
    if (!ctx || !n_bad) return SWB200_ERR_ARG;
    if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SWB200_ERR_ARG, "device_index out of range");
    Device* d = ctx->devs[device_index];
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::lock_guard<std::mutex> lock(d->mu);     // the counter d_bad is one per device: calls on other streams / threads wait
    SWB_CUDA(ctx, cudaSetDevice(d->id));
    SWB_CUDA(ctx, cudaMemsetAsync(d->d_bad, 0, sizeof(unsigned long long), st));
    if (n_bytes) {
        count_bad_codes_kernel<<<d->prop.multiProcessorCount * 8, 256, 0, st>>>(d_codes, n_bytes, d->d_bad);
        SWB_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    unsigned long long h = 0;
    SWB_CUDA(ctx, cudaMemcpyAsync(&h, d->d_bad, sizeof h, cudaMemcpyDeviceToHost, st));
    SWB_CUDA(ctx, cudaStreamSynchronize(st));
    *n_bad = h;
    return SWB200_OK;
}

int swb200_kernel_info_len(swb200_ctx* ctx, int device_index, int seq_len, const int8_t* sm, int8_t gap, swb200_kernel_info* info)
{
    if (!ctx || !info || !sm) return SWB200_ERR_ARG;
    if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SWB200_ERR_ARG, "device_index out of range");
    if (sw_check_domain(sm, gap) != SW_DOMAIN_OK) return fail(ctx, SWB200_ERR_DOMAIN, "matrix/gap outside the domain");
    int rc = check_len(ctx, sm, seq_len);
    if (rc != SWB200_OK) return rc;
    Device* d = ctx->devs[device_index];
    SWB_CUDA(ctx, cudaSetDevice(d->id));
    const SwParams prm = sw_make_params(sm, gap, ctx->force_general, seq_len);
    cudaFuncAttributes fa{};
    int blocks = 0, nt = 0, smem = 0;
    cudaError_t e = cudaErrorInvalidValue;
    switch (seq_len) {
    case 128: e = prm.fast ? kernel_resources<true, 128>(&fa, &blocks, &nt, &smem) : kernel_resources<false, 128>(&fa, &blocks, &nt, &smem); break;
    case 256: e = prm.fast ? kernel_resources<true, 256>(&fa, &blocks, &nt, &smem) : kernel_resources<false, 256>(&fa, &blocks, &nt, &smem); break;
    case 512: e = prm.fast ? kernel_resources<true, 512>(&fa, &blocks, &nt, &smem) : kernel_resources<false, 512>(&fa, &blocks, &nt, &smem); break;
    }
    SWB_CUDA(ctx, e);
    info->fast_path = prm.fast;
    info->regs_per_thread = fa.numRegs;
    info->threads_per_block = nt;
    info->blocks_per_sm = blocks;
    info->smem_bytes_per_block = smem;
    info->sm_count = d->prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d->id);
    info->sm_clock_khz = khz;
    return SWB200_OK;
}

int swb200_kernel_info_for(swb200_ctx* ctx, int device_index, const int8_t* sm, int8_t gap, swb200_kernel_info* info)
{
    return swb200_kernel_info_len(ctx, device_index, SWB200_SEQ_LEN, sm, gap, info);
}

// Semi-global X-drop aligner: the C-ABI entry points (a textual part of this translation unit)
#include "sg_abi.inc"

int swb200_measure_alu_peak(swb200_ctx* ctx, int device_index, double target_ms, double* tinstr_per_s, double* elapsed_ms)
{
    if (!ctx || !tinstr_per_s) return SWB200_ERR_ARG;
    if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SWB200_ERR_ARG, "device_index out of range");
    if (!(target_ms > 0.0) || target_ms > 2000.0) return fail(ctx, SWB200_ERR_ARG, "target_ms must be in (0, 2000]");
    return measure_alu_peak(ctx, ctx->devs[device_index], target_ms, tinstr_per_s, elapsed_ms);
}

uint64_t swb200_launch_count(const swb200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int swb200_set_host_pack_threads(swb200_ctx* ctx, int threads_per_gpu)
{
    if (!ctx) return SWB200_ERR_ARG;
    ctx->pack_threads = threads_per_gpu < 0 ? -1 : threads_per_gpu;    // the lane pool is rebuilt with the new size by the next batch
    return SWB200_OK;
}

int swb200_host_pack_stats(const swb200_ctx* ctx, uint64_t* packed_pairs, uint64_t* raw_pairs, int* threads_per_gpu)
{
    if (!ctx) return SWB200_ERR_ARG;
    if (packed_pairs) *packed_pairs = ctx->packed_pairs.load();
    if (raw_pairs) *raw_pairs = ctx->raw_pairs.load();
    if (threads_per_gpu) *threads_per_gpu = pack_threads_per_gpu(ctx);
    return SWB200_OK;
}

int swb200_host_pack_tuning(const swb200_ctx* ctx, int device_index, int* lanes_in_use, int* raw_lane_in_use, double* pairs_per_s)
{
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) return SWB200_ERR_ARG;
    Device* d = ctx->devs[device_index];
    std::lock_guard<std::mutex> lock(d->mu);
    const FeedState& fs = *d->feed;
    const int n_pack = pack_threads_per_gpu(ctx);
    const bool tuned = ctx->pack_threads < 0 && fs.tune_calls > 0;
    if (lanes_in_use) *lanes_in_use = tuned ? feed_tune_lanes(fs.tune_best, n_pack) : n_pack;
    if (raw_lane_in_use) *raw_lane_in_use = tuned ? (feed_tune_raw(fs.tune_best) ? 1 : 0) : 1;
    if (pairs_per_s) for (int k = 0; k < FeedState::kTuneCandidates; ++k) pairs_per_s[k] = fs.tune_rate[k];
    return SWB200_OK;
}

int swb200_pair_path_stats(const swb200_ctx* ctx, uint64_t* server_launches, uint64_t* doorbell_calls, uint32_t* last_sweep_ns)
{
    if (!ctx || ctx->devs.empty()) return SWB200_ERR_ARG;
    Device* d = ctx->devs[0];
    PairSlot* ps = __atomic_load_n(&d->pair, __ATOMIC_ACQUIRE);
    uint64_t launches = 0, calls = 0;
    uint32_t ns = 0;
    if (ps) {
        std::lock_guard<std::mutex> lock(ps->mu);
        launches = ps->server_launches;
        calls = ps->doorbell_calls;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d->id);
        const uint32_t cycles = const_cast<const volatile PairMailbox*>(ps->h_mb)->sweep_ns;
        ns = khz > 0 ? (uint32_t)((uint64_t)cycles * 1000000ull / (uint64_t)khz) : 0;
    }
    if (server_launches) *server_launches = launches;
    if (doorbell_calls) *doorbell_calls = calls;
    if (last_sweep_ns) *last_sweep_ns = ns;
    return SWB200_OK;
}

int swb200_unpack2bit_host(const uint8_t* packed, uint8_t* codes, uint64_t n_codes)
{
    if ((!codes || !packed) && n_codes) return SWB200_ERR_ARG;
    unpack2bit_host(packed, codes, (size_t)n_codes);
    return SWB200_OK;
}

int swb200_pack2bit_host(const uint8_t* codes, uint8_t* packed, uint64_t n_codes)
{
    if ((!codes || !packed) && n_codes) return SWB200_ERR_ARG;
    if (n_codes % 8) return SWB200_ERR_ARG;
    pack2bit_host(codes, packed, (size_t)n_codes);
    return SWB200_OK;
}

int swb200_set_latency_path(swb200_ctx* ctx, int on)
{
    if (!ctx) return SWB200_ERR_ARG;
    ctx->latency_path = on ? 1 : 0;
    ctx->pair_doorbell = (on != 2) ? 1 : 0;      // 2: the latency kernel with one launch per call, also for a single pair
    return SWB200_OK;
}

int swb200_set_force_general(swb200_ctx* ctx, int on)
{
    if (!ctx) return SWB200_ERR_ARG;
    ctx->force_general = on ? 1 : 0;
    return SWB200_OK;
}

} // extern "C"

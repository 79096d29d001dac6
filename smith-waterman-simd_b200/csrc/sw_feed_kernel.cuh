// sw_feed_kernel.cuh -- the persistent CONSUMER kernel of the host batch packer.
//
// One launch scores one "epoch" of a host batch (up to the staging capacity of a region,
// feed.inc) while its input is still crossing PCIe.  The grid is the resident block count of
// the scoring kernel (SMs x 6 blocks of 64 threads); a block repeatedly
//   1. claims the next work item (128 consecutive pairs) from a counter,
//   2. waits until the 4096-pair tile holding the item has LANDED: the host enqueues, behind
//      every input copy and on the same stream, a 4-byte copy that writes the tile's flag, so
//      the copy engine itself publishes "tile t is in HBM, in format f" in stream order;
//   3. if the tile arrived in the reference's 2-bit packing (source.cpp:1580-1583) expands
//      its own 128 pairs into the byte staging (8 KiB in, 32 KiB out, L2-resident);
//   4. scores them with the same per-thread code as sw_kernel (sw_core.cuh), and
//   5. writes the scores straight to where the caller wants them -- mapped pinned host memory
//      (zero-copy stores, 256 B per warp) or a device array.
// So a host batch is ONE kernel launch however many copies feed it, nothing waits for a whole
// chunk, and the only exposed PCIe time is the first tile's.
//
// THE RULE that keeps this safe: a launch only ever waits for copies that were ENQUEUED BEFORE IT.
// Streams are multiplexed onto a few in-order hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by
// default); a copy enqueued behind a spinning kernel on a queue it happens to share would never
// start, and the kernel would wait for it for ever.  Work enqueued earlier is ahead of the kernel
// on every queue.  Input that the host produces while the GPU is already scoring (byte-coded
// batches compressed by the PACK lanes) is therefore cut into WINDOWS, each with its own launch,
// issued when the window's last piece has been enqueued (feed.inc).
//
// The wait is bounded: a block that sees no flag for `timeout_ns` (or an abort word set by the
// host after a failed enqueue) reports through `status` and every block leaves.
#pragma once
#include <cuda_runtime.h>
#include "sw_kernel.cuh"

namespace swb {

constexpr int FEED_NT = 64;                    // threads per block (= sw_kernel's shape at L = 128)
constexpr int FEED_MINB = 6;                   // resident blocks per SM
constexpr int FEED_ITEM_PAIRS = 2 * FEED_NT;   // pairs per work item
constexpr int FEED_TILE_PAIRS = 4096;          // flag granularity: one `ready` word per 4096 pairs
constexpr int FEED_ITEMS_PER_TILE = FEED_TILE_PAIRS / FEED_ITEM_PAIRS;

// `ready[t]` holds 2*epoch + format once tile t of launch `epoch` is in HBM.
constexpr uint32_t FEED_FMT_RAW = 0;           // byte codes in raw1/raw2
constexpr uint32_t FEED_FMT_PACKED = 1;        // 2-bit codes in pk1/pk2; the consuming block expands them into raw1/raw2

enum : uint32_t { FEED_STATUS_OK = 0, FEED_STATUS_TIMEOUT = 1, FEED_STATUS_ABORTED = 2 };

struct FeedArgs {
    uint8_t* raw1;                  // [cap][128] byte-coded staging (device)
    uint8_t* raw2;
    const uint8_t* pk1;             // [cap][32] 2-bit staging (device)
    const uint8_t* pk2;
    int32_t* scores;                // [n] device array, or mapped pinned host memory
    const uint32_t* ready;          // [ceil(n / 4096)] tile flags, written by the copy engine (and mirrored there by the relay)
    const uint32_t* host_flags;     // LANES epochs: the PACK lanes' own flag words in mapped pinned memory (else NULL), see below
    uint32_t* next_item;            // this launch's work counter: never reset, the host knows where it stands ...
    uint32_t item_base;             // ... before this launch (a launch of n items on g blocks adds exactly n + g)
    uint32_t first_item;            // the launch scores items [first_item, first_item + n_items) of the epoch
    uint32_t n_items;
    volatile uint32_t* status;      // mapped pinned host word: FEED_STATUS_*
    const volatile uint32_t* abort; // device word the host sets after a failed enqueue: non-zero = give up
    unsigned long long timeout_ns;
    uint32_t n;                     // pairs in this epoch
    uint32_t epoch;                 // flags of this launch are >= 2*epoch
    uint32_t scores_vec2;           // 1: scores is 8-byte aligned (int2 stores)
    uint32_t all_there;             // 1: every tile was complete before the launch (device-resident input, or 2-bit input the
                                    //    blocks pull straight from the caller's pinned arrays): no flags, no fence; format = fmt_all
    uint32_t fmt_all;
};

// The flag is polled with a RELAXED system-scope load (one LDG.E.STRONG.SYS, served by L2 where the copy engine's
// write lands).  An acquire load would be followed by CCTL.IVALL -- the whole SM's L1 invalidated on every poll, under
// the feet of the five other blocks that are scoring -- so the acquire is a single fence once the flag has been seen.
__device__ __forceinline__ uint32_t feed_ld_flag(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void feed_acquire_fence()
{
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

__device__ __forceinline__ unsigned long long feed_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// 4 packed bytes (16 bases) -> 16 byte codes: code(4i + j) = (byte[i] >> 2j) & 3   (source.cpp:1580-1583)
__device__ __forceinline__ uint4 feed_expand_word(uint32_t w)
{
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const uint32_t x = (w >> (8 * b)) & 0xffu;
        o[b] = (x & 3u) | (((x >> 2) & 3u) << 8) | (((x >> 4) & 3u) << 16) | (((x >> 6) & 3u) << 24);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// The block's own m pairs of one sequence array: m*32 packed bytes -> m*128 byte codes.  One 32-bit word (16 bases) per
// thread and step, so a warp reads 128 contiguous bytes and writes 512; all 16 loads of a thread are issued before the
// first store (the expansion is a latency chain otherwise: 9 us per item measured, 4.7 % of the kernel).
__device__ __forceinline__ void feed_expand_item(const uint8_t* pk, uint8_t* raw, uint32_t m)
{
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pk);
    uint4* dst = reinterpret_cast<uint4*>(raw);
    const uint32_t n_words = m * 8u;                                   // 8 words per sequence
    constexpr int PER = FEED_ITEM_PAIRS * 8 / FEED_NT;                 // 16 words per thread for a full item
    uint32_t w[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t i = threadIdx.x + (uint32_t)j * FEED_NT;
        w[j] = (i < n_words) ? __ldcg(src + i) : 0u;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t i = threadIdx.x + (uint32_t)j * FEED_NT;
        if (i < n_words) dst[i] = feed_expand_word(w[j]);
    }
}

template <bool FAST>
__global__ void __launch_bounds__(FEED_NT, FEED_MINB)
sw_feed_kernel(const FeedArgs fa, const SwParams prm)
{
    constexpr int L = SW_L;
    constexpr int V = SW_DEFAULT_VARIANT | SW_V_COHERENT_LD;
    extern __shared__ uint32_t smem[];
    uint32_t* t4s = smem + L * FEED_NT;
    __shared__ uint32_t s_item, s_fmt;
    if (threadIdx.x < 4) t4s[threadIdx.x] = prm.t4[threadIdx.x];

    const uint32_t want = 2u * fa.epoch;
    SmemFifo<FEED_NT> fifo{smem + threadIdx.x};
    SmemTable t4{t4s};

    for (;;) {
        __syncthreads();                                   // s_item / s_fmt of the previous round have been read
        if (threadIdx.x == 0) {
            uint32_t item = atomicAdd(fa.next_item, 1u) - fa.item_base;
            uint32_t fmt = 0xffffffffu;                    // "leave"
            if (item < fa.n_items && fa.all_there) {
                item += fa.first_item;
                fmt = fa.fmt_all;
            } else if (item < fa.n_items) {
                item += fa.first_item;
                const uint32_t* flag = fa.ready + item / FEED_ITEMS_PER_TILE;
                uint32_t v = feed_ld_flag(flag);
                if (v < want || v > want + 1u) {           // not landed yet (a flag of an older epoch is smaller)
                    const unsigned long long t0 = feed_now_ns();
                    unsigned spins = 0, nap = 100;
                    for (;;) {
                        __nanosleep(nap);
                        if (nap < 1600) nap *= 2;          // back off: hundreds of blocks may be waiting when the link is the bottleneck
                        v = feed_ld_flag(flag);
                        if (v >= want && v <= want + 1u) break;
                        // Progress must not depend on the relay warp finding a place on an SM (the consumer blocks leave it
                        // barely a warp's worth of registers): now and then a waiting block looks at its tile's flag in host
                        // memory itself -- a PCIe read, so rarely.
                        if (fa.host_flags && (spins & 31u) == 31u) {
                            const uint32_t h = feed_ld_flag(fa.host_flags + (flag - fa.ready));
                            if (h == want + FEED_FMT_PACKED) { v = h; break; }
                        }
                        if ((++spins & 63u) == 0u) {
                            if (*fa.abort) { *fa.status = FEED_STATUS_ABORTED; v = 0xffffffffu; break; }
                            if (feed_now_ns() - t0 > fa.timeout_ns) { *fa.status = FEED_STATUS_TIMEOUT; v = 0xffffffffu; break; }
                        }
                    }
                }
                feed_acquire_fence();                      // the tile's bytes were written before its flag
                fmt = (v == 0xffffffffu) ? v : (v - want);
                if (fmt == 0xffffffffu) atomicAdd(fa.next_item, 0x40000000u);   // every later claim is past the end: all blocks leave
            }
            s_item = item;
            s_fmt = fmt;
        }
        __syncthreads();
        const uint32_t item = s_item, fmt = s_fmt;
        if (fmt == 0xffffffffu) return;

        const uint32_t p0 = item * FEED_ITEM_PAIRS;
        const uint32_t m = (fa.n - p0 < (uint32_t)FEED_ITEM_PAIRS) ? fa.n - p0 : (uint32_t)FEED_ITEM_PAIRS;
        if (fmt == FEED_FMT_PACKED) {
            feed_expand_item(fa.pk1 + (size_t)p0 * 32, fa.raw1 + (size_t)p0 * L, m);
            feed_expand_item(fa.pk2 + (size_t)p0 * 32, fa.raw2 + (size_t)p0 * L, m);
            __syncthreads();                               // the block's stores are visible to the block's (plain) loads
        }
        const uint32_t p = p0 + 2u * threadIdx.x;
        if (p < fa.n) {
            const bool two = p + 1u < fa.n;                // odd tail: the high half repeats the low pair
            int32_t lo, hi;
            sw_two_pairs<FAST, L, V>(fa.raw1 + (size_t)p * L, fa.raw2 + (size_t)p * L, two ? (uint32_t)L : 0u, two ? (uint32_t)L : 0u,
                                     fifo, t4, prm, lo, hi);
            if (two && fa.scores_vec2) *reinterpret_cast<int2*>(fa.scores + p) = make_int2(lo, hi);
            else { fa.scores[p] = lo; if (two) fa.scores[p + 1] = hi; }
        }
    }
}

// The RELAY of a laned epoch: one warp that mirrors tile flags the host's PACK lanes set with plain CPU stores (in mapped
// pinned memory) into the device flag array the consumer blocks poll -- so that hundreds of waiting blocks poll L2, and
// only this warp polls across PCIe.  A host flag holds 2*epoch + format once the tile's 2-bit bytes are in the pinned
// staging (the consumer pulls them over PCIe), or 2*epoch + FEED_FMT_BY_DMA when the RAW lane has enqueued the tile's
// copies (the copy engine then writes the device flag itself).  The warp leaves when every tile has been accounted for.
constexpr uint32_t FEED_FMT_BY_DMA = 2;

__global__ void __launch_bounds__(32) feed_relay_kernel(const uint32_t* h_flags, uint32_t* d_flags, uint32_t n_tiles, uint32_t epoch,
                                                         const volatile uint32_t* abort, unsigned long long timeout_ns)
{
    const uint32_t want = 2u * epoch;
    const unsigned long long t0 = feed_now_ns();
    uint32_t done_mask_lo = 0;                       // bit j: tile lane + 32*j accounted for (n_tiles <= 32 * 32 * ... see loop)
    unsigned spins = 0;
    for (;;) {
        bool all = true;
        for (uint32_t j = 0, t = threadIdx.x; t < n_tiles; ++j, t += 32) {
            if (j < 32 && (done_mask_lo >> j) & 1u) continue;
            const uint32_t v = feed_ld_flag(h_flags + t);                  // a PCIe read of host memory
            if (v >= want && v <= want + FEED_FMT_BY_DMA) {
                if (v != want + FEED_FMT_BY_DMA) {
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(d_flags + t), "r"(v) : "memory");
                }
                if (j < 32) done_mask_lo |= 1u << j;
            } else {
                all = false;
            }
        }
        if (__all_sync(0xffffffffu, all)) return;
        if ((++spins & 127u) == 0u) {
            if (*abort) return;
            if (feed_now_ns() - t0 > 4ull * timeout_ns) return;            // the consumers report the timeout
        }
        __nanosleep(500);
    }
}

constexpr size_t feed_smem_bytes() { return sw_smem_bytes<SW_L, FEED_NT>(); }

} // namespace swb

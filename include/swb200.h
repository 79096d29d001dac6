/*
 * swb200.h -- C ABI of the B200-native replacement for the one hot path of
 * eukaryo/smith-waterman-simd: score-only Smith-Waterman (linear gap, 4x4 int8
 * substitution matrix) of pairs of 128-base DNA sequences.
 *
 * The reference has no library boundary at all -- its kernels are free functions in one
 * translation unit (/root/reference/source.cpp) -- so each entry point below cites the
 * reference interface it stands in for.  INTEGRATION.md shows the binding a maintainer of
 * the reference would add.  Plain C: pointers and sizes only, no C++ or torch types.
 *
 * Semantics shared by all scoring entry points (the reference's contract):
 *   - sequences are bytes holding codes 0..3, 128 per sequence, sequences laid end to end
 *     (what `const std::array<uint8_t,128>&` is in memory, source.cpp:36-37);
 *   - score_matrix is 16 int8, indexed seq1_code*4 + seq2_code (source.cpp:38,50);
 *   - gap_penalty is the linear gap cost, subtracted (source.cpp:39,51-52);
 *   - the score is max(0, best local alignment score) as an int (source.cpp:41,53,59).
 * Parameter domain: matrix entries in [-127,127], gap in [0,127] -- the domain on which
 * the reference's AVX2 kernels agree with its scalar kernel.  Outside it
 * the call returns SWB200_ERR_DOMAIN instead of a silently different number.
 * Codes above 3 are the caller's error, as in the reference (which indexes unchecked,
 * source.cpp:50); the library never reads out of bounds for them but the score is
 * unspecified.  swb200_validate_codes() checks a buffer on the device.
 *
 * There is NO CPU fallback: every scoring call runs the sm_100a kernel or fails.
 */
#ifndef SWB200_H
#define SWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB200_SEQ_LEN 128   /* std::array<uint8_t,128>, source.cpp:36-37 */

enum {
    SWB200_OK = 0,
    SWB200_ERR_ARG = -1,        /* null pointer, bad handle, misaligned device pointer */
    SWB200_ERR_DOMAIN = -2,     /* matrix entry == -128 or gap outside [0,127] */
    SWB200_ERR_NO_DEVICE = -3,  /* no CUDA device / not an sm_100 device */
    SWB200_ERR_CUDA = -4,       /* a CUDA call failed; swb200_last_error() has the text */
    SWB200_ERR_NOMEM = -5,
    SWB200_ERR_TICKET = -6      /* unknown or already-waited ticket */
};

typedef struct swb200_ctx swb200_ctx;

/* ---- lifetime --------------------------------------------------------------------- */

/* Number of usable (compute capability 10.x) CUDA devices, or a negative error. */
int swb200_device_count(void);

/* Creates a context on `n_devices` GPUs.  devices == NULL means 0..n_devices-1;
 * n_devices == 0 means all visible.  The context owns streams, device staging buffers
 * and one worker per GPU; host arrays always stay the caller's.
 * (No reference counterpart: the reference keeps no state, source.cpp:35-60.) */
int swb200_init(swb200_ctx** ctx, const int* devices, int n_devices);
void swb200_shutdown(swb200_ctx* ctx);
int swb200_n_devices(const swb200_ctx* ctx);

/* Text of the last failure on this context (or of the last failed swb200_init when ctx == NULL). */
const char* swb200_last_error(const swb200_ctx* ctx);
const char* swb200_strerror(int code);

/* ---- the batch packer's host side ------------------------------------------------- */

/* Page-locked host memory: copies from/to it overlap with kernels.  Pageable host
 * arrays are accepted everywhere, they are just slower to move. */
int swb200_alloc_pinned(void** ptr, size_t bytes);
int swb200_free_pinned(void* ptr);

/* ---- scoring ---------------------------------------------------------------------- */

/* Replaces ONE call of
 *   int SmithWaterman_simdN(const std::array<uint8_t,128>& seq1, const std::array<uint8_t,128>& seq2,
 *                           const std::array<int8_t,16>& score_matrix, const int8_t gap_penalty)
 * (source.cpp:462-466; identical signatures at 35-39, 62, 210, 341, 573, 666, 758, 852, 953).
 * The score is written to *score.  The call rings a DOORBELL: it writes the pair, the matrix and the
 * gap into 320 bytes of mapped pinned memory that a resident one-warp server kernel polls, and spins on
 * the mapped word the server stores the tagged score into -- no launch, no copy, no stream
 * synchronisation in a run of calls (the shape SpeedTest times, source.cpp:3036-3054).  The server
 * leaves by itself when no call has come for SWB200_PAIR_LINGER_US (environment, default 200; 0 = no
 * resident server: one launch per call with the pair in the launch parameters); the next call launches
 * a new one.  Calls from several threads are serialised per context. */
int swb200_score_pair(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2,
                      const int8_t* score_matrix, int8_t gap_penalty, int32_t* score);

/* Replaces the loop `for (pair p) scores[p] = SmithWaterman_simdN(a[p], b[p], sm, gap)`
 * that the reference's harnesses run (source.cpp:2947-2970, 3049-3051, 3209-3235).
 * The batched-call precedent in the reference is SmithWaterman_8b111x32mark1
 * (source.cpp:1227-1234: 32 row-major 128-mers in, results into a `dest` array).
 * HOST arrays: seq1[n][128], seq2[n][128], scores[n].  Pairs are split in contiguous
 * index ranges over the context's GPUs.  Per GPU one persistent kernel consumes the range
 * while its input is still crossing PCIe (tile flags written by the copy engine) and stores
 * the scores straight into `scores` when that is pinned memory (swb200_alloc_pinned);
 * up to 2048 pairs take a one-warp-per-pair latency kernel instead (as swb200_score_pair). */
int swb200_score_batch(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2,
                       const int8_t* score_matrix, int8_t gap_penalty,
                       int32_t* scores, uint64_t n);

/* Many queries against ONE target: replaces
 *   int SmithWaterman_8b111x32markN(const std::array<uint8_t,128*32>& seq1, const std::array<uint8_t,128>& seq2,
 *                                   std::array<int,32>& dest)           (source.cpp:1227-1230, 1299, 1383)
 * for any n (not only 32) and any matrix/gap of the domain (the reference fixes 1/-1/1 there):
 * seq1s[n][128] row-major, seq2[128], scores[n] = score(seq1s[p], seq2). */
int swb200_score_one_vs_many(swb200_ctx* ctx, const uint8_t* seq1s, const uint8_t* seq2,
                             const int8_t* score_matrix, int8_t gap_penalty,
                             int32_t* scores, uint64_t n);

/* Fixed match/mismatch/gap = 1/1/1 scoring, replacing
 *   int SmithWaterman_111(const std::array<uint8_t,128>& seq1, const std::array<uint8_t,128>& seq2)      (source.cpp:1073-1103)
 *   int SmithWaterman_8bit111simd(same arguments)                                                        (source.cpp:1105-1225)
 * for n pairs.  The reference needs a separate 8-bit kernel for this scoring because AVX2 doubles
 * its lane count at 8 bits; sm_100a has no native u8x4 min/max, and the packed int16x2 kernel
 * already spends no instruction on the generality (the score comes from one PRMT either way), so
 * this entry runs the same kernel with the matrix fixed at +1/-1 and gap 1. */
int swb200_score_batch_111(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, int32_t* scores, uint64_t n);

/* Batch scoring with inputs as the reference's 2-bit packing: 32 bytes per sequence,
 * code(i*4+j) = (byte[i] >> 2j) & 3   (source.cpp:1580-1583, `unpack`).
 * seq1_packed[n][32], seq2_packed[n][32].  Unpacking happens on the device. */
int swb200_score_batch_packed(swb200_ctx* ctx, const uint8_t* seq1_packed, const uint8_t* seq2_packed,
                              const int8_t* score_matrix, int8_t gap_penalty,
                              int32_t* scores, uint64_t n);

/* Asynchronous form of swb200_score_batch for streaming callers: returns at once with a
 * ticket; the arrays must stay valid and untouched until swb200_wait(ticket) returns. */
typedef uint64_t swb200_ticket;
int swb200_submit(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2,
                  const int8_t* score_matrix, int8_t gap_penalty,
                  int32_t* scores, uint64_t n, swb200_ticket* ticket);
int swb200_submit_packed(swb200_ctx* ctx, const uint8_t* seq1_packed, const uint8_t* seq2_packed,
                         const int8_t* score_matrix, int8_t gap_penalty,
                         int32_t* scores, uint64_t n, swb200_ticket* ticket);
int swb200_wait(swb200_ctx* ctx, swb200_ticket ticket);

/* DEVICE arrays on GPU `device_index` of the context (16-byte aligned), enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = the legacy default stream); returns without
 * synchronising.  This is the kernel launch alone. */
int swb200_score_batch_device(swb200_ctx* ctx, int device_index,
                              const uint8_t* d_seq1, const uint8_t* d_seq2,
                              const int8_t* score_matrix, int8_t gap_penalty,
                              int32_t* d_scores, uint64_t n, void* cuda_stream);
/* The packed form is ONE launch per 2 M pairs of the persistent consumer kernel: a block expands its own 128 pairs
 * into the context's byte staging (L2-resident) and scores them.  Calls on different streams share that staging:
 * a later call waits, on the device, for the earlier one. */
int swb200_score_batch_packed_device(swb200_ctx* ctx, int device_index,
                                     const uint8_t* d_seq1_packed, const uint8_t* d_seq2_packed,
                                     const int8_t* score_matrix, int8_t gap_penalty,
                                     int32_t* d_scores, uint64_t n, void* cuda_stream);

/* Length sweep (BASELINE.json configs[3]): the same scoring for square pairs of seq_len bases,
 * seq_len in {128, 256, 512}; arrays are [n][seq_len].  The oracle is the scalar recurrence of
 * source.cpp:35-60 restated for any length.  SWB200_ERR_DOMAIN when seq_len * max(score_matrix)
 * does not fit the packed int16 arithmetic (possible only at 512). */
int swb200_score_batch_len(swb200_ctx* ctx, int seq_len, const uint8_t* seq1, const uint8_t* seq2,
                           const int8_t* score_matrix, int8_t gap_penalty, int32_t* scores, uint64_t n);
int swb200_score_batch_len_device(swb200_ctx* ctx, int device_index, int seq_len,
                                  const uint8_t* d_seq1, const uint8_t* d_seq2,
                                  const int8_t* score_matrix, int8_t gap_penalty,
                                  int32_t* d_scores, uint64_t n, void* cuda_stream);

/* Counts bytes > 3 in a DEVICE buffer (the reference never checks, source.cpp:50). */
int swb200_validate_codes_device(swb200_ctx* ctx, int device_index, const uint8_t* d_codes,
                                 uint64_t n_bytes, uint64_t* n_bad, void* cuda_stream);

/* ---- adaptive-banded X-drop semi-global aligner --------------------
 * Replaces, for n pairs,
 *   std::pair<int, std::vector<std::pair<int,int>>>
 *   SemiGlobal_AdaptiveBanded_XDrop_111_32_70(const std::array<uint8_t,16384>& seq1,
 *                                             const std::array<uint8_t,16384>& seq2)     (source.cpp:1836-1976)
 * and its AVX2 forms _simd, _simd_mark2, _simd_mark3, _simd_mark4 (source.cpp:1978-2725), which the
 * reference asserts equal to it (source.cpp:2774-2784).  Fixed by the reference's name: match /
 * mismatch / gap = 1/1/1, band 32, X-drop 70; the alignment starts at (0,0) and ends at the best cell.
 * seq1, seq2: [n][seq_len] codes 0..3 (the reference's shape is seq_len = 16384; 1..32768 accepted).
 * Per pair: scores = the pair's .first; (end_y, end_x) = the last element of its traceback vector;
 * ops[p][0..n_ops[p]) = the traceback as moves in forward order from (0,0): 0 = diagonal (y+1,x+1),
 * 1 = down (y+1), 2 = right (x+1) -- the vector itself is the running sum of the moves, (0,0) first
 * (host/smith_waterman_b200.hpp rebuilds it).  ops is [n][2*seq_len]; ops and n_ops may both be NULL
 * (score and end cell only; the traceback is then skipped on the device as well).  Bytes of a row past
 * n_ops[p] are unspecified.
 * Large batches (two chunks of sm_count x 32 pairs or more, seq_len a multiple of 16) cross the PCIe link
 * four to a byte in both directions: host threads pack the sequences to the reference's 2-bit format
 * (source.cpp:1580-1583) and expand the 2-bit move strings into `ops`; the arrays need not be pinned.
 * SWB200_SG_PIPE=0 (environment) keeps the plain chunked copies. */
int swb200_semiglobal_xdrop_batch(swb200_ctx* ctx, const uint8_t* seq1, const uint8_t* seq2, int32_t seq_len, uint64_t n,
                                  int32_t* scores, int32_t* end_y, int32_t* end_x, int32_t* n_ops, uint8_t* ops);
/* The kernel launch alone on DEVICE arrays (same meaning), stream-ordered on `cuda_stream`.  Launches
 * of one device share its trace scratch: issue them on one stream, or order them yourself. */
int swb200_semiglobal_xdrop_batch_device(swb200_ctx* ctx, int device_index, const uint8_t* d_seq1, const uint8_t* d_seq2,
                                         int32_t seq_len, uint64_t n, int32_t* d_scores, int32_t* d_end_y, int32_t* d_end_x,
                                         int32_t* d_n_ops, uint8_t* d_ops, void* cuda_stream);

/* ---- introspection for the benchmark harness --------------------------------------- */

typedef struct swb200_kernel_info {
    int fast_path;          /* 1: anti-diagonal offset DP (2 instr/word); 0: general form */
    int regs_per_thread;
    int threads_per_block;
    int blocks_per_sm;      /* resident, by the occupancy calculator */
    int smem_bytes_per_block;
    int sm_count;
    int sm_clock_khz;       /* cudaDevAttrClockRate */
} swb200_kernel_info;

/* Which kernel swb200_score_batch* would launch for this matrix/gap, and its resources. */
int swb200_kernel_info_for(swb200_ctx* ctx, int device_index, const int8_t* score_matrix,
                           int8_t gap_penalty, swb200_kernel_info* info);

int swb200_kernel_info_len(swb200_ctx* ctx, int device_index, int seq_len, const int8_t* score_matrix,
                           int8_t gap_penalty, swb200_kernel_info* info);

/* ---- host-side 2-bit packing lanes of swb200_score_batch ------------------------------
 * The caller's byte codes carry 2 bits each but cross PCIe as 8.  For large batches
 * swb200_score_batch therefore runs, per GPU, one RAW lane (the caller's bytes, 256 B per pair)
 * and `threads_per_gpu` PACK lanes (a host core compresses a sub-chunk to the reference's
 * 2-bit layout, source.cpp:1580-1583, into pinned staging; 64 B per pair cross the link and
 * the device expands them).  All lanes draw pieces of 8192 pairs from one counter, so
 * the split adapts to the machine.  This is wire compression only: no score is ever computed
 * on the host.  threads_per_gpu: -1 = auto ((CPUs this process may run on - GPUs) / GPUs: the calling
 * thread of every GPU is its RAW lane; env SWB200_PACK_THREADS overrides), 0 = off (every pair
 * travels as bytes). */
int swb200_set_host_pack_threads(swb200_ctx* ctx, int threads_per_gpu);
/* Pairs sent packed / as bytes by host batches so far, and the lane count in effect. */
int swb200_host_pack_stats(const swb200_ctx* ctx, uint64_t* packed_pairs, uint64_t* raw_pairs, int* threads_per_gpu);
/* With threads_per_gpu = -1 the library also tunes which of its lanes take part: large byte-coded batches try, in turn,
 * all PACK lanes + the RAW lane, all PACK lanes with the calling thread packing too (no DMA), half the PACK lanes + the
 * RAW lane, and the RAW lane alone, and keep the fastest (what pays depends on whether the PCIe link, the host's cores or
 * its DRAM is the scarcer resource, i.e. on how many GPUs share the host).  lanes_in_use / raw_lane_in_use: what is
 * currently preferred on GPU device_index; pairs_per_s[4]: the throughput seen with each candidate in the order above
 * (0 = not tried yet). */
int swb200_host_pack_tuning(const swb200_ctx* ctx, int device_index, int* lanes_in_use, int* raw_lane_in_use, double* pairs_per_s);
/* The per-pair call's resident server (swb200_score_pair): how many server kernels this context has launched so far,
 * how many calls went through its doorbell, and what the server itself measured for the last one -- nanoseconds from
 * the poll that brought the request to the store of the result (the sweep of the 128 x 128 matrix by one warp).
 * Any pointer may be NULL. */
int swb200_pair_path_stats(const swb200_ctx* ctx, uint64_t* server_launches, uint64_t* doorbell_calls, uint32_t* last_sweep_ns);

/* The packer itself (inverse of the reference's `unpack`, source.cpp:1580-1583), for callers
 * that want to feed swb200_score_batch_packed: n_codes bytes (a multiple of 8) -> n_codes/4. */
int swb200_pack2bit_host(const uint8_t* codes, uint8_t* packed, uint64_t n_codes);
/* Replaces `void unpack(const uint8_t* src, uint8_t* dest, int n)` and its SIMD forms unpack_simd1..4
 * (source.cpp:1580-1774) on the host: codes[4i + j] = (packed[i] >> 2j) & 3 for n_codes codes (any count). */
int swb200_unpack2bit_host(const uint8_t* packed, uint8_t* codes, uint64_t n_codes);

/* The integer-ALU issue peak of GPU `device_index`, measured now: independent chains of VIADDMNMX.S16x2 (the hot
 * loop's own instruction) on every SM for about target_ms milliseconds; *tinstr_per_s = thread-level instructions per
 * second / 1e12.  This is the roofline denominator SURVEY.md 8(d) asks to be measured on the box. */
int swb200_measure_alu_peak(swb200_ctx* ctx, int device_index, double target_ms, double* tinstr_per_s, double* elapsed_ms);

/* Resources of the semi-global aligner's kernel (fast_path is 0 there). */
int swb200_semiglobal_kernel_info(swb200_ctx* ctx, int device_index, swb200_kernel_info* info);

/* Kernel launches issued by this context since creation (all GPUs). */
uint64_t swb200_launch_count(const swb200_ctx* ctx);

/* Test hook: 0 sends small host batches (<= 2048 pairs) through the throughput kernel's chunk pipeline instead of
 * the one-warp-per-pair latency kernels; 1 (default) restores the latter; 2 = the latency kernels with one launch per
 * call even for a single pair (no resident server). */
int swb200_set_latency_path(swb200_ctx* ctx, int on);

/* Test hook: 1 forces the general (non-offset) kernel even where the fast one is exact. */
int swb200_set_force_general(swb200_ctx* ctx, int on);

/* ---- synthetic inputs and checksum for the verification / benchmark harness ---------- */

/* Pairs [0,n) of the reference's own test stream (source.cpp:2944-2953: mt19937_64(seed),
 * one draw per base taken as draw>>62, a[i] and b[i] alternately).  Host arrays [n][128]. */
int swb200_gen_reference_stream(uint64_t seed, uint64_t n, uint8_t* seq1, uint8_t* seq2);

/* Pairs [first, first+n) of a counter-based stream (pair k depends only on (seed,k)), so
 * any index range can be generated by any thread, rank or shard.  Byte codes [n][128] or
 * the reference's 2-bit packing [n][32] (source.cpp:1580-1583). */
int swb200_gen_counter_pairs(uint64_t seed, uint64_t first, uint64_t n, uint8_t* seq1, uint8_t* seq2, int threads);
int swb200_gen_counter_pairs_packed(uint64_t seed, uint64_t first, uint64_t n,
                                    uint8_t* seq1_packed, uint8_t* seq2_packed, int threads);

/* Long related pairs in the manner of the reference's TestSemiGlobal (source.cpp:2750-2771): seq2 is seq1
 * with sub_pct % mismatches, ins_pct % insertions and del_pct % deletions (the reference uses 10/10/10),
 * counter-based like the stream above.  Host arrays [n][seq_len]. */
int swb200_gen_related_pairs(uint64_t seed, uint64_t first, uint64_t n, int32_t seq_len, int sub_pct, int ins_pct, int del_pct,
                             uint8_t* seq1, uint8_t* seq2, int threads);

/* Benchmark helper (bench.py `host_ceiling`): `threads` host threads stream-read contiguous shares of
 * [buf, buf+bytes) `passes` times; *bytes_per_s = bytes * passes / wall time.  A byte-coded host batch must be read
 * from host memory once, by a packing core or by the DMA engine: this is that roofline's CPU side. */
int swb200_host_read_bandwidth(const void* buf, uint64_t bytes, int threads, int passes, double* bytes_per_s);

/* FNV-1a-64 over int32 scores: h = 1469598103934665603; h = (h ^ (uint32)s) * 1099511628211. */
uint64_t swb200_fnv1a64_i32(const int32_t* scores, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */

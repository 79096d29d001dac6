// ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Compiles the UNMODIFIED reference translation unit where it lies
// (/root/reference/source.cpp, found through -I by oracle/Makefile) and exports
// its scalar and AVX2 kernels behind a C ABI, so that tests can pin
// oracle/sw_oracle.c and the CUDA path against the reference itself, and so that
// bench.py can time the reference's own CPU path (`--impl reference`,
// cpu_baseline.kind == "reference").  No reference source is copied into this
// repository: the only coupling is the #include below.  Built WITHOUT -DNDEBUG,
// as the reference's own asserts expect.
//
// Entry points wrap:
//   SmithWaterman        source.cpp:35-60     (scalar; the oracle of record)
//   SmithWaterman_simd   source.cpp:62-208    ... SmithWaterman_simd9 source.cpp:953-1071
//   unpack               source.cpp:1580-1583
//   mt19937_64(seed) + uniform_int_distribution<int>(0,3), a/b interleaved: source.cpp:2944-2953
#define main swref_reference_main
#include "source.cpp"
#undef main

#include <pthread.h>
#include <thread>

namespace {
using Seq = std::array<uint8_t, 128>;
using Mat = std::array<int8_t, 16>;
typedef int (*KernelFn)(const Seq&, const Seq&, const Mat&, const int8_t);

KernelFn pick(int variant)
{
    switch (variant) {
    case 0: return SmithWaterman;
    case 1: return SmithWaterman_simd;
    case 2: return SmithWaterman_simd2;
    case 3: return SmithWaterman_simd3;
    case 4: return SmithWaterman_simd4;
    case 5: return SmithWaterman_simd5;
    case 6: return SmithWaterman_simd6;
    case 7: return SmithWaterman_simd7;
    case 8: return SmithWaterman_simd8;
    case 9: return SmithWaterman_simd9;
    default: return nullptr;
    }
}

void run_range(KernelFn fn, const uint8_t* s1, const uint8_t* s2, const int8_t* sm, int gap,
               int32_t* out, uint64_t lo, uint64_t hi)
{
    Mat m;
    std::memcpy(m.data(), sm, 16);
    Seq a, b;
    for (uint64_t p = lo; p < hi; ++p) {
        std::memcpy(a.data(), s1 + p * 128, 128);
        std::memcpy(b.data(), s2 + p * 128, 128);
        out[p] = fn(a, b, m, (int8_t)gap);
    }
}
} // namespace

extern "C" {

// variant: 0 = scalar, 1..9 = SmithWaterman_simd .. SmithWaterman_simd9.  Returns -1 for an unknown variant.
int swref_score_batch(int variant, const uint8_t* seq1, const uint8_t* seq2, const int8_t* score_matrix,
                      int gap_penalty, int32_t* scores, uint64_t n, int threads)
{
    KernelFn fn = pick(variant);
    if (!fn) return -1;
    if (threads <= 1) {
        run_range(fn, seq1, seq2, score_matrix, gap_penalty, scores, 0, n);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back(run_range, fn, seq1, seq2, score_matrix, gap_penalty, scores,
                          n * (uint64_t)t / (uint64_t)threads, n * (uint64_t)(t + 1) / (uint64_t)threads);
    for (auto& th : pool) th.join();
    return 0;
}

// The reference's own way of timing (source.cpp:3047-3055): ONE pair, `calls` times.  Returns the last score.
int swref_score_repeat(int variant, const uint8_t* seq1, const uint8_t* seq2, const int8_t* score_matrix,
                       int gap_penalty, uint64_t calls)
{
    KernelFn fn = pick(variant);
    if (!fn) return -1;
    Mat m; std::memcpy(m.data(), score_matrix, 16);
    Seq a, b;
    std::memcpy(a.data(), seq1, 128);
    std::memcpy(b.data(), seq2, 128);
    volatile int score = 0;
    for (uint64_t i = 0; i < calls; ++i) score = fn(a, b, m, (int8_t)gap_penalty);
    return score;
}

// The reference's generator exactly as written (std::uniform_int_distribution), source.cpp:2944-2953.
void swref_reference_stream(uint64_t seed, uint64_t n, uint8_t* seq1, uint8_t* seq2)
{
    std::mt19937_64 rnd(seed);
    std::uniform_int_distribution<int> dna(0, 3);
    for (uint64_t p = 0; p < n; ++p)
        for (int i = 0; i < 128; ++i) {
            seq1[p * 128 + i] = (uint8_t)dna(rnd);
            seq2[p * 128 + i] = (uint8_t)dna(rnd);
        }
}

// source.cpp:1580-1583
void swref_unpack(const uint8_t* src32, uint8_t* dest128)
{
    std::array<uint8_t, 32> s;
    std::array<uint8_t, 128> d;
    std::memcpy(s.data(), src32, 32);
    unpack(s, d);
    std::memcpy(dest128, d.data(), 128);
}

// SmithWaterman_111 (source.cpp:1073-1103): fixed match/mismatch/gap = 1/1/1
int swref_111(const uint8_t* seq1, const uint8_t* seq2)
{
    Seq a, b;
    std::memcpy(a.data(), seq1, 128);
    std::memcpy(b.data(), seq2, 128);
    return SmithWaterman_111(a, b);
}

// SmithWaterman_8bit111simd (source.cpp:1105-1225): the 8-bit AVX2 kernel for the same fixed scoring
int swref_8bit111(const uint8_t* seq1, const uint8_t* seq2)
{
    Seq a, b;
    std::memcpy(a.data(), seq1, 128);
    std::memcpy(b.data(), seq2, 128);
    return SmithWaterman_8bit111simd(a, b);
}

// SmithWaterman_8b111x32mark1/2/3 (source.cpp:1227, 1299, 1383): 32 queries x one target, fixed 1/1/1
int swref_x32(int mark, const uint8_t* seq1_32x128, const uint8_t* seq2, int32_t* dest32)
{
    std::array<uint8_t, 128 * 32> a;
    std::array<uint8_t, 128> b;
    std::array<int, 32> d;
    std::memcpy(a.data(), seq1_32x128, 128 * 32);
    std::memcpy(b.data(), seq2, 128);
    int rc;
    switch (mark) {
    case 1: rc = SmithWaterman_8b111x32mark1(a, b, d); break;
    case 2: rc = SmithWaterman_8b111x32mark2(a, b, d); break;
    case 3: rc = SmithWaterman_8b111x32mark3(a, b, d); break;
    default: return -1;
    }
    for (int i = 0; i < 32; ++i) dest32[i] = d[i];
    return rc;
}

// ---- the adaptive-banded X-drop semi-global aligner ------------------------
// SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (source.cpp:1836-1976, scalar) and its AVX2 forms
// _simd (1978-2165), _simd_mark2 (2167-2353), _simd_mark3 (2355-2541), _simd_mark4 (2543-2725).
// They keep several MB of tables on the stack, so every call runs on a thread with a 64 MiB stack.
namespace {
using LongSeq = std::array<uint8_t, 16384>;
using SgResult = std::pair<int, std::vector<std::pair<int, int>>>;
typedef SgResult (*SgFn)(const LongSeq&, const LongSeq&);

SgFn pick_sg(int variant)
{
    switch (variant) {
    case 0: return SemiGlobal_AdaptiveBanded_XDrop_111_32_70;
    case 1: return SemiGlobal_AdaptiveBanded_XDrop_111_32_70_simd;
    case 2: return SemiGlobal_AdaptiveBanded_XDrop_111_32_70_simd_mark2;
    case 3: return SemiGlobal_AdaptiveBanded_XDrop_111_32_70_simd_mark3;
    case 4: return SemiGlobal_AdaptiveBanded_XDrop_111_32_70_simd_mark4;
    default: return nullptr;
    }
}

struct SgJob {
    SgFn fn; const uint8_t* s1; const uint8_t* s2; uint64_t lo, hi;
    int32_t* scores; int32_t* tb; int64_t tb_cap; int32_t* tb_len;   // tb: (y,x) pairs of pair `lo` only (may be null)
};

void* sg_thread(void* arg)
{
    SgJob* j = (SgJob*)arg;
    LongSeq* a = new LongSeq;
    LongSeq* b = new LongSeq;
    for (uint64_t p = j->lo; p < j->hi; ++p) {
        std::memcpy(a->data(), j->s1 + p * 16384, 16384);
        std::memcpy(b->data(), j->s2 + p * 16384, 16384);
        const SgResult r = j->fn(*a, *b);
        j->scores[p] = r.first;
        if (j->tb && p == j->lo) {
            *j->tb_len = (int32_t)r.second.size();
            for (int64_t k = 0; k < (int64_t)r.second.size() && k < j->tb_cap; ++k) {
                j->tb[2 * k] = r.second[k].first;
                j->tb[2 * k + 1] = r.second[k].second;
            }
        }
    }
    delete a;
    delete b;
    return nullptr;
}

int sg_run(std::vector<SgJob>& jobs)
{
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 64u << 20);
    std::vector<pthread_t> th(jobs.size());
    int rc = 0;
    for (size_t k = 0; k < jobs.size(); ++k) rc |= pthread_create(&th[k], &attr, sg_thread, &jobs[k]);
    for (size_t k = 0; k < jobs.size(); ++k) pthread_join(th[k], nullptr);
    pthread_attr_destroy(&attr);
    return rc;
}
} // namespace

// One pair: score and the traceback as (y,x) int pairs from (0,0) to the best cell.
int swref_semiglobal(int variant, const uint8_t* seq1, const uint8_t* seq2, int32_t* score,
                     int32_t* tb_yx, int64_t tb_cap, int32_t* tb_len)
{
    SgFn fn = pick_sg(variant);
    if (!fn) return -1;
    std::vector<SgJob> jobs(1);
    jobs[0] = SgJob{fn, seq1, seq2, 0, 1, score, tb_yx, tb_cap, tb_len};
    return sg_run(jobs);
}

// n pairs over `threads` threads (contiguous ranges), scores only: the CPU baseline of the aligner.
int swref_semiglobal_batch(int variant, const uint8_t* seq1, const uint8_t* seq2, uint64_t n, int32_t* scores, int threads)
{
    SgFn fn = pick_sg(variant);
    if (!fn || threads < 1) return -1;
    std::vector<SgJob> jobs((size_t)threads);
    for (int k = 0; k < threads; ++k)
        jobs[k] = SgJob{fn, seq1, seq2, n * k / threads, n * (k + 1) / threads, scores, nullptr, 0, nullptr};
    return sg_run(jobs);
}

// The inputs of TestSemiGlobal (source.cpp:2734-2771): iteration `it` of mt19937_64(seed) with
// dna(0,3) / dice(0,99): a = iid, b = a with 10 % mismatch, 10 % insert, 10 % delete.  a, b: [n][16384].
void swref_semiglobal_test_inputs(uint64_t seed, uint64_t n, uint8_t* a_out, uint8_t* b_out)
{
    std::mt19937_64 rnd(seed);
    std::uniform_int_distribution<int> dna(0, 3);
    std::uniform_int_distribution<int> dice(0, 99);
    for (uint64_t it = 0; it < n; ++it) {
        uint8_t* a = a_out + it * 16384;
        uint8_t* b = b_out + it * 16384;
        for (int i = 0; i < 16384; ++i) a[i] = (uint8_t)dna(rnd);
        for (int i = 0, j = 0; i < 16384;) {
            if (j == 16384) b[i++] = (uint8_t)dna(rnd);
            else {
                const int p = dice(rnd);
                if (p < 10) { b[i++] = (uint8_t)dna(rnd); ++j; }
                else if (p < 20) { b[i++] = (uint8_t)dna(rnd); }
                else if (p < 30) { ++j; }
                else { b[i++] = a[j++]; }
            }
        }
    }
}

// The single pair of SpeedtestSemiGlobal (source.cpp:2804-2813): dice(0,19), 5 % substitutions.
void swref_semiglobal_speedtest_input(uint64_t seed, uint8_t* a, uint8_t* b)
{
    std::mt19937_64 rnd(seed);
    std::uniform_int_distribution<int> dna(0, 3);
    std::uniform_int_distribution<int> dice(0, 19);
    for (int i = 0; i < 16384; ++i) {
        a[i] = (uint8_t)dna(rnd);
        if (dice(rnd)) b[i] = a[i];
        else b[i] = (uint8_t)dna(rnd);
    }
}

int swref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

} // extern "C"

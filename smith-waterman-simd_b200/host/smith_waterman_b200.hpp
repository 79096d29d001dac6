// smith_waterman_b200.hpp -- C++ host side above the C ABI (include/swb200.h), mirroring the
// reference's interface for the hot path so that the new functions can sit next to
// SmithWaterman_simd .. SmithWaterman_simd9 in the reference's own harnesses
// (TestSimdSmithWaterman source.cpp:2943-2982, SpeedTest source.cpp:3032-3147).
//
//   int SmithWaterman_b200(seq1, seq2, score_matrix, gap_penalty)
//       -- the exact signature of source.cpp:462-466; forwards to swb200_score_pair.
//   void SmithWaterman_b200_batch(seq1s, seq2s, score_matrix, gap_penalty, dest)
//       -- the batched form (precedent: SmithWaterman_8b111x32mark1, source.cpp:1227-1234:
//          row-major 128-mers in, results into `dest`).
//
// Error behaviour: the reference has none (no validation, no return codes).  Here a failed
// call throws std::runtime_error carrying the library's message -- there is no CPU fallback
// to hide behind.  Header-only; link with -lswb200.
#pragma once
#include <array>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/swb200.h"

namespace swb200 {

class Context {
public:
    explicit Context(int n_devices = 1)
    {
        const int rc = swb200_init(&ctx_, nullptr, n_devices);
        if (rc != SWB200_OK) throw std::runtime_error(std::string("swb200_init: ") + swb200_last_error(nullptr));
    }
    ~Context() { swb200_shutdown(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    swb200_ctx* get() const { return ctx_; }
    void check(int rc) const
    {
        if (rc != SWB200_OK) throw std::runtime_error(std::string("swb200: ") + swb200_last_error(ctx_));
    }
private:
    swb200_ctx* ctx_ = nullptr;
};

// One process-wide context on GPU 0, created on first use (the reference's kernels are
// stateless free functions; this keeps the call sites identical).
inline Context& default_context()
{
    static Context ctx(1);
    return ctx;
}

} // namespace swb200

// Drop-in for SmithWaterman_simdN (source.cpp:462-466): same arguments, same return value.
// The per-pair call serialises on a mutex; use the batch form for throughput.
inline int SmithWaterman_b200(
    const std::array<uint8_t, 128>& seq1,
    const std::array<uint8_t, 128>& seq2,
    const std::array<int8_t, 16>& score_matrix,
    const int8_t gap_penalty)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    swb200::Context& c = swb200::default_context();
    int32_t score = 0;
    c.check(swb200_score_pair(c.get(), seq1.data(), seq2.data(), score_matrix.data(), gap_penalty, &score));
    return score;
}

// Drop-in for SmithWaterman_111 (source.cpp:1073-1076) and SmithWaterman_8bit111simd
// (source.cpp:1105-1107): fixed match/mismatch/gap = 1/1/1, same two arguments.
inline int SmithWaterman_111_b200(
    const std::array<uint8_t, 128>& seq1,
    const std::array<uint8_t, 128>& seq2)
{
    swb200::Context& c = swb200::default_context();
    int32_t score = 0;
    c.check(swb200_score_batch_111(c.get(), seq1.data(), seq2.data(), &score, 1));
    return score;
}

// n pairs at once: seq1s[p], seq2s[p] are the reference's std::array<uint8_t,128>, which are
// contiguous 128-byte objects, so a vector of them is exactly the [n][128] layout of the ABI.
inline void SmithWaterman_b200_batch(
    const std::vector<std::array<uint8_t, 128>>& seq1s,
    const std::vector<std::array<uint8_t, 128>>& seq2s,
    const std::array<int8_t, 16>& score_matrix,
    const int8_t gap_penalty,
    std::vector<int>& dest,
    swb200::Context* ctx = nullptr)
{
    if (seq1s.size() != seq2s.size()) throw std::invalid_argument("SmithWaterman_b200_batch: size mismatch");
    static_assert(sizeof(std::array<uint8_t, 128>) == 128, "std::array<uint8_t,128> must be 128 contiguous bytes");
    static_assert(sizeof(int) == sizeof(int32_t), "int must be 32 bits");
    swb200::Context& c = ctx ? *ctx : swb200::default_context();
    dest.resize(seq1s.size());
    c.check(swb200_score_batch(c.get(), seq1s.empty() ? nullptr : seq1s[0].data(), seq2s.empty() ? nullptr : seq2s[0].data(),
                               score_matrix.data(), gap_penalty, reinterpret_cast<int32_t*>(dest.data()), seq1s.size()));
}

// Drop-in for SmithWaterman_8b111x32mark1/2/3 (source.cpp:1227-1230, 1299, 1383): 32 row-major
// 128-mers against one target, results in dest, return value dest[0] (as the reference does,
// "for no deep reason", source.cpp:1233).  The reference fixes match/mismatch/gap = 1/1/1
// there (source.cpp:1238); this form takes them as defaulted arguments.
inline int SmithWaterman_b200_x32(
    const std::array<uint8_t, 128 * 32>& seq1,
    const std::array<uint8_t, 128>& seq2,
    std::array<int, 32>& dest,
    const std::array<int8_t, 16>& score_matrix = {1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1},
    const int8_t gap_penalty = 1)
{
    swb200::Context& c = swb200::default_context();
    c.check(swb200_score_one_vs_many(c.get(), seq1.data(), seq2.data(), score_matrix.data(), gap_penalty,
                                     reinterpret_cast<int32_t*>(dest.data()), 32));
    return dest[0];
}

// Drop-in for SemiGlobal_AdaptiveBanded_XDrop_111_32_70 and its AVX2 forms (source.cpp:1836-1838,
// 1978, 2167, 2355, 2543): same arguments, same pair<score, traceback> -- the traceback runs from
// (0,0) to the best cell, one (y, x) per step, rebuilt here from the library's move string.
#include <utility>
inline std::pair<int, std::vector<std::pair<int, int>>> SemiGlobal_AdaptiveBanded_XDrop_111_32_70_b200(
    const std::array<uint8_t, 16384>& seq1,
    const std::array<uint8_t, 16384>& seq2)
{
    swb200::Context& c = swb200::default_context();
    int32_t score = 0, end_y = 0, end_x = 0, n_ops = 0;
    std::vector<uint8_t> ops(2 * 16384);
    c.check(swb200_semiglobal_xdrop_batch(c.get(), seq1.data(), seq2.data(), 16384, 1, &score, &end_y, &end_x, &n_ops, ops.data()));
    std::vector<std::pair<int, int>> traceback;
    traceback.reserve((size_t)n_ops + 1);
    int y = 0, x = 0;
    traceback.emplace_back(y, x);
    for (int k = 0; k < n_ops; ++k) {
        if (ops[k] != 2) ++y;
        if (ops[k] != 1) ++x;
        traceback.emplace_back(y, x);
    }
    return std::make_pair((int)score, std::move(traceback));
}

"""Generates tests/golden/counter_stream_sums.json: the sum of the reference's scores over prefixes [0, N) of the
counter-based pair stream (swb200.counter_pairs, seed 10000), N up to 100 000 000 = BASELINE.json configs[2].

    python tests/golden/make_counter_sums.py        (authoring container: needs oracle/_ref; about two minutes on 8 cores)

Every score comes from the UNMODIFIED reference's SmithWaterman_simd9 (source.cpp:953-1071) compiled by oracle/Makefile;
the first 200 000 pairs are also scored by its scalar SmithWaterman (source.cpp:35-60) and simd4 and must agree.
A sum is independent of how the index range is cut into batches, ranks or GPUs, so it pins the streaming / sharded
configurations at their full size (bench.py --workload stream, tests/test_zz_fullsize_pins_gpu.py)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
from oracle import oracle as O  # noqa: E402
import swb200  # noqa: E402  (the generator only: a host function of the library, no GPU)

CHECKPOINTS = [1_000_000, 2_000_000, 8_000_000, 10_000_000, 16_000_000, 50_000_000, 100_000_000]
SETTINGS = {"speedtest_10_-30_15": (swb200.MATRIX_SPEEDTEST, 15), "x32_1_-1_1": (swb200.MATRIX_111, 1)}


def main():
    assert O.have_ref(), "oracle/_ref/libswref.so is needed (make -C oracle in the authoring container)"
    threads = os.cpu_count() or 1
    chunk = 1_000_000
    a = np.empty((chunk, 128), np.uint8)
    b = np.empty((chunk, 128), np.uint8)
    sums = {k: {} for k in SETTINGS}
    blocks = {k: {} for k in SETTINGS}      # pairs [r*1M, (r+1)*1M), r = 0..7: what rank r of `bench.py --gpus N` scores (r >= 1)
    run = {k: 0 for k in SETTINGS}
    t0 = time.time()
    for first in range(0, CHECKPOINTS[-1], chunk):
        swb200.counter_pairs(first, chunk, out=(a, b), threads=threads)
        if first == 0:
            an, bn = swb200.counter_pairs_numpy(0, 4096)      # the numpy restatement of the generator
            assert np.array_equal(a[:4096], an) and np.array_equal(b[:4096], bn)
        for key, (m, g) in SETTINGS.items():
            if key != "speedtest_10_-30_15" and first >= 10_000_000:
                continue                                       # the second scoring only up to 10 M pairs
            s = O.ref_score_batch(9, a, b, m, g, threads=threads)
            if first < 200_000:
                n = 200_000
                assert np.array_equal(s[:n], O.ref_score_batch(0, a[:n], b[:n], m, g, threads=threads))
                assert np.array_equal(s[:n], O.ref_score_batch(4, a[:n], b[:n], m, g, threads=threads))
            run[key] += int(s.sum(dtype=np.int64))
            if first < 8 * chunk:
                blocks[key][str(first // chunk)] = int(s.sum(dtype=np.int64))
            if first + chunk in CHECKPOINTS:
                sums[key][str(first + chunk)] = run[key]
        if (first // chunk) % 10 == 9:
            print(f"{first + chunk} pairs, {time.time() - t0:.0f} s", flush=True)
    out = {"stream": "swb200.counter_pairs(first=0, seed=10000): pair k from splitmix64 of (seed, 8k..8k+7), 32 bases per draw",
           "scored_by": "SmithWaterman_simd9 of the unmodified reference (oracle/_ref), first 200000 pairs also scalar and simd4",
           "sum_of_scores_over_prefix": sums, "sum_of_scores_block_1M": blocks}
    with open(os.path.join(HERE, "counter_stream_sums.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(sums), json.dumps(blocks))


if __name__ == "__main__":
    main()

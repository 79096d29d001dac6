"""Times swb200_semiglobal_xdrop_batch end to end at several batch sizes (development tool)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "smith-waterman-simd_b200"))
import swb200

ctx = swb200.Context(devices=[0])
N = 9472
pa, pb = swb200.PinnedArray((N, 16384), np.uint8), swb200.PinnedArray((N, 16384), np.uint8)
swb200.related_pairs(0, N, 16384, out=(pa.array, pb.array))
meta = [swb200.PinnedArray((N,), np.int32) for _ in range(4)]
ops = swb200.PinnedArray((N, 32768), np.uint8)
for n in (1184, 2368, 4736, 9472):
    for tb in (False, True):
        def call():
            ctx._check(ctx._lib.swb200_semiglobal_xdrop_batch(ctx._h, pa.array.ctypes.data, pb.array.ctypes.data, 16384, n,
                       meta[0].array.ctypes.data, meta[1].array.ctypes.data, meta[2].array.ctypes.data,
                       meta[3].array.ctypes.data if tb else None, ops.array.ctypes.data if tb else None))
        call(); call()
        t = time.perf_counter()
        for _ in range(5): call()
        dt = (time.perf_counter() - t) / 5
        print(json.dumps({"pairs": n, "traceback": tb, "ms": dt * 1e3, "pairs_per_s": n / dt}), flush=True)
if os.environ.get("SWB200_SG_TIMELINE"):
    pass

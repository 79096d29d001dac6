import os, sys, time, json
import numpy as np
sys.path.insert(0, "/root/repo/smith-waterman-simd_b200")
import swb200
ctx = swb200.Context(devices=[0])
N = 16384
pa, pb = swb200.PinnedArray((N, 16384), np.uint8), swb200.PinnedArray((N, 16384), np.uint8)
swb200.related_pairs(0, N, 16384, out=(pa.array, pb.array))
meta = [swb200.PinnedArray((N,), np.int32) for _ in range(4)]
ops = swb200.PinnedArray((N, 32768), np.uint8)
def call():
    ctx._check(ctx._lib.swb200_semiglobal_xdrop_batch(ctx._h, pa.array.ctypes.data, pb.array.ctypes.data, 16384, N,
               meta[0].array.ctypes.data, meta[1].array.ctypes.data, meta[2].array.ctypes.data, meta[3].array.ctypes.data, ops.array.ctypes.data))
call(); call()
t=time.perf_counter(); call(); print("ms", (time.perf_counter()-t)*1e3)

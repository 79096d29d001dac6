"""Host 2-bit packing rate (GB/s of byte-coded input) vs thread count -- development tool.
ctypes releases the GIL during the call, so Python threads measure the C++ packer itself."""
import json, os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "smith-waterman-simd_b200"))
import swb200

lib = swb200.load_library()
n = 1_000_000
a = np.random.default_rng(1).integers(0, 4, (n, 128), dtype=np.uint8)
out = np.empty((n, 32), np.uint8)
for threads in (1, 2, 4, 8, 12, 15, 16, 24, 32):
    if threads > (os.cpu_count() or 1):
        break
    bounds = [n * k // threads for k in range(threads + 1)]
    def work(k):
        lo, hi = bounds[k], bounds[k + 1]
        sub = 16384
        for c in range(lo, hi, sub):
            m = min(sub, hi - c)
            lib.swb200_pack2bit_host(a[c:c + m].ctypes.data, out[c:c + m].ctypes.data, m * 128)
    best = 1e9
    for rep in range(4):
        ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"threads": threads, "ms_per_128MB": best * 1e3, "GBps_in": n * 128 / best / 1e9}), flush=True)

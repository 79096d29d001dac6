"""Experiment: the persistent consumer kernel PULLING 2-bit input straight from pinned host memory (zero-copy loads over PCIe)
instead of being fed by copy-engine transfers.  Calls swb200_score_batch_packed_device with HOST pointers (pinned memory is
mapped on the device under unified addressing).  Prints ms per 1 M pairs; scores are checked against the resident run."""
import ctypes as C, os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import swb200

n = 1_000_000
ctx = swb200.Context(devices=[0])
a, b = swb200.reference_stream(n)
ka, kb = swb200.PinnedArray((n, 32), np.uint8), swb200.PinnedArray((n, 32), np.uint8)
ka.array[...] = swb200.pack2bit(a); kb.array[...] = swb200.pack2bit(b)
hs = swb200.PinnedArray((n,), np.int32)
m = np.asarray(swb200.MATRIX_SPEEDTEST, np.int8)
lib, h = ctx._lib, ctx._h
d_s = torch.empty(n, dtype=torch.int32, device="cuda")
d_ka, d_kb = torch.from_numpy(ka.array).cuda(), torch.from_numpy(kb.array).cuda()
stream = torch.cuda.current_stream().cuda_stream

def run(p1, p2, out, reps=10):
    for _ in range(3):
        rc = lib.swb200_score_batch_packed_device(h, 0, p1, p2, m.ctypes.data, 15, out, n, stream); assert rc == 0, rc
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.swb200_score_batch_packed_device(h, 0, p1, p2, m.ctypes.data, 15, out, n, stream)
        torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps

res = {}
res["resident_in_resident_out_ms"] = run(d_ka.data_ptr(), d_kb.data_ptr(), d_s.data_ptr())
ref = d_s.cpu().numpy().copy()
res["resident_in_host_out_ms"] = run(d_ka.data_ptr(), d_kb.data_ptr(), hs.array.ctypes.data)
res["host_out_equal"] = bool(np.array_equal(hs.array, ref))
res["host_in_resident_out_ms"] = run(ka.array.ctypes.data, kb.array.ctypes.data, d_s.data_ptr())
res["host_in_equal"] = bool(np.array_equal(d_s.cpu().numpy(), ref))
hs.array[:] = 0
res["host_in_host_out_ms"] = run(ka.array.ctypes.data, kb.array.ctypes.data, hs.array.ctypes.data)
res["host_in_host_out_equal"] = bool(np.array_equal(hs.array, ref))
res["fnv_ok"] = f"{swb200.fnv1a64(ref):016x}" == "ae56a1e6a1d57492"
print(json.dumps(res))

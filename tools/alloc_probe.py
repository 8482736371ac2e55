"""One-off probe: first-touch cost of cudaMalloc vs cudaMallocAsync for multi-GB scratch (development aid)."""
import ctypes as C
import time

rt = C.CDLL("libcudart.so.12")
rt.cudaFree(0)
p = C.c_void_p()


def t(f):
    rt.cudaDeviceSynchronize()
    t0 = time.perf_counter()
    f()
    rt.cudaDeviceSynchronize()
    return (time.perf_counter() - t0) * 1e3


for gb in (1, 4):
    n = C.c_size_t(gb << 30)
    ms = t(lambda: rt.cudaMalloc(C.byref(p), n))
    ms_f = t(lambda: rt.cudaFree(p))
    print(f"cudaMalloc {gb} GB: {ms:.2f} ms, cudaFree {ms_f:.2f} ms")
    ms = t(lambda: rt.cudaMallocAsync(C.byref(p), n, None))
    ms_f = t(lambda: rt.cudaFreeAsync(p, None))
    print(f"cudaMallocAsync {gb} GB (first, pool grows): {ms:.2f} ms, free {ms_f:.2f} ms")
    ms = t(lambda: rt.cudaMallocAsync(C.byref(p), n, None))
    ms_f = t(lambda: rt.cudaFreeAsync(p, None))
    print(f"cudaMallocAsync {gb} GB (again, default threshold): {ms:.2f} ms, free {ms_f:.2f} ms")
# many medium allocations
ptrs = [C.c_void_p() for _ in range(40)]
ms = t(lambda: [rt.cudaMalloc(C.byref(q), C.c_size_t(100 << 20)) for q in ptrs])
print(f"40 x cudaMalloc 100 MB: {ms:.2f} ms")
ms = t(lambda: [rt.cudaFree(q) for q in ptrs])
print(f"40 x cudaFree: {ms:.2f} ms")

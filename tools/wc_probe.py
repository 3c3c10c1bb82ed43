"""H2D bandwidth from ordinary pinned vs write-combined pinned host memory (cudaHostAlloc flags 0 / 4), with and without a concurrent D2H."""
import ctypes as C, torch, time
rt = C.CDLL("libcudart.so.12")
n = 512 * 640 * 480
d = torch.empty(n, dtype=torch.uint8, device="cuda"); do = torch.empty(34_000_000, dtype=torch.uint8, device="cuda")
ho = torch.empty(34_000_000, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for flags, name in ((0, "pinned"), (4, "write-combined")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 7, n)
    for both in (False, True):
        torch.cuda.synchronize(); t = time.perf_counter()
        for r in range(8):
            c = n // 8
            for k in range(8):
                rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr() + k * c), C.c_void_p(p.value + k * c), C.c_size_t(c), 1, C.c_void_p(s1.cuda_stream))
            if both:
                with torch.cuda.stream(s2):
                    ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 8
        print("%-15s %s  H2D %.1f GB/s" % (name, "with D2H" if both else "alone   ", n / dt / 1e9))
    rt.cudaFreeHost(p)

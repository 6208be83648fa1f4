"""H2D bandwidth of a 272.6 MB buffer (one bench step of fp32 descriptors): pinned vs write-combined pinned, one copy vs two
concurrent halves on two streams."""
import ctypes as C, time, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL([l.split()[-1] for l in open("/proc/self/maps") if "libcudart" in l][0])
n = 272630280
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    return p
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def copy(p, parts):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for rep in range(10):
        if parts == 1:
            rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, C.c_size_t(n), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        else:
            h = n // 2
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, C.c_size_t(h), 1, C.c_void_p(s1.cuda_stream))
            rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + h), C.c_void_p(p.value + h), C.c_size_t(n - h), 1, C.c_void_p(s2.cuda_stream))
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); e1.synchronize()
    return 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
for name, flags in (("pinned", 0), ("write-combined", 4)):
    p = alloc(flags)
    for parts in (1, 2):
        copy(p, parts)
        print(f"{name:15s} {parts} stream(s): {copy(p, parts):.1f} GB/s")

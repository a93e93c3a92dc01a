"""Developer tool: what the PCIe link of this box gives the host-buffer e2e path -- one contiguous
H2D copy of the C2 logits (49 MB, pinned) against the column-sliced cudaMemcpy2DAsync the session
issues (n slices of N / n utterances: rows of (N / n) * V * 4 bytes at a pitch of N * V * 4)."""
import ctypes, sys, time
import torch

rt = ctypes.CDLL("libcudart.so.12")
T, N, V = 1000, 256, 48
host = torch.randn(T, N, V).pin_memory()
dev = torch.empty(T, N, V, device="cuda")
st = torch.cuda.current_stream().cuda_stream
H2D = 1
pitch = N * V * 4


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def one_d():
    rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(host.data_ptr()),
                       ctypes.c_size_t(T * pitch), H2D, ctypes.c_void_p(st))


def two_d(n):
    def f():
        for k in range(n):
            b0, b1 = N * k // n, N * (k + 1) // n
            rt.cudaMemcpy2DAsync(ctypes.c_void_p(dev.data_ptr() + b0 * V * 4), ctypes.c_size_t(pitch),
                                 ctypes.c_void_p(host.data_ptr() + b0 * V * 4), ctypes.c_size_t(pitch),
                                 ctypes.c_size_t((b1 - b0) * V * 4), ctypes.c_size_t(T), H2D, ctypes.c_void_p(st))
    return f


t = timed(one_d)
print(f"1-D copy of {T * pitch / 1e6:.1f} MB: {t * 1e3:.3f} ms = {T * pitch / t / 1e9:.1f} GB/s")
for n in (1, 2, 4, 8, 16):
    t = timed(two_d(n))
    print(f"2-D copies, {n:2d} column slices: {t * 1e3:.3f} ms = {T * pitch / t / 1e9:.1f} GB/s")

"""Pure-write bandwidth of the box: torch fill kernels and cudaMemsetAsync on 1 GiB (is the voxel fill's 5.3 TB/s near the write ceiling?)."""
import torch, ctypes
n = 1 << 30
t = torch.empty(n, dtype=torch.uint8, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
def timed(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timed(lambda: t.zero_()); print(f"torch zero_   {ms*1e3:7.1f} us  {n/ms/1e9:.2f} TB/s")
ms = timed(lambda: t.fill_(1)); print(f"torch fill_   {ms*1e3:7.1f} us  {n/ms/1e9:.2f} TB/s")
s = torch.cuda.current_stream().cuda_stream
ms = timed(lambda: rt.cudaMemsetAsync(ctypes.c_void_p(t.data_ptr()), 0, ctypes.c_size_t(n), ctypes.c_void_p(s))); print(f"cudaMemsetAsync {ms*1e3:7.1f} us  {n/ms/1e9:.2f} TB/s")
u = torch.empty(n // 2, dtype=torch.uint8, device="cuda"); v = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
ms = timed(lambda: u.copy_(v)); print(f"torch copy 512 MiB {ms*1e3:7.1f} us  read+write {n/ms/1e9:.2f} TB/s")

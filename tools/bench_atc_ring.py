import os, sys, torch
sys.path.insert(0, '/root/repo')
from s2m2_b200.ops import CudaOps
ops = CudaOps()
g = torch.Generator().manual_seed(0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for S, N in ((4096, 176), (4096, 128), (4096, 304)):
    C = 256
    qkv = torch.randn((1, S, N, 3 * C), generator=g).cuda().half()
    o = torch.empty((1, S, N, C), dtype=torch.float16, device="cuda")
    t = timeit(lambda: ops.attention(qkv, o, S, N, 1, S // 2))
    fl = 4.0 * S * N * N * 256
    print(f"NST={os.environ.get('S2M2_ATC_NST','2')} S={S} N={N} d=256: {t:8.1f} us {fl/t/1e6:7.1f} TFLOP/s {(qkv.numel()+o.numel())*2/t/1e3:7.0f} GB/s", flush=True)

import torch, time
n = 6_400_000_000
a = torch.empty(n, dtype=torch.uint8, device='cuda')
b = torch.empty(n, dtype=torch.uint8, device='cuda')
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
t=timeit(lambda: a.zero_()); print('memset  %.3f ms  %.0f GB/s write' % (t, n/t/1e6))
t=timeit(lambda: a.fill_(7)); print('fill    %.3f ms  %.0f GB/s write' % (t, n/t/1e6))
t=timeit(lambda: b.copy_(a)); print('copy    %.3f ms  %.0f GB/s r+w' % (t, 2*n/t/1e6))
a32=a.view(torch.int32)
t=timeit(lambda: a32.sum()); print('sum     %.3f ms  %.0f GB/s read' % (t, n/t/1e6))
t=timeit(lambda: torch.max(a32)); print('max     %.3f ms  %.0f GB/s read' % (t, n/t/1e6))

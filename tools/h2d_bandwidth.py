import torch, time
x = torch.empty(43*1024*1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print('H2D', round(x.numel() / ms / 1e6, 2), 'GB/s', round(ms, 3), 'ms per 43 MiB')

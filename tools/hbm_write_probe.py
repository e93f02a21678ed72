import torch
x = torch.empty(1<<29, dtype=torch.float32, device='cuda')   # 2 GiB
src = torch.empty(1<<27, dtype=torch.float32, device='cuda')  # 0.5 GiB
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n*1e-3
s = t(lambda: x.fill_(1.5)); print("pure write  %.0f GB/s" % (x.numel()*4/s/1e9))
y = torch.empty_like(x)
s = t(lambda: y.copy_(x)); print("copy (r+w)  %.0f GB/s" % (2*x.numel()*4/s/1e9))
s = t(lambda: src.sum()); print("pure read   %.0f GB/s" % (src.numel()*4/s/1e9))
# 1 read : 3.6 write mix like layer 4 (0.54 GB read, 1.94 GB write): expand a small tensor into a large one
big = torch.empty(4, 1<<27, dtype=torch.float32, device='cuda')
s = t(lambda: big.copy_(src.unsqueeze(0).expand(4, -1))); print("1r:4w mix   %.0f GB/s" % ((src.numel()*4 + big.numel()*4)/s/1e9))

import torch, time
def t(f, n=10):
    f(); torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(n):
        a.record(); f(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
N=1<<30
x=torch.empty(N,dtype=torch.float32,device='cuda'); y=torch.empty(N,dtype=torch.float32,device='cuda')
print("fill  4GiB write-only  GB/s", 4*N/ t(lambda: x.fill_(1.0))/1e6)
print("zero  4GiB write-only  GB/s", 4*N/ t(lambda: x.zero_())/1e6)
print("copy  4+4GiB           GB/s", 8*N/ t(lambda: y.copy_(x))/1e6)
print("sum   4GiB read-only   GB/s", 4*N/ t(lambda: x.sum())/1e6)
xd=x.view(torch.float64)
print("fill f64 write-only GB/s", 4*N/ t(lambda: xd.fill_(1.0))/1e6)

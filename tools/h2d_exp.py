import torch, time
torch.cuda.init()
n=120_000_000
for wc in (False,):
    h=torch.empty(n,dtype=torch.uint8).pin_memory()
    d=torch.empty(n,dtype=torch.uint8,device='cuda')
    o=torch.empty(17_000_000,dtype=torch.uint8,device='cuda'); ho=torch.empty(17_000_000,dtype=torch.uint8).pin_memory()
    s=torch.cuda.Stream(); s2=torch.cuda.Stream()
    for _ in range(3): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(20): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
    print("H2D 120MB pinned: %.3f ms  %.1f GB/s"%(dt*1e3,n/dt/1e9))
    t=time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s): d.copy_(h,non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(o,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
    print("H2D 120MB + D2H 17MB concurrent: %.3f ms"%(dt*1e3))
    # chunked 8 x 15MB
    c=n//8
    t=time.perf_counter()
    for _ in range(20):
        for k in range(8): d[k*c:(k+1)*c].copy_(h[k*c:(k+1)*c],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
    print("H2D 8 chunks: %.3f ms"%(dt*1e3))
# two copy streams at once (does a second copy engine raise the H2D rate?)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
for _ in range(3):
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
print("H2D 120MB split over two streams: %.3f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))

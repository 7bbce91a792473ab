import torch, time
n=2<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,(src,dst) in {'d2h':(d,h),'h2d':(h,d)}.items():
    for _ in range(2): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(5): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize(); dt=time.time()-t
    print(name, 5*n/dt/1e9,'GB/s')
# concurrent both directions
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.time()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=time.time()-t
print('bidir each', 5*n/dt/1e9)

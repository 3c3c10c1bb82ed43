import torch, time
n=512*640*480
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
ho=torch.empty(34_000_000,dtype=torch.uint8).pin_memory(); do=torch.empty(34_000_000,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(both, chunks=8):
    torch.cuda.synchronize(); t=time.perf_counter()
    for r in range(5):
        with torch.cuda.stream(s1):
            c=n//chunks
            for k in range(chunks): d[k*c:(k+1)*c].copy_(h[k*c:(k+1)*c],non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                c=34_000_000//chunks
                for k in range(chunks): ho[k*c:(k+1)*c].copy_(do[k*c:(k+1)*c],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print('both' if both else 'h2d ', chunks, 'ms per 157MB', dt*1e3, 'GB/s', n/dt/1e9)
for b in (False,True):
    for ch in (1,8,16): run(b,ch)

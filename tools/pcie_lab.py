import torch, time
n=232965; D=256
hx=torch.empty(n,D,pin_memory=True); hg=torch.empty(n,D,pin_memory=True); ho=torch.empty(n,D,pin_memory=True)
dx=torch.empty(n,D,device='cuda'); dg=torch.empty(n,D,device='cuda'); do=torch.empty(n,D,device='cuda')
def run(streams_h2d, chunks, with_d2h):
    ss=[torch.cuda.Stream() for _ in range(streams_h2d)]; sd=torch.cuda.Stream()
    step=(n+chunks-1)//chunks
    def once():
        i=0
        for src,dst in ((hx,dx),(hg,dg)):
            for lo in range(0,n,step):
                with torch.cuda.stream(ss[i%len(ss)]):
                    dst[lo:lo+step].copy_(src[lo:lo+step],non_blocking=True)
                i+=1
        if with_d2h:
            with torch.cuda.stream(sd):
                for lo in range(0,n,step): ho[lo:lo+step].copy_(do[lo:lo+step],non_blocking=True)
    for _ in range(2): once()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): once()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print('h2d streams %d chunks %d d2h %d: %.2f ms  h2d %.1f GB/s'%(streams_h2d,chunks,with_d2h,dt*1e3, 2*n*D*4/dt/1e9))
for s,c,d in ((1,1,0),(1,8,0),(2,8,0),(1,8,1),(2,8,1),(2,2,1),(4,8,1)): run(s,c,d)

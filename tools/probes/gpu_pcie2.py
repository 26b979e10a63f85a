import torch, time
torch.cuda.init()
n=64; sz=64<<20
src=[torch.empty(sz,dtype=torch.uint8,device='cuda') for _ in range(8)]
dst=[torch.empty(sz,dtype=torch.uint8).pin_memory() for _ in range(n)]
def run(streams):
    ss=[torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize(); t=time.time()
    for rep in range(3):
        for i in range(n):
            with torch.cuda.stream(ss[i%streams]):
                dst[i].copy_(src[i%8],non_blocking=True)
    torch.cuda.synchronize(); dt=time.time()-t
    return 3*n*sz/dt/1e9
for s in (1,2,4): print("D2H streams",s, round(run(s),1),"GB/s")
# with concurrent compute
a=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16)
def busy():
    for _ in range(200): torch.matmul(a,a)
cs=torch.cuda.Stream()
with torch.cuda.stream(cs): busy()
print("D2H with matmul running", round(run(2),1))
torch.cuda.synchronize()
# with concurrent H2D
h=torch.empty(128<<20,dtype=torch.uint8).pin_memory(); d=torch.empty(128<<20,dtype=torch.uint8,device='cuda')
hs=torch.cuda.Stream()
with torch.cuda.stream(hs):
    for _ in range(100): d.copy_(h,non_blocking=True)
print("D2H with H2D running", round(run(2),1))

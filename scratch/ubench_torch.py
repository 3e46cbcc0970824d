import torch, time
torch.backends.cuda.matmul.allow_tf32=False
for n in (4096, 8192):
    a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
    for _ in range(2): c=a@b
    torch.cuda.synchronize()
    best=1e9
    for _ in range(5):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); c=a@b; e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(f"cuBLAS DGEMM {n}^3: {best:.3f} ms {2*n**3/best*1e-9:.2f} TFLOP/s")
n=5200
a=torch.randn(n,n,dtype=torch.float64,device='cuda'); spd=a@a.T+n*torch.eye(n,dtype=torch.float64,device='cuda')
for _ in range(2): L=torch.linalg.cholesky(spd)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); L=torch.linalg.cholesky(spd); e1.record(); torch.cuda.synchronize()
print(f"cuSOLVER potrf {n}: {e0.elapsed_time(e1):.3f} ms {n**3/3/e0.elapsed_time(e1)*1e-9:.2f} TFLOP/s")
x=torch.empty(1<<28,dtype=torch.float64,device='cuda'); y=torch.empty_like(x)
for _ in range(2): y.copy_(x)
torch.cuda.synchronize()
e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
print(f"copy 2GiB+2GiB: {2*x.numel()*8/e0.elapsed_time(e1)*1e-6:.1f} GB/s")

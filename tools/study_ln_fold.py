"""CPU numerics study for round 2: LayerNorm folded into the FOLLOWING GEMM.

    LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (b + W beta),   W' = W * diag(gamma)

so the GEMM can read the raw residual stream (no LayerNorm kernel, no extra pass over the activations) and its epilogue
applies two per-row scalars (mean, rstd: partial sums emitted by the PRODUCING GEMM's epilogue) and one per-column vector.
Prints the rel-L2 error vs an fp64 reference of today's path (LN in fp32 -> bf16 -> GEMM) and of the folded path.
Measured: 2.3e-3 vs 1.6e-3, independent of mean/std up to 10 - the folded path is the more accurate one."""
import torch
torch.manual_seed(0)
def bf(x): return x.to(torch.bfloat16).float()
M,K,N=2048,320,960
for mean_over_std in (0.0,0.5,2.0,10.0):
    x = bf(torch.randn(M,K)*1.7 + mean_over_std*1.7*torch.randn(M,1).sign())
    gamma = 1+0.2*torch.randn(K); beta=0.1*torch.randn(K)
    W = torch.randn(N,K)*K**-0.5; b=0.1*torch.randn(N)
    # exact (fp64)
    xd=x.double(); mu=xd.mean(1,keepdim=True); var=xd.var(1,unbiased=False,keepdim=True)
    ln=(xd-mu)/torch.sqrt(var+1e-5)*gamma.double()+beta.double()
    ref=ln@W.double().T+b.double()
    # path A: LN in fp32 -> bf16 -> GEMM (bf16 W) -> fp32 acc
    mu32=x.mean(1,keepdim=True); var32=x.var(1,unbiased=False,keepdim=True)
    lnA=bf((x-mu32)*torch.rsqrt(var32+1e-5)*gamma+beta)
    yA=lnA@bf(W).T+b
    # path B: folded
    Wp=bf(W*gamma[None,:]); colsum=Wp.sum(1); bp=b+W@beta   # b' uses unrounded W (fp32 vector)
    acc=x@Wp.T
    rstd=torch.rsqrt(var32+1e-5)
    yB=rstd*(acc-mu32*colsum[None,:])+bp
    e=lambda y:(y.double()-ref).norm()/ref.norm()
    print(f"mean/std {mean_over_std:4.1f}: LN->bf16->GEMM rel-L2 {e(yA):.2e} | folded {e(yB):.2e}")

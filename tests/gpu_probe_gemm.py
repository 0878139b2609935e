"""Bring-up probe (not a pytest): runs each GEMM layout in isolation under a watchdog and prints
error statistics, so one gpurun call localises descriptor / layout bugs."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import sys, torch
sys.path.insert(0, %r)
from cocodr_b200 import kernels as k
torch.manual_seed(0)
name = sys.argv[1]
def rnd(*s): return (torch.randn(*s)).half().cuda()
def rep(got, ref, tag):
    got, ref = got.float(), ref.float()
    err = (got-ref).abs()
    print(f"{tag}: maxerr={err.max().item():.4g} scale={ref.abs().max().item():.4g} frac_bad={(err > 0.05*ref.abs().max()).float().mean().item():.4f}", flush=True)
if name == "nt":
    for (M,N,K) in [(128,128,64),(128,128,128),(128,256,64),(256,512,256)]:
        a,b = rnd(M,K), rnd(N,K)
        out = torch.zeros(M,N,dtype=torch.float16,device="cuda")
        k.gemm(a,b,out,M=M,N=N,K=K); torch.cuda.synchronize()
        rep(out, a.float()@b.float().t(), f"nt {M}x{N}x{K}")
elif name == "nn":
    for (M,N,K) in [(128,128,64),(128,128,128),(128,256,64),(256,512,256)]:
        a,b = rnd(M,K), rnd(K,N)
        out = torch.zeros(M,N,dtype=torch.float16,device="cuda")
        k.gemm(a,b,out,M=M,N=N,K=K,b_major=1); torch.cuda.synchronize()
        rep(out, a.float()@b.float(), f"nn {M}x{N}x{K}")
elif name.startswith("tn"):
    lbo, sbo = (int(x) for x in name.split(":")[1:]) if ":" in name else (0,0)
    for (M,N,K) in [(128,128,64),(128,128,128),(256,256,256)]:
        a,b = rnd(K,M), rnd(K,N)
        out = torch.zeros(M,N,dtype=torch.float32,device="cuda")
        k.gemm(a,b,out,M=M,N=N,K=K,a_major=1,b_major=1,epilogue=k.EPI_F32_STORE,dbg_lbo=lbo,dbg_sbo=sbo); torch.cuda.synchronize()
        rep(out, a.float().t()@b.float(), f"tn[{lbo},{sbo}] {M}x{N}x{K}")
elif name == "perf":
    import time
    for (M,N,K) in [(16384,768,768),(16384,2304,768),(16384,3072,768),(16384,768,3072)]:
        a,b = rnd(M,K), rnd(N,K)
        out = torch.zeros(M,N,dtype=torch.float16,device="cuda")
        for _ in range(3): k.gemm(a,b,out,M=M,N=N,K=K)
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): k.gemm(a,b,out,M=M,N=N,K=K)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/20
        print(f"perf nt {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        t0=time.time()
        for _ in range(20): torch.matmul(a,b.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): torch.matmul(a,b.t())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/20
        print(f"     cublas {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
''' % ROOT

def main():
    cases = sys.argv[1:] or ["nt", "nn", "tn", "tn:8192:1024", "tn:1024:8192", "perf"]
    for c in cases:
        print(f"=== case {c}", flush=True)
        try:
            r = subprocess.run([sys.executable, "-c", CASE, c], timeout=120, capture_output=True, text=True)
            print(r.stdout[-3000:], r.stderr[-1500:], f"rc={r.returncode}", flush=True)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT (hang)", (e.stdout or b"")[-2000:], flush=True)

if __name__ == "__main__":
    main()

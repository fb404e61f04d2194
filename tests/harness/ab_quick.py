"""Quick same-box A/B of library builds: python tests/harness/ab_quick.py lib1.so lib2.so ...
Each build runs in its own process: N=8192 full and causal (B1 H32 D128), ~150 ms of back-to-back launches each."""
import os
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import flash_attention_cuda_b200 as fa
def tf(B,H,N,D,causal,iters):
    g = torch.Generator(device="cuda").manual_seed(0)
    q,k,v = ((torch.rand((B,H,N,D), device="cuda", generator=g)-0.5).half() for _ in range(3))
    o = torch.empty_like(q)
    for _ in range(5): fa.flash_attn_fwd(q,k,v,causal=bool(causal),out=o)
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fa.flash_attn_fwd(q,k,v,causal=bool(causal),out=o)
    e1.record(); torch.cuda.synchronize()
    return 4.0*B*H*N*N*D/(2 if causal else 1)/(e0.elapsed_time(e1)/iters)/1e9
print("full8192 %%7.1f  causal8192 %%7.1f  causal2048 %%7.1f  d64 %%7.1f" %% (tf(1,32,8192,128,0,150), tf(1,32,8192,128,1,300), tf(1,32,2048,128,1,1000), tf(32,16,2048,64,0,200)), fa.watchdog_status()["aborted"])
'''
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, FLASH_ATTN_B200_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CODE % REPO], env=env, capture_output=True, text=True)
        print(f"round {rnd} {os.path.basename(lib):24s} {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)

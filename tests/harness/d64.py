import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa
def tf(B,H,N,D,causal,iters):
    g=torch.Generator(device="cuda").manual_seed(0)
    q,k,v=((torch.rand((B,H,N,D),device="cuda",generator=g)-0.5).half() for _ in range(3)); o=torch.empty_like(q)
    for _ in range(5): fa.flash_attn_fwd(q,k,v,causal=bool(causal),out=o)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(iters): fa.flash_attn_fwd(q,k,v,causal=bool(causal),out=o)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/iters
    return 4.0*B*H*N*N*D/(2 if causal else 1)/ms/1e9
print(os.path.basename(fa.LIB_PATH), "cfg4 B32 H16 N2048 D64 full: %.1f | D64 N8192 H32 causal: %.1f | D128 N8192 causal %.1f" % (tf(32,16,2048,64,0,300), tf(1,32,8192,64,1,300), tf(1,32,8192,128,1,150)))

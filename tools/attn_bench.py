import torch, sys, json
sys.path.insert(0, '/root/repo')
from perceiverio_pytorch_b200 import engine, ops
sys.path.insert(0, '/root/repo/tools')
from chain_bench import chain_time
def run(B,H,N,d,fp16=False):
    engine.set_precision("fp16" if fp16 else "bf16")
    ld = 3*H*d
    qkv = torch.randn(B*N, ld, device='cuda').to(ops.dtype16())
    with torch.inference_mode():
        t = chain_time(lambda i: engine.attention(qkv, ld, 0, qkv, ld, H*d, qkv, ld, 2*H*d, B=B, H=H, Nq=N, Nk=N, dqk=d, dv=d, scale=d**-0.5), 20)
        o = engine.attention(qkv, ld, 0, qkv, ld, H*d, qkv, ld, 2*H*d, B=B, H=H, Nq=N, Nk=N, dqk=d, dv=d, scale=d**-0.5)
        q = qkv[:, :H*d].float().view(B,N,H,d).transpose(1,2); k = qkv[:, H*d:2*H*d].float().view(B,N,H,d).transpose(1,2); v = qkv[:, 2*H*d:].float().view(B,N,H,d).transpose(1,2)
        ref = torch.softmax(q @ k.transpose(-1,-2) * d**-0.5, -1) @ v
        ref = ref.transpose(1,2).reshape(B,N,H*d)
        err = float((o.float()[..., :H*d] - ref).abs().max() / ref.abs().max())
    engine.set_precision("bf16")
    return round(t,2), err
print(json.dumps({"tower_cls_64x8x512x128_us": run(64,8,512,128), "fp16": run(64,8,512,128,True), "flow_1x16x2048x32": run(1,16,2048,32), "mm_1x8x784x64": run(1,8,784,64), "lang_1x8x256_qk32": run(1,8,256,32)}))

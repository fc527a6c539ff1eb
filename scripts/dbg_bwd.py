import sys, torch, numpy as np
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load()
DEV="cuda:0"
def run(T, E=32, B=1, H=1):
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn(B,T,3*H*E, device=DEV, generator=g)
    d_out = torch.randn(B,T,H*E, device=DEV, generator=g)
    out = torch.empty(B,T,H*E, device=DEV); Tp=(T+127)//128*128
    lse = torch.zeros(B*H,Tp, device=DEV); d_qkv = torch.full((B,T,3*H*E), float('nan'), device=DEV)
    scratch = torch.empty(lib.v1t_attn_scratch_bytes(B,H,T,E), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    impl = _lib.IMPL_NAMES[sys.argv[1]] if len(sys.argv) > 1 else _lib.IMPL_BF16X3
    assert lib.v1t_attn_forward(qkv.data_ptr(),B,H,T,E,impl,0.0,0,0,out.data_ptr(),lse.data_ptr(),scratch.data_ptr(),st)==0
    assert lib.v1t_attn_backward(qkv.data_ptr(),out.data_ptr(),d_out.data_ptr(),lse.data_ptr(),B,H,T,E,impl,0.0,0,0,d_qkv.data_ptr(),scratch.data_ptr(),st)==0
    torch.cuda.synchronize()
    q64 = qkv.double().requires_grad_(True)
    q,k,v = [x.reshape(B,T,H,E).transpose(1,2) for x in q64.chunk(3,dim=-1)]
    p = torch.softmax(q@k.transpose(-1,-2)*E**-0.5, dim=-1)
    o = (p@v).transpose(1,2).reshape(B,T,H*E)
    o.backward(d_out.double())
    ref = q64.grad[...,:E][0].cpu().numpy(); got = d_qkv[...,:E][0].cpu().numpy()
    err = got-ref
    for nm, lo in (("dk", E), ("dv", 2*E)):
        r2 = q64.grad[..., lo:lo+E][0].cpu().numpy(); g2 = d_qkv[..., lo:lo+E][0].cpu().numpy()
        print(f"   {nm}: rel err {np.abs(g2-r2).max()/np.abs(r2).max():.3g}", end="")
    print()
    dp = (d_out.double().reshape(B,T,H,E).transpose(1,2) @ v.transpose(-1,-2))
    delta = (p*dp).sum(-1,keepdim=True)
    ds = (p*(dp-delta))*E**-0.5
    nb = (T+15)//16
    parts = [(ds[0,0][:, b*16:(b+1)*16] @ k[0,0][b*16:(b+1)*16]).detach().cpu().numpy() for b in range(nb)]
    A = np.stack([p_.ravel() for p_ in parts], 1)
    coef, *_ = np.linalg.lstsq(A, err.ravel(), rcond=None)
    resid = np.linalg.norm(err.ravel() - A@coef)/ (np.linalg.norm(err)+1e-30)
    print(f"T={T}: max err {np.abs(err).max():.3g}; lstsq coefficients of per-key-block contributions: {np.round(coef,3)} resid {resid:.3g}")
for T in (16, 64, 200):
    run(T)

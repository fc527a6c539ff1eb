import os, sys
sys.path.insert(0, ".")
sys.argv = ["x"]
exec(open("scripts/dbg_pair.py").read().split('print("---- detail')[0].split("for (B, H, T, E, p) in")[0])
for prec in ("0", "256"):
    os.environ["V1T_ATTN_PREC"] = prec
    for (B, H, T, E, p) in [(1, 1, 64, 64, 0.0), (1, 1, 200, 155, 0.0)]:
        a = run(B, H, T, E, p, "three")
        b = run(B, H, T, E, p, "pair")
        I = H * E
        print(f"== swap={prec} T{T} E{E}")
        for name, sl in (("dq", slice(0, I)), ("dk", slice(I, 2 * I)), ("dv", slice(2 * I, 3 * I))):
            x, y = a[0, :, sl], torch.nan_to_num(b[0, :, sl], nan=1e30, posinf=1e30, neginf=-1e30)
            print("  ", name, [f"{((x[:, c:c+32] - y[:, c:c+32]).abs().max() / x.abs().max()).item():.1e}" for c in range(0, E, 32)])

TAG=a CATSEG_DBG_SKIP_MLP=1 python tools/dbg_stages.py precise:swin_attn 1 1 2>&1 | tail -8 | head -2
TAG=b CATSEG_DBG_SKIP_MLP=1 CATSEG_ATTN_V=2 CATSEG_A2_FORCE_SPLIT=1 python tools/dbg_stages.py fast:swin_attn 1 1 2>&1 | tail -8 | head -2
TAG=c CATSEG_DBG_SKIP_MLP=1 python tools/dbg_stages.py precise:swin_mlp 1 1 2>&1 | tail -8 | head -2
python - <<'PY'
import torch
a,b,c=[torch.load("/tmp/taps_%s.pt"%t) for t in "abc"]
for n in ["swin_l0_b1"]:
    print(n, "precise-route vs fast-route:", (a[n]-b[n]).abs().max().item(), " precise-route vs exact-attn:", (a[n]-c[n]).abs().max().item(), " fast-route vs exact-attn:", (b[n]-c[n]).abs().max().item())
    d=(a[n]-c[n]).abs()[0,0]   # [576,128]
    print("rows with err>1e-3:", (d.max(1)[0]>1e-3).sum().item(), "cols:", (d.max(0)[0]>1e-3).sum().item())
PY

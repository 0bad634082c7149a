import sys, torch
sys.path.insert(0, "unet-segment-pytorch_b200")
from unet import kernels as K
N = 4
LAYERS = [("inc.3", 512, 64, 0, 64, 9), ("down1.0", 256, 64, 0, 128, 9), ("down1.3", 256, 128, 0, 128, 9),
    ("down2.0", 128, 128, 0, 256, 9), ("down2.3", 128, 256, 0, 256, 9), ("down3.0", 64, 256, 0, 512, 9),
    ("down3.3", 64, 512, 0, 512, 9), ("down4.0", 32, 512, 0, 512, 9), ("down4.3", 32, 512, 0, 512, 9), ("up1.0", 64, 512, 512, 512, 9),
    ("up1.3", 64, 512, 0, 256, 9), ("up2.0", 128, 256, 256, 256, 9), ("up2.3", 128, 256, 0, 128, 9),
    ("up3.0", 256, 128, 128, 128, 9), ("up3.3", 256, 128, 0, 64, 9), ("up4.0", 512, 64, 64, 64, 9),
    ("up4.3", 512, 64, 0, 64, 9), ("g1.Wx", 64, 512, 0, 256, 1), ("g1.Wg", 32, 512, 0, 256, 1), ("g2.Wx", 128, 256, 0, 128, 1), ("g2.Wg", 64, 256, 0, 128, 1),
    ("g3.Wx", 256, 128, 0, 64, 1), ("g3.Wg", 128, 128, 0, 64, 1), ("g4.Wx", 512, 64, 0, 32, 1), ("g4.Wg", 256, 64, 0, 32, 1)]
tot = 0
for name, h, c0, c1, cout, taps in LAYERS:
    x0 = torch.randn(N, h, h, c0, device="cuda").bfloat16()
    x1 = torch.randn(N, h, h, c1, device="cuda").bfloat16() if c1 else None
    dy = torch.randn(N, h, h, cout, device="cuda").bfloat16()
    p = K.conv_wgrad(x0, dy, taps, x1=x1)
    mb = p.numel() * 4 / 1e6
    tot += mb
    print(f"{name:8s} variant {K.last_conv_variant()} splits {p.shape[0]:4d} partial {mb:7.1f} MB (params {p.shape[1]*p.shape[2]*4/1e6:6.2f} MB)")
print("total partial MB", round(tot, 1))

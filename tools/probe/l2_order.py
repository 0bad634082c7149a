"""Probe: does a consumer that reads a just-written tensor in REVERSE order (most recently written part first) run
faster than one that reads it front to back?  The B200's 126 MB L2 holds roughly the tail of a 134 MB tensor."""
import torch
dev = "cuda"


def timeit(fn, iters=20):
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(e0, e1)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for mb in (34, 67, 100, 134, 200, 268):
    n = mb * (1 << 20) // 2 // 4
    x = torch.randn(4, n, device=dev).bfloat16()        # 4 slabs (images)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    idx_f = list(range(4))
    res = {}
    for name, order in (("forward", idx_f), ("reverse", idx_f[::-1])):
        def run(e0, e1):
            torch.mul(x, 2, out=y)                    # producer: writes y front to back
            e0.record()
            for i in order:                            # consumer: the slabs in the given order
                torch.mul(y[i], 2, out=z[i])
            e1.record()
        res[name] = timeit(run)
    print(f"{mb:4d} MB tensor: consumer forward {res['forward']:7.1f} us ({3 * mb * 1.048576 / res['forward'] * 1e0:5.2f} TB/s incl. write)  "
          f"reverse {res['reverse']:7.1f} us  ratio {res['forward'] / res['reverse']:.3f}")

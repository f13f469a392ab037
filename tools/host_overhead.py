"""Host-side cost of one forward call (python shim + C ABI replay of the launch list): timed on a frame small enough
that the GPU is never the bottleneck, and at full size with the queue kept short.   host_overhead.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmc_b200 as D  # noqa: E402

dev = torch.device("cuda:0")
for (H, W) in ((128, 192), (1280, 1920)):
    frames, masks = D.clips.synthetic_clip(5, 1, 3, H, W)
    x = torch.cat([frames, masks], 2).to(dev)
    torch.manual_seed(1)
    mp = D.build_p_model("performance").eval().to(dev)
    with torch.no_grad():
        r = mp(x[:, 1], 40, {"frame": x[:, 0, :3].contiguous(), "feature": None}, after_i=True)
        for _ in range(5):
            r = mp(x[:, 2], 32, r["dpb"], after_i=False)
        torch.cuda.synchronize()
        n = 100
        t0 = time.perf_counter()
        for i in range(n):
            r = mp(x[:, 1 + i % 2], 32, r["dpb"], after_i=False)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        # the C ABI call alone (same buffers every time)
        import ctypes
        lib = D._capi.load()
        h, stream = mp._engine(1, H, W, dev)
        xi, mk = x[:, 1, :3].contiguous(), x[:, 1, 3:4].contiguous()
        feat_in = r["dpb"]["feature"].contiguous()
        x_hat, feat, bpp3 = torch.empty_like(xi), torch.empty_like(feat_in), torch.empty(1, 3, device=dev)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        for i in range(6):                      # queue empty: pure host time of the call
            lib.dmc_forward(h, P(xi), P(mk), ctypes.c_void_p(), P(feat_in), 32, 0, P(x_hat), P(feat), P(bpp3),
                            ctypes.c_void_p(), ctypes.c_void_p(), stream)
        t6 = time.perf_counter()
        for i in range(6):
            r = mp(x[:, 1 + i % 2], 32, r["dpb"], after_i=False)
        t7 = time.perf_counter()
        print(f"{H}x{W}: host time with an empty queue: dmc_forward {1e3 * (t6 - t5) / 6:.3f} ms, module call {1e3 * (t7 - t6) / 6:.3f} ms")
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        for i in range(n):
            lib.dmc_forward(h, P(xi), P(mk), ctypes.c_void_p(), P(feat_in), 32, 0, P(x_hat), P(feat), P(bpp3),
                            ctypes.c_void_p(), ctypes.c_void_p(), stream)
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{H}x{W}: dmc_forward alone {1e3 * (t4 - t3) / n:.3f} ms per call (issue only)")
    print(f"{H}x{W}: host loop {1e3 * (t1 - t0) / n:.3f} ms per forward (issue only), {1e3 * (t2 - t0) / n:.3f} ms with the final sync")

"""Dev aid: pinned H2D / D2H rates alone and concurrently (what bounds bench.py's e2e number)."""
import time, torch
dev = torch.device("cuda:0")
for mb_in, mb_out in ((34, 33), (25, 20)):
    hin = torch.empty(mb_in << 20, dtype=torch.uint8).pin_memory(); din = torch.empty_like(hin, device=dev)
    dout = torch.empty(mb_out << 20, dtype=torch.uint8, device=dev); hout = torch.empty(mb_out << 20, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    def run(do_in, do_out, reps=50):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
    run(True, True, 5)
    a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"{mb_in} MB in / {mb_out} MB out: H2D {a:.3f} ms ({mb_in/a:.1f} GB/s)  D2H {b:.3f} ms ({mb_out/b:.1f} GB/s)  both {c:.3f} ms")

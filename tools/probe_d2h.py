"""D2H rate of the int64 mask batch (537 MB, pinned destination): one copy vs the same bytes split over several streams."""
import torch
dev = torch.device("cuda", 0)
src = torch.empty(64, 1024, 1024, dtype=torch.int64, device=dev).fill_(3)
dst = torch.empty(64, 1024, 1024, dtype=torch.int64).pin_memory()
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    chunks_s, chunks_d = src.chunk(parts), dst.chunk(parts)
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for st in streams:
            st.wait_event(a)
        for st, s, d in zip(streams, chunks_s, chunks_d):
            with torch.cuda.stream(st):
                d.copy_(s, non_blocking=True)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        b.record(); torch.cuda.synchronize()
        if rep: best = min(best, a.elapsed_time(b))
    print(f"{parts} stream(s): {best:.2f} ms = {src.numel() * 8 / best / 1e6:.1f} GB/s")
h = torch.empty(64, 3, 1024, 1024, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    torch.cuda.synchronize(); a.record(); d.copy_(h, non_blocking=True); b.record(); torch.cuda.synchronize()
print(f"H2D 201 MB: {a.elapsed_time(b):.2f} ms = {h.numel() / a.elapsed_time(b) / 1e6:.1f} GB/s")

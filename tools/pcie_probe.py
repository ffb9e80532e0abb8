"""Measures what the host link gives: H2D alone, D2H alone, and both directions at once (pinned memory, two streams).
The e2e numbers of bench.py (host buffers in, host buffers out) are bounded by the last figure."""
import json, torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); e1.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 2)
print(json.dumps({"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "each_way_when_both_gbs": run(True, True)}))

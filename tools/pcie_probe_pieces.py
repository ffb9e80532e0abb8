import json, torch
def run(piece, total=256<<20, reps=4):
    k = total // piece
    hi = [torch.empty(piece, dtype=torch.uint8).pin_memory() for _ in range(k)]
    ho = [torch.empty(piece, dtype=torch.uint8).pin_memory() for _ in range(k)]
    di = [torch.empty(piece, dtype=torch.uint8, device="cuda") for _ in range(k)]
    do = [torch.empty(piece, dtype=torch.uint8, device="cuda") for _ in range(k)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        for j in range(k):
            with torch.cuda.stream(s1): di[j].copy_(hi[j], non_blocking=True)
            with torch.cuda.stream(s2): ho[j].copy_(do[j], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); e1.synchronize()
    return total * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
run(16<<20, reps=1)
print(json.dumps({f"{p}MiB_pieces_each_way_gbs": round(run(p << 20), 1) for p in (2, 4, 8, 16, 32, 64, 256)}))

"""Shared timing helper of the probe tools."""


def time_kernel(torch, fn, nbuf, iters=200, warm=20):
    """Device time per call (us): the nbuf calls (rotating weight buffers, pool > L2) are captured into ONE CUDA graph
    so that Python / ctypes launch overhead is not what gets measured; CUDA events around graph replays."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(min(nbuf, 3)):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(nbuf):
            fn(i)
    reps = max(1, iters // nbuf)
    for _ in range(max(1, warm // nbuf)):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (reps * nbuf) * 1e3

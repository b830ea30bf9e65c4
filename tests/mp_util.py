"""Multi-process test helper."""


def run_ranks(worker, P, args, timeout=900):
    """Spawn P ranks; if one reports an error the others (blocked in recv) are terminated instead of waited for."""
    import queue
    import time

    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, P, *args, q)) for r in range(P)]
    for p in procs:
        p.start()
    res = {}
    t0 = time.time()
    try:
        while len(res) < P and time.time() - t0 < timeout:
            try:
                r, out = q.get(timeout=5)
            except queue.Empty:
                if any(p.exitcode not in (None, 0) for p in procs):
                    raise RuntimeError("a rank died") from None
                continue
            res[r] = out
            if "error" in out:
                raise RuntimeError(f"rank {r} failed:\n{out['error']}")
        if len(res) < P:
            raise TimeoutError("site-parallel ranks timed out")
        for p in procs:
            p.join(timeout=120)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    return res

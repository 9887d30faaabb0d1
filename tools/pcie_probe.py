"""Host<->device copy bandwidth of this box with N GPUs copying at the same time (development probe; names the fabric limit
behind the multi-GPU e2e and descriptor-download numbers).
  python tools/pcie_probe.py [--gpus 1,2,4,8]
One process per GPU (spawned here), 531 MB pinned buffers, H2D then D2H then both directions at once; prints the aggregate."""
import argparse
import multiprocessing as mp
import time


def worker(dev, barrier, q):
    import torch

    torch.cuda.set_device(dev)
    a = torch.empty((64, 1080, 1920), dtype=torch.float32).pin_memory()
    h = torch.empty((64, 1080, 1920), dtype=torch.float32).pin_memory()
    d = torch.empty_like(a, device="cuda")
    d2 = torch.empty_like(a, device="cuda")
    s2 = torch.cuda.Stream()
    nbytes = a.numel() * 4
    out = {}
    for name in ("h2d", "d2h", "both"):
        for it in range(2):
            barrier.wait()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(6):
                if name in ("h2d", "both"):
                    d.copy_(a, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h.copy_(d2, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            barrier.wait()
        out[name] = 6 * nbytes * (2 if name == "both" else 1) / dt / 1e9
    q.put((dev, out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    args = ap.parse_args()
    mp.set_start_method("spawn")
    for n in [int(x) for x in args.gpus.split(",")]:
        barrier, q = mp.Barrier(n), mp.Queue()
        ps = [mp.Process(target=worker, args=(i, barrier, q)) for i in range(n)]
        [p.start() for p in ps]
        res = [q.get() for _ in ps]
        [p.join() for p in ps]
        tot = {k: sum(r[1][k] for r in res) for k in ("h2d", "d2h", "both")}
        print(f"{n} GPUs at once: H2D {tot['h2d']:.1f} GB/s aggregate ({tot['h2d']/n:.1f} per GPU), D2H {tot['d2h']:.1f} ({tot['d2h']/n:.1f}), "
              f"both directions {tot['both']:.1f} ({tot['both']/n:.1f})", flush=True)

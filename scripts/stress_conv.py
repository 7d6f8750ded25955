"""One small conv (the self-probe's shape) against torch in P concurrent processes: where does it go wrong under contention?"""
import multiprocessing as mp, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(idx, n, q, direct):
    os.environ["SS4K_DESC_MODE"] = "0"
    os.environ["SS4K_SKIP_PROBE"] = "1"
    import torch
    import ss4k_b200
    eng = ss4k_b200.Engine.get(0)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(1, 64, 5, 200, generator=g) - 0.5).cuda()
    w = ((torch.rand(64, 64, 3, 3, generator=g) - 0.5) * 0.2).half().float().cuda()
    b = (torch.rand(64, generator=g) - 0.5).cuda()
    ref = torch.nn.functional.conv2d(x.half().float(), w, b, padding=1)
    out = []
    for i in range(n):
        y = eng.conv3x3(x, w, b, direct_f32=direct)
        torch.cuda.synchronize()
        e = (y - ref).abs()
        if e.max().item() > 0.02:
            bad = (e > 0.02).nonzero()
            out.append((i, round(e.max().item(), 4), bad.shape[0], sorted(set(bad[:, 1].tolist()))[:8], sorted(set(bad[:, 2].tolist())),
                        (min(bad[:, 3].tolist()), max(bad[:, 3].tolist()))))
    q.put((idx, len(out), out[:4]))


if __name__ == "__main__":
    P, N, direct = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=work, args=(i, N, q, bool(direct))) for i in range(P)]
    [p.start() for p in ps]
    for r in sorted(q.get(timeout=600) for _ in ps):
        print(r)
    [p.join() for p in ps]

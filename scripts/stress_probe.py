"""Start-up self-probe under contention: P processes create / destroy engines in a loop on the same GPU.  Prints the failures
per process.  (Found: uploads made with pageable cudaMemcpy / NULL-stream memsets were not ordered before kernels on the
engine's non-blocking stream -- wrong probe results in ~85 % of the starts with three processes; fixed by a device
synchronisation at the end of every set-up path.)"""
import ctypes, multiprocessing as mp, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(idx, n, q):
    import ss4k_b200
    from ss4k_b200 import _lib as L
    lib = L.load()
    bad = []
    for i in range(n):
        h = ctypes.c_void_p()
        rc = lib.ss4k_create(0, ctypes.byref(h))
        if rc != 0:
            bad.append((i, rc, (lib.ss4k_last_error(None) or b"").decode()[:200]))
        else:
            lib.ss4k_destroy(h)
    q.put((idx, bad))


if __name__ == "__main__":
    P, N = int(sys.argv[1]), int(sys.argv[2])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=work, args=(i, N, q)) for i in range(P)]
    [p.start() for p in ps]
    res = [q.get(timeout=600) for _ in ps]
    [p.join() for p in ps]
    for idx, bad in sorted(res):
        print("process", idx, "failures", len(bad), bad[:3])

import mmap, time, numpy as np, os
from concurrent.futures import ThreadPoolExecutor
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| defrag:", open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip(), "| cores", os.cpu_count())
n = 64
src = np.ones((2, 1400, 1400), np.float32)
def fresh(huge):
    nbytes = n * src.nbytes
    if huge:
        mm = mmap.mmap(-1, nbytes + (2 << 20))
        mm.madvise(mmap.MADV_HUGEPAGE)
        return np.frombuffer(mm, dtype=np.float32, count=n * src.size).reshape((n,) + src.shape), mm
    return np.empty((n,) + src.shape, np.float32), None
for huge in (False, True):
    for thr in (1, 3, 6):
        dst, keep = fresh(huge)
        pool = ThreadPoolExecutor(thr)
        t0 = time.perf_counter()
        list(pool.map(lambda i: np.copyto(dst[i], src), range(n)))
        dt = time.perf_counter() - t0
        print(f"huge={huge} threads={thr}: first-touch copy {dt/n*1e3:.3f} ms per 15.68 MB = {src.nbytes*n/dt/1e9:.1f} GB/s")
        t0 = time.perf_counter()
        list(pool.map(lambda i: np.copyto(dst[i], src), range(n)))
        dt = time.perf_counter() - t0
        print(f"            second pass      {dt/n*1e3:.3f} ms = {src.nbytes*n/dt/1e9:.1f} GB/s")
        pool.shutdown(); del dst, keep

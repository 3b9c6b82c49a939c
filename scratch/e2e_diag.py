import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in ("tests", "mappy-rs_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import bench
from mappy_rs import _mmg
lib = _mmg.Lib()
ref, coff, names, buf, offs = bench.workload(200000, 0)
io, mopt = _mmg.IdxOpt(), _mmg.MapOpt()
lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mopt)))
mopt.flag = 0
t0 = time.perf_counter()
idx = _mmg.Index.build(lib, io, names, [ref.tobytes()])
print("index build s", time.perf_counter() - t0)
lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mopt), idx.h))
al = _mmg.DeviceAligner(lib, idx, mopt, device=0)
hbuf = torch.empty(len(buf), dtype=torch.uint8, pin_memory=True)
hbuf.numpy()[:] = buf
hptr = hbuf.numpy()
offs = np.ascontiguousarray(offs)
for prof in (1, 0):
    al.set("profile", prof)
    for it in range(3):
        t = [time.perf_counter()]
        b = al.upload(hptr, offs); t.append(time.perf_counter())
        al.run(b); t.append(time.perf_counter())
        al.fetch(b); t.append(time.perf_counter())
        res = _mmg.Batch(lib, b, len(offs) - 1); t.append(time.perf_counter())
        al.free(b); t.append(time.perf_counter())
        print("profile", prof, "upload/run/fetch/Batch/free ms:", ["%.1f" % ((t[i + 1] - t[i]) * 1e3) for i in range(5)], "dev ms %.1f" % al.last_run_ms())
# pageable
for it in range(2):
    t0 = time.perf_counter(); r = al.map_batch(buf, offs); print("pageable map_batch ms %.1f" % ((time.perf_counter() - t0) * 1e3))
for it in range(2):
    t0 = time.perf_counter(); r = al.map_batch(hptr, offs); print("pinned map_batch ms %.1f" % ((time.perf_counter() - t0) * 1e3))

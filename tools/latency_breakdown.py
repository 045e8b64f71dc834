import time, numpy as np, cv2, sys, os
sys.path.insert(0, '/root/repo')
import deepcharuco_b200 as dc
g = np.load('/root/repo/tests/golden/sample_image.npz'); img = g['bgr']
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device='cuda')
for _ in range(10): dc.infer_image(img, 16, deepc, refinenet)
def t(fn, n=1000):
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
eng = deepc._ctx.engine(240, 320, max_batch=1)
print('cvtColor us', t(lambda: cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)))
print('engine.infer_batch_host us', t(lambda: eng.infer_batch_host(gray[None], 16, True)))
print('dc.infer_batch us', t(lambda: dc.infer_batch(gray[None], 16, deepc, refinenet)))
print('dc.infer_image us', t(lambda: dc.infer_image(img, 16, deepc, refinenet)))
import ctypes as C
from deepcharuco_b200 import _native as N
f = np.ascontiguousarray(gray[None])
def raw():
    N.lib().dcu_infer_batch_host(eng._h, f.ctypes.data, 1, 16, 1, eng._counts.ctypes.data, eng._offsets.ctypes.data, C.addressof(eng._total), eng._kpts.ctypes.data, eng._refined.ctypes.data, None)
print('raw C call us', t(raw))

import sys, ctypes, time, numpy as np, os, torch
sys.path.insert(0,'.')
from audio_transformers_b200 import _lib
lib=_lib.load()
rng=np.random.default_rng(0)
n=64
pin=torch.empty((n,480000),dtype=torch.float32,pin_memory=True)
for dt in (np.float64,np.float32):
    arrs=[rng.standard_normal(480000).astype(dt) for _ in range(n)]
    ptrs=(ctypes.c_void_p*n)(*[a.ctypes.data for a in arrs])
    lens=np.full(n,480000,dtype=np.int64)
    out=np.zeros(n,dtype=np.int32)
    for th in (1,4,16):
        t0=time.perf_counter()
        for _ in range(3):
            lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), n, int(dt==np.float64), 480000, ctypes.c_void_p(pin.data_ptr()), 480000, out.ctypes.data_as(ctypes.c_void_p), th)
        dtm=(time.perf_counter()-t0)/3
        print(dt.__name__, 'threads',th, f'{dtm*1e3:.1f} ms')
    t0=time.perf_counter(); x=[a.astype(np.float32) for a in arrs]; print('numpy astype loop', f'{(time.perf_counter()-t0)*1e3:.1f} ms')
print(os.cpu_count(), len(os.sched_getaffinity(0)))

"""Probe: does cudaHostRegister accept NumPy-owned memory here (aligned / unaligned, flags)?"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from varnet_b200 import _capi
lib = _capi.load_library()
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12") if False else None
for n in (1 << 20, 40 << 20):
    a = np.zeros(n, dtype=np.float64)
    t0 = time.perf_counter()
    rc = lib.vn_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
    print("register", n * 8, "addr%%4096=%d" % (a.ctypes.data % 4096), "rc", rc, lib.vn_last_error().decode() if rc else "", "%.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    if rc == 0:
        print("unregister rc", lib.vn_host_unregister(C.c_void_p(a.ctypes.data)))
    b = a[1:]
    rc = lib.vn_host_register(C.c_void_p(b.ctypes.data), b.nbytes)
    print("register view+8", "rc", rc, lib.vn_last_error().decode() if rc else "", flush=True)
    if rc == 0:
        print("unregister rc", lib.vn_host_unregister(C.c_void_p(b.ctypes.data)))

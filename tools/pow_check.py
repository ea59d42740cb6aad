"""Is libdevice pow(w,2)/pow(w,3) exact for dyadic w?  (decides whether an integer bicubic path is usable)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_stream_b200 import vpp
w = np.array([0.0, 0.25, 0.5, 0.75, 0.125, 0.375, 0.625, 0.875] + [k / 64 for k in range(1, 64, 6)] + [k / 4096 for k in (1, 1000, 3001)], dtype=np.float64)
w2, w3 = vpp.device_pow23(w)
bad2 = [(float(a), float(b), float(a * a)) for a, b in zip(w, w2) if b != a * a]
bad3 = [(float(a), float(b), float(a * a * a)) for a, b in zip(w, w3) if b != a * a * a]
print("pow(w,2) inexact:", bad2)
print("pow(w,3) inexact:", bad3)

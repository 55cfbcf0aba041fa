"""Development aid: host time of the native SSML / CSV formatter (pb_ssml_csv) at config-3 scale (46 k rows, 42 MB of CSV)."""
import sys, time, ctypes as C
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import prosody_b200 as pb
from prosody_b200 import ssml as SSML, _native as N
rng=np.random.default_rng(0)
n=46000
words=[" ".join("mot%d" % rng.integers(0,999) for _ in range(rng.integers(2,9))) + rng.choice([",", ".", "", "?"]) for _ in range(n)]
segn=["segment_ph%d" % (i//23) for i in range(n)]
pools=SSML.TextPools(segn, words)
import os; lib=pb._native.load(os.environ['PB_LIB']) if os.environ.get('PB_LIB') else pb._native.load()
pa=rng.integers(0,900,n).astype(np.int32); p=rng.normal(0,3,n); r=rng.normal(0,3,n); v=rng.normal(0,3,n)
i64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64)); dbl = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
for nt in (1,8,16):
  for it in range(3):
    outs = [C.c_void_p() for _ in range(3)]; lens = [C.c_int64() for _ in range(3)]
    t0=time.perf_counter()
    rc = lib.pb_ssml_csv(n, pools.seg, i64(pools.seg_off), pools.txt, i64(pools.txt_off), pa.ctypes.data_as(C.POINTER(C.c_int32)), dbl(p), dbl(r), dbl(v), 1.0, b"fr-FR-HenriNeural", nt, C.byref(outs[0]), C.byref(lens[0]), C.byref(outs[1]), C.byref(lens[1]), C.byref(outs[2]), C.byref(lens[2]))
    t1=time.perf_counter()
    res=tuple(C.string_at(o, l.value) for o, l in zip(outs, lens))
    t2=time.perf_counter()
    for o in outs: lib.pb_ssml_free(o)
    t3=time.perf_counter()
  print(nt, 'native ms', round((t1-t0)*1e3,1), 'string_at ms', round((t2-t1)*1e3,1), 'free', round((t3-t2)*1e3,1))

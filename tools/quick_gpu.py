import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import ctypes as C
from tiktak_b200 import *
from tiktak_b200 import lib
L=lib.load()
fl=C.c_double(); ms=C.c_double()
print('fp64 peak rc', L.tiktak_cuda_fp64_peak(0,C.byref(fl),C.byref(ms)), fl.value/1e12,'TFLOP/s', ms.value,'ms')
for name in ['C2','C4','C3']:
    cfg=baseline_config(name, round_size=256 if name!='C3' else 1024)
    t0=time.time()
    with TikTakSearch(cfg) as s:
        for rep in range(3):
            _,ne,nl=s.solveAtSobolPoints()
            print(name,'pretest',ne,nl,'ms',s.stage_ms('pretest'),'eval ms',s.stage_ms('pretest_eval'), 'evals/s %.3e'%(ne/ (s.stage_ms('pretest_eval')*1e-3)))
        s.chooseSobol(); print(' choose ms',s.stage_ms('choose'))
        if name!='C3':
            t1=time.time(); s.LocalMinimizations('a'); t2=time.time()
            print(' local wall',t2-t1,'s; restarts/s',s.K/(t2-t1),'mean evals',s.evals.mean(),'best',s.getBestLine()[0])
    print(name,'total wall',time.time()-t0)

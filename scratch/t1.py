import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes
from qsvc_b200 import yuv, _lib
from qsvc_b200.mctf import Context
from oracle import oracle as orc
X,Y,bs,sr,a=352,288,16,8,2
clip=yuv.synthetic_clip(X,Y,5,3,max_shift=24)
even,odd=clip[0::2],clip[1::2]
with Context(0) as c:
    c.set_me_mode(2)
    try:
        mv=c.motion_estimate(even,odd,X,Y,bs,sr,a)
        ref=orc.motion_estimate(even,odd,X,Y,bs,sr,a)
        print('diff', (mv!=ref).sum(), 'of', mv.size)
    except Exception as e:
        print('ERR', e)
    L=ctypes.CDLL(_lib.SO_PATH); print('tma timeouts', L.qsvc_debug_tma_timeouts())

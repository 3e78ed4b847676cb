import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from qsvc_b200 import yuv
from qsvc_b200.mctf import Context
X,Y,GOPs,TRLs,bs,sr,a=1920,1080,1,4,16,16,2
clip=yuv.synthetic_clip(X,Y,GOPs*2**(TRLs-1)+1,2,max_shift=48)
with Context(0) as c:
    for uf in (0.0, 0.25):
        c.analyze(clip,X,Y,GOPs,TRLs,bs,sr,a,uf,always_B=1,block_size_min=bs)
        t0=time.time(); c.analyze(clip,X,Y,GOPs,TRLs,bs,sr,a,uf,always_B=1,block_size_min=bs); print('uf',uf,'9-frame 1080p analyze', round((time.time()-t0)*1e3,1),'ms')

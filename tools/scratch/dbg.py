import sys; sys.path.insert(0,'.')
import numpy as np, nnsp_b200 as nb
from oracle.pyoracle import Oracle
O=Oracle()
S,T=8,4
pcm=nb.synth_pcm(S,T)
m=nb.Model.from_blob(nb.MODEL_DIR+'/vad.nnspm')
b=nb.NNSPBatch(m,S)
res,taps=b.exec(pcm,taps=True)
mo=O.model(1,False)
for s in (0,5):
    r,tp=O.nnsp_run(mo,pcm[s])
    print('stream',s)
    print('gpu  L0', taps['act'][s,0,:28])
    print('orac L0', tp.act[0,:28])
    print('feat eq', (taps['feat'][s]==tp.feat).all())
    print('gpu  L1', taps['act'][s,0,28:56])
    print('orac L1', tp.act[0,28:56])

import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import make_labels
from test_ctc_gpu import _gpu_loss
from oracle import ctc_oracle as oracle
for (T,B,C,maxlen,scale) in [(250,5,20,120,2.0),(250,5,20,120,1.0),(250,5,20,100,2.0),(250,5,20,70,2.0),(250,5,20,20,2.0),(120,5,20,50,2.0),(250,8,20,40,2.0),(509,4,30,30,1.0)]:
    rng = np.random.default_rng(T * 1000 + B)
    x = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
    seq_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    seq_len[0] = T
    labels = make_labels(rng, B, seq_len, max_len=maxlen, num_labels=C - 1, repeat_p=0.3)
    loss, grad, st = _gpu_loss(x, labels, seq_len)
    l64, g64, s64 = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    err = np.abs(grad-g64)
    print(T,B,C,maxlen,scale,'Ls',[len(l) for l in labels],'Tb',seq_len.tolist())
    for b in range(B):
        t,k = np.unravel_index(err[:,b].argmax(), err[:,b].shape)
        print('  b',b,'maxerr %.3g at t=%d k=%d'%(err[:,b].max(),t,k),'loss',loss[b],l64[b], 'rowsum@t %.3g'%grad[t,b].sum())

import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import make_labels
from test_ctc_gpu import _gpu_loss
from oracle import ctc_oracle as oracle
T,B,C,maxlen,scale = 250,5,20,20,2.0
rng = np.random.default_rng(T * 1000 + B)
x = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
seq_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
seq_len[0] = T
labels = make_labels(rng, B, seq_len, max_len=maxlen, num_labels=C - 1, repeat_p=0.3)
for path in (0,2):
    loss, grad, st = _gpu_loss(x, labels, seq_len, path=path)
    l64, g64, s64 = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    print('path',path)
    for b in range(B):
        bad = np.argwhere(~np.isfinite(grad[:,b]))
        print(' b',b,'Tb',seq_len[b],'L',len(labels[b]),'lab',labels[b],'n_nonfinite',len(bad), bad[:40].tolist())
        if len(bad):
            t=bad[0][0]
            print('   row',t,grad[t,b], 'ref', g64[t,b])
            print('   x row', x[t,b])

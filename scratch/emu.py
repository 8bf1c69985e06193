import sys, numpy as np
sys.path.insert(0,'tests')
from util import make_labels
T,B,C,maxlen,scale = 250,5,20,20,2.0
rng = np.random.default_rng(T * 1000 + B)
x = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
seq_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
seq_len[0] = T
labels = make_labels(rng, B, seq_len, max_len=maxlen, num_labels=C - 1, repeat_p=0.3)
b=3; Tb=int(seq_len[b]); lab=labels[b]; L=len(lab); blank=C-1
xb=x[:Tb,b]; e=np.exp(xb-xb.max(1,keepdims=True)).astype(np.float32)
f=np.float32
def rescale(m):
    bits=np.float32(m).view(np.uint32); ee=int(bits>>23)
    if ee==0 or ee==255: return f(1),0
    return np.uint32((254-ee)<<23).view(np.float32), ee-127
# A chain
ab=np.zeros(L+1,f); al=np.zeros(L+1,f); sc=f(1)
A_b=np.zeros((Tb,L+1),f); A_l=np.zeros((Tb,L+1),f)
for t in range(Tb):
    nb=np.zeros(L+1,f); nl=np.zeros(L+1,f)
    for i in range(L+1):
        if t==0:
            nb[i]=e[t,blank] if i==0 else 0; nl[i]=(e[t,lab[0]] if (i==0 and L>0) else 0)
        else:
            pl=al[i-1] if i>=1 else f(0)
            nb[i]=e[t,blank]*(ab[i]+pl)
            sk=1 if (i>=1 and i<L and lab[i-1]!=lab[i]) else 0
            nl[i]=(e[t,lab[i]]*(al[i]+ab[i]+sk*pl)) if i<L else 0
    for i in range(L+1):
        tdl = Tb-(L-i) if i<L else Tb
        ab[i]=nb[i]*sc if t<tdl else 0
        al[i]=nl[i]*sc if (t<=tdl and i<L) else 0
    m=max(ab.max(),al.max()); sc,ex=rescale(m)
    A_b[t]=ab; A_l[t]=al
print('A max per row range', A_b.max(), A_l.max(), 'rows 100-110', [float(max(A_b[t].max(),A_l[t].max())) for t in range(100,110)])
# B chain: canonical arrays bb[i]=beta(blank i), bl[i]=beta(label i)
bb=np.zeros(L+1,f); bl=np.zeros(L+1,f); sc=f(1)
B_b=np.zeros((Tb,L+1),f); B_l=np.zeros((Tb,L+1),f)
for t in range(Tb-1,-1,-1):
    nbb=np.zeros(L+1,f); nbl=np.zeros(L+1,f)
    if t==Tb-1:
        nbb[L]=1
        if L>=1: nbl[L-1]=1
    else:
        wb=bb*e[t+1,blank]; wl=np.array([bl[i]*e[t+1,lab[i]] if i<L else 0 for i in range(L+1)],f)
        for i in range(L+1):
            nbb[i]=wb[i]+(wl[i] if i<L else 0)
            if i<L:
                sk=1 if (i+1<L and lab[i+1]!=lab[i]) else 0
                nbl[i]=wl[i]+wb[i+1]+sk*(wl[i+1] if i+1<L else 0)
    for i in range(L+1):
        bb[i]=nbb[i]*sc if t>=i else 0
        bl[i]=nbl[i]*sc if (t>=i and i<L) else 0
    m=max(bb.max(),bl.max()); sc,ex=rescale(m)
    B_b[t]=bb; B_l[t]=bl
print('B max', B_b.max(), B_l.max(), [float(max(B_b[t].max(),B_l[t].max())) for t in range(100,110)])
P=A_b*B_b; Pl=A_l*B_l
print('prod max', P.max(), Pl.max(), 'rows', [float(P[t].sum()+Pl[t].sum()) for t in range(100,110)])

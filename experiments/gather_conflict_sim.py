"""Probe (NOT product code, CPU only): bank-conflict model of the SRP gather over the real C2 lag table in the k-d slot
order of asw_srp_create (restated here): wavefronts per LDS.32 for the product layout, skewed layouts and a
lane-per-tap mapping.  See profiles/r02_notes.md section 10.1."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth, native
sys.setrecursionlimit(10000)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
from acousticswarms_speech_b200.srp_phat import SRP_PHAT
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05, threshold=list(SRP_THRESHOLDS), WIDTH=8, build_native=False)
lag = native.pair_lags(node.grids, scene.mic_positions, 48000, 343.0)
G,P = lag.shape
print(G,P)
def kd(idx):
    n=len(idx)
    if n<=32: return idx
    sub=lag[idx]
    bp=np.argmax(sub.max(0)-sub.min(0))
    mid=((n//2+31)//32)*32
    if mid>=n: mid = ((n-1)//32)*32 if n-32>0 else n//2
    order=np.lexsort((idx, sub[:,bp]))
    idx=idx[order]
    return np.concatenate([kd(idx[:mid]), kd(idx[mid:])])
perm=kd(np.arange(G))
U=4
lo=np.floor(lag.min(0)).astype(int)-2
pos=(lag-lo)*U          # entries
i0=np.floor(pos).astype(int)[perm]   # [G][P] in slot order
def wavefronts(addr):   # addr: [32] word addresses -> wavefronts for one LDS.32
    banks={}
    for a in addr: banks.setdefault(a%32,set()).add(a)
    return max(len(v) for v in banks.values())
def total(fn):
    tot=0; ideal=0
    for w in range(0,G-31,32):
        blk=i0[w:w+32]
        for p in range(P):
            a=fn(blk[:,p])
            tot+=wavefronts(a); ideal+=1
    return tot/ideal
print('baseline', total(lambda a:a))
print('skew+1 per 32', total(lambda a:a+(a>>5)))
print('skew+1 per 16', total(lambda a:a+(a>>4)))
print('skew+3 per 32', total(lambda a:a+3*(a>>5)))
# span stats
sp=[]
for w in range(0,G-31,32):
    blk=i0[w:w+32]; sp.append(blk.max(0)-blk.min(0))
sp=np.array(sp); print('span mean per pair', sp.mean(0).round(0)); print('span overall mean', sp.mean(), 'max', sp.max())
def total8():
    tot=0; cnt=0
    for w in range(0,G-7,8):
        blk=i0[w:w+8]
        for p in range(P):
            a=(blk[:,p][:,None]-1+np.arange(4)[None,:]).ravel()
            tot+=wavefronts(a); cnt+=1
    return tot/cnt
print('8 hyp x 4 taps per LDS: wavefronts', total8(), ' -> per hypercube-tap-set', total8()/8, 'vs baseline', 4*1.6577/32)


def total16():
    tot = 0; cnt = 0
    for w in range(0, G - 15, 16):
        blk = i0[w:w + 16]
        for p in range(P):
            for first in (0, 1):                      # lane (h, j): taps i0 - 1 + 2 j + first
                a = (blk[:, p][:, None] - 1 + 2 * np.arange(2)[None, :] + first).ravel()
                tot += wavefronts(a); cnt += 1
    return tot / cnt


t16 = total16()
print('16 hyp x 2 lanes, 2 loads each: wavefronts per load', t16, ' -> per hypercube', 2 * t16 / 16, 'vs baseline', 4 * 1.6577 / 32)

"""Developer probe (CPU, ~7 min on 8 cores): the oracle's posterior sums over all 44 850 pairs of the bundled ASMC example
against the reference's golden sumOverPairs (golden G4).  Measured: L1 relative difference 7.8e-6, largest elementwise
relative difference 1.2e-4 (the golden is printed at 6 significant digits).  Usage: python tests/probes/g4_oracle_check.py"""
import sys, gzip, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import ASMC_EXAMPLE, DQ_69
from oracle import pyoracle
pyoracle.build()
o = pyoracle.Oracle(ASMC_EXAMPLE, DQ_69, "/tmp/x", hashing=False, FastSMC=False, asmcMode=True, batchSize=64, useKnownSeed=True)
gold=np.loadtxt(gzip.open(os.path.join(ROOT, 'tests', 'golden', 'asmc_sum_over_pairs.gz'),'rt'))
n=o.num_haps//2
a=[];b=[]
for i in range(n):
    for j in range(i):
        for ih in (0,1):
            for jh in (0,1):
                a.append(2*j+jh); b.append(2*i+ih)
    a.append(2*i); b.append(2*i+1)
a=np.array(a,np.uint32); b=np.array(b,np.uint32)
print(len(a), flush=True)
N=int(sys.argv[1]) if len(sys.argv)>1 else len(a)
tot=np.zeros((o.sites,o.states))
t=time.time()
for s in range(0,N,256):
    p=o.decode_posterior(a[s:s+256],b[s:s+256])
    tot+=p.astype(np.float64).sum(axis=0)
    if s%2560==0: print(s, time.time()-t, flush=True)
np.save('/tmp/g4_oracle_sum.npy',tot)
if N==len(a):
    rel=np.abs(tot-gold)/np.maximum(gold,1e-3)
    print("max rel",rel.max(),"mean rel",rel.mean(),"L1 rel",np.abs(tot-gold).sum()/gold.sum())

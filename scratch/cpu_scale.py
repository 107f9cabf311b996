import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
from oracle import cref_binding as C
from common import *
from pyref import poseidon as P
ctx=C.Ctx(resource(20,'rln_final.arkzkey'),resource(20,'graph.bin'))
fs=fr_stream(5); pe=[P.poseidon([i+7]) for i in range(20)]
import os
n=256; inp=[];rs=[]
for j in range(n):
    inp.append(ctx.inputs_buffer(next(fs),100,j%100,pe,[(j>>i)&1 for i in range(20)],next(fs),12345)); rs+=[next(fs),next(fs)]
inp=b''.join(inp); rs=fr_bytes(rs)
for th,k in ((1,2),(8,16),(32,64),(64,128),(128,256)):
    t=time.perf_counter(); ctx.prove_batch(inp[:k*ctx.inputs_size*32],rs[:64*k],k,th); dt=time.perf_counter()-t
    print(th,'threads',k,'proofs',round(dt,3),'s ->',round(k/dt,2),'proofs/s', round(dt*th/k,3),'thread-s/proof')

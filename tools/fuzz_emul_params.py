#!/usr/bin/env python3
"""Parameter fuzz of the DEVICE code without a GPU: random star_params_t settings (seeding, windows, stitching, scores, filters) on a few
reads; the kernel sources run as emulated CTAs (oracle/engine_emul.cpp) and must give the oracle's alignments field by field, on the
default flat pipeline or on the lane path.  usage: python tools/fuzz_emul_params.py SEED [NREADS]   (reads from /tmp/tp/tiny)"""
import os, random, sys, ctypes as C
import numpy as np
ROOT="/root/repo"; sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+"/tests")
import conftest as cf, oracle_capi as oc, star_b200 as sb
from star_b200 import capi
seed=int(sys.argv[1]); nreads=int(sys.argv[2]) if len(sys.argv)>2 else 8
random.seed(seed)
g="/tmp/tp/tiny"
lib=sb.load_library(); ol=oc.load_oracle()
p=capi.default_params(lib)
def ch(*a): return random.choice(a)
desc=[]
def setp(prob,name,*vals):
    if random.random()<prob:
        v=ch(*vals); setattr(p,name,v); desc.append("%s=%s"%(name,v))
setp(.3,"seedSearchStartLmax",20,30,50,70); setp(.2,"seedSearchStartLmaxOverLread",0.5,1.0,0.3); setp(.2,"seedMultimapNmax",100,1000,10000)
setp(.2,"seedPerWindowNmax",10,30,50); setp(.2,"seedSplitMin",8,12,20); setp(.2,"seedMapMin",3,5,10); setp(.3,"winAnchorMultimapNmax",20,50,200)
setp(.2,"alignIntronMax",0,5000,100000,1000000); setp(.2,"alignMatesGapMax",0,1000,100000); setp(.2,"alignIntronMin",10,21,50)
setp(.2,"alignSJoverhangMin",3,5,8); setp(.2,"alignSJDBoverhangMin",1,3,5); setp(.2,"alignSplicedMateMapLmin",0,20); setp(.2,"alignSplicedMateMapLminOverLmate",0.3,0.66,0.9)
setp(.15,"alignTranscriptsPerWindowNmax",10,100); setp(.15,"alignSoftClipAtReferenceEnds",0,1); setp(.15,"alignInsertionFlushRight",0,1)
if random.random()<.3:
    t=ch("EndToEnd","Extend5pOfRead1","Extend5pOfReads12","Extend3pOfRead1"); desc.append("alignEndsType="+t)
    e={"EndToEnd":((1,1),(1,1)),"Extend5pOfRead1":((1,0),(0,0)),"Extend5pOfReads12":((1,0),(1,0)),"Extend3pOfRead1":((0,1),(0,0))}[t]
    for i in range(2):
        for j in range(2): p.alignEndsTypeExt[i][j]=e[i][j]
if random.random()<.15:
    p.alignEndsProtrudeNbasesMax=ch(5,10); p.alignEndsProtrudeConcordantPair=ch(0,1); desc.append("protrude=%d,%d"%(p.alignEndsProtrudeNbasesMax,p.alignEndsProtrudeConcordantPair))
setp(.2,"scoreGap",0,-2,-4); setp(.2,"scoreGapNoncan",-8,-4,-12); setp(.15,"scoreGapGCAG",-4,-2); setp(.15,"scoreGapATAC",-8,-4); setp(.2,"scoreGenomicLengthLog2scale",0,0.25,0.5,1)
setp(.2,"scoreDelOpen",-2,-1,-4); setp(.2,"scoreDelBase",-2,-1); setp(.2,"scoreInsOpen",-2,-1,-4); setp(.2,"scoreInsBase",-2,-1); setp(.2,"scoreStitchSJshift",0,1,2); setp(.2,"sjdbScore",0,1,2,4)
setp(.3,"outFilterMismatchNmax",3,10,999); setp(.3,"outFilterMismatchNoverLmax",0.05,0.1,0.3,1); setp(.2,"outFilterMismatchNoverReadLmax",0.04,0.1,1)
setp(.3,"outFilterMultimapNmax",1,3,10,20); setp(.2,"outFilterMultimapScoreRange",0,1,3); setp(.2,"outFilterScoreMin",0,50); setp(.2,"outFilterScoreMinOverLread",0.3,0.66,0.9)
setp(.2,"outFilterMatchNmin",0,50); setp(.2,"outFilterMatchNminOverLread",0.3,0.66,0.9); setp(.2,"outFilterIntronMotifs",0,1,2); setp(.2,"outFilterIntronStrandsRemoveInconsistent",0,1)
setp(.2,"outSAMstrandFieldType",0,1); setp(.2,"outSAMprimaryFlagAllBestScore",0,1); setp(.2,"outSAMmultNmax",2**64-1,1,2)
ds=ch(("std",2),("hard",2),("se",1),("hard",1))
idxdir=ch(g+"/idx",g+"/idx0")
flat=random.random()<0.6
env={} if flat else {"STAR_B200_HEAVY_NA":"2000000000","STAR_B200_HEAVY_EST":"0"}
for k,v in env.items(): os.environ[k]=v
n0=nreads if ds[0]!="hard" else max(1,nreads//4)
if not flat: n0*=2
start=random.randrange(0,300)
mates=[cf.read_fastq_seqs("%s/%s_%d.fq"%(g,ds[0],m+1))[start:start+n0] for m in range(ds[1])]
seq,off,n,nm=sb.pack_reads(mates)
idx=sb.Index(lib,idxdir,params=p)
oe=oc.OracleEngine(ol,idx)
res_o,al_o,_=oe.map_chunk(seq,off,n,nm)
batch=oe._batch(seq,off,n,nm); res,al,ab=oe._out(n,oe.n_out); oe.close()
em=C.CDLL(ROOT+"/oracle/_build/libengine_emul.so")
em.engine_emul_map_chunk.argtypes=[C.POINTER(capi.IndexView),C.POINTER(capi.Params),C.POINTER(capi.ReadBatch),C.POINTER(capi.AlignBatch),C.c_void_p]
info4=np.zeros(4,dtype=np.uint64)
rc=em.engine_emul_map_chunk(idx.view,C.byref(idx.params),C.byref(batch),C.byref(ab),info4.ctypes.data)
diffs=oc.compare_outputs(res_o,al_o,res,al[:ab.nAligns]) if rc==0 and int(info4[2])==0 else ["rc %d overflow %d"%(rc,int(info4[2]))]
print("seed",seed,"OK" if not diffs else "MISMATCH",ds,os.path.basename(idxdir),"flat" if flat else "lane","n",n,"info",info4.tolist()," ".join(desc), diffs[:3], flush=True)

#!/usr/bin/env python3
"""Differential fuzz of --runMode genomeGenerate against the LIVE reference (build container only: needs oracle/_ref/STAR).
usage: python tools/fuzz_genome_generate.py SEED      (works in /tmp/tp/fg<SEED>)
Random genomes (repeats, reverse-complement copies, N runs, low complexity, IUPAC / lowercase, identical chromosomes, bin-sized
chromosomes, one or two FASTA files, random --genomeChrBinNbits / --genomeSAindexNbases) are indexed by the reference, by the
oracle-driven CLI (comparison sort) and by the emulated kernels (prefix doubling); Genome, SA, SAindex and the chr files must be
identical.  Round 1: seeds 1-60 identical."""
import random, subprocess, os, sys, shutil, filecmp
seed=int(sys.argv[1]); random.seed(seed)
d="/tmp/tp/fg%d"%seed; shutil.rmtree(d,ignore_errors=True); os.makedirs(d); os.chdir(d)
def rnd(n,alpha="ACGT"): return "".join(random.choice(alpha) for _ in range(n))
comp={"A":"T","C":"G","G":"C","T":"A"}
def rc(s): return "".join(comp.get(c,"N") for c in reversed(s))
pool=[rnd(random.randint(50,4000)) for _ in range(4)]
chrs=[]
for c in range(random.randint(1,6)):
    s=""
    for k in range(random.randint(1,8)):
        r=random.random()
        if r<0.35: s+=rnd(random.randint(1,3000))
        elif r<0.55: s+=random.choice(pool)
        elif r<0.7: s+=rc(random.choice(pool))
        elif r<0.8: s+="N"*random.randint(1,300)
        elif r<0.9: s+=random.choice(["A","AC","ACG","T"])*random.randint(5,400)
        else: s+=rnd(random.randint(1,200),"ACGTNacgtnRY")
    if not s: s="A"
    chrs.append(s)
if random.random()<0.3: chrs.append(chrs[0])
bits=random.choice([8,9,10,12])
if random.random()<0.3: chrs.append(rnd((1<<bits)*random.randint(1,2)))
nb=random.choice([3,4,5,6])
files=["a.fa"] if random.random()<0.6 else ["a.fa","b.fa"]
per=[chrs] if len(files)==1 else [chrs[:len(chrs)//2] or chrs[:1], chrs[len(chrs)//2:] or chrs[:1]]
ci=0
for fn,cs in zip(files,per):
    with open(fn,"w") as f:
        for s in cs:
            ci+=1; f.write(">c%d desc\n"%ci); w=random.choice([50,60,80,1000])
            for i in range(0,len(s),w): f.write(s[i:i+w]+"\n")
args=["--runMode","genomeGenerate","--genomeFastaFiles"]+files+["--genomeSAindexNbases",str(nb),"--genomeChrBinNbits",str(bits)]
env=dict(os.environ)
for tag,b,e in (("ref","/root/repo/oracle/_ref/STAR",None),("our","/root/repo/oracle/_build/star_cli_oracle",None),("emu","/root/repo/oracle/_build/star_cli_oracle",dict(env,STAR_CLI_SJDB_EMUL="/root/repo/oracle/_build/libengine_emul.so"))):
    os.makedirs(tag)
    rc=subprocess.call([b]+args+["--genomeDir",tag,"--outFileNamePrefix",tag+"_","--runThreadN","2"],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL,env=e)
    if rc: print("seed",seed,tag,"rc",rc)
res={}
for tag in ("our","emu"):
    res[tag]=[f for f in ("Genome","SA","SAindex","chrStart.txt","chrLength.txt","chrName.txt") if not (os.path.exists("ref/"+f) and os.path.exists(tag+"/"+f) and filecmp.cmp("ref/"+f,tag+"/"+f,shallow=False))]
print("seed",seed,"nchr",len(chrs),"bits",bits,"nb",nb,"bad",res)
os.chdir("/tmp"); 
if not res["our"] and not res["emu"]: shutil.rmtree(d)

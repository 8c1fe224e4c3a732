#!/usr/bin/env python3
"""Differential fuzz of the on-the-fly junction insertion against the LIVE reference (build container only: needs oracle/_ref/STAR).
usage: cd <unpacked tiny/ with idx/ and idx0/>; python tools/fuzz_sjdb_insert.py SEED
Random junction lists (known junctions moved / re-stranded / duplicated, random introns, missing strand column) are inserted with
--sjdbInsertSave All by the reference and by the oracle-driven CLI; sjdbInfo.txt, sjdbList.out.tab, SA, SAindex, Genome and the SAM
must be identical.  Round 1: seeds 1-70 identical (after the missing-strand-column behaviour was mirrored)."""
import random, subprocess, os, sys, shutil, filecmp
seed=int(sys.argv[1]); random.seed(seed)
rows=[l.split("\t") for l in open("idx/sjdbList.out.tab").read().splitlines()]
chrs={}
for l in open("idx/chrNameLength.txt"): n,L=l.split(); chrs[n]=int(L)
out=[]
for k in range(random.randint(5,120)):
    r=random.random()
    if r<0.5:
        c,s,e,st=random.choice(rows); s=int(s); e=int(e)
        if random.random()<0.4: d=random.randint(-4,4); s+=d; e+=d
        if random.random()<0.3: e+=random.randint(-3,3)
        st=random.choice("+-.12 0x")
    else:
        c=random.choice(list(chrs)); s=random.randint(300,chrs[c]-3000); e=s+random.randint(21,2000); st=random.choice("+-.")
    if e<=s: continue
    out.append("%s\t%d\t%d\t%s\n"%(c,s,e,st))
    if random.random()<0.2: out.append(out[-1])
open("fz.tab","w").writelines(out)
idx=random.choice(["idx","idx0"])
ov=random.choice([99,99,30,5]) if idx=="idx0" else 99
args=["--genomeDir",idx,"--readFilesIn","se_1.fq","--readMapNumber","50","--sjdbFileChrStartEnd","fz.tab","--sjdbInsertSave","All","--sjdbOverhang",str(ov),"--runThreadN","1"]
for tag,b in (("fz_ref","/root/repo/oracle/_ref/STAR"),("fz_our","/root/repo/oracle/_build/star_cli_oracle")):
    shutil.rmtree(tag,ignore_errors=True); os.makedirs(tag)
    rc=subprocess.call([b]+args+["--outFileNamePrefix",tag+"/"],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
    if rc: print("seed",seed,tag,"rc",rc)
bad=[f for f in ("sjdbInfo.txt","sjdbList.out.tab","SA","SAindex","Genome") if not (os.path.exists("fz_ref/_STARgenome/"+f) and os.path.exists("fz_our/_STARgenome/"+f) and filecmp.cmp("fz_ref/_STARgenome/"+f,"fz_our/_STARgenome/"+f,shallow=False))]
sam= open("fz_ref/Aligned.out.sam").read().split("@CO")[-1].split("\n",1)[1]==open("fz_our/Aligned.out.sam").read().split("@CO")[-1].split("\n",1)[1] if os.path.exists("fz_ref/Aligned.out.sam") and os.path.exists("fz_our/Aligned.out.sam") else False
print("seed",seed,idx,ov,len(out),"bad",bad,"sam",sam)

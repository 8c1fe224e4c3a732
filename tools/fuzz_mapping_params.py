#!/usr/bin/env python3
"""Differential fuzz of the mapping path against the LIVE reference (build container only: needs oracle/_ref/STAR).
usage: python tools/fuzz_mapping_params.py SEED   (works in /tmp/tp/tiny = unpacked tiny.tar.gz + idx0 of twopass.tar.gz)
Random combinations of ~55 seeding / window / stitching / scoring / filter / output parameters (incl. BySJout, 2-pass, attribute sets)
on the std / hard / se read sets, annotated or plain index: exit code, SAM body, SJ.out.tab and Log.final.out counters of the
oracle-driven CLI must equal the reference's.  Round 1: seeds 1-60 identical (after mirroring "ch requires BAM output")."""
import random, subprocess, os, sys, shutil
seed=int(sys.argv[1]); random.seed(seed)
os.chdir("/tmp/tp/tiny")
def ch(*a): return random.choice(a)
opts=[]
def maybe(p, name, *vals):
    if random.random()<p: opts.extend([name]+[str(v) for v in ch(*vals)] if isinstance(vals[0],(list,tuple)) else [name,str(ch(*vals))])
maybe(.3,"--seedSearchStartLmax",20,30,50,70)
maybe(.2,"--seedSearchStartLmaxOverLread",0.5,1.0,0.3)
maybe(.2,"--seedMultimapNmax",100,1000,10000)
maybe(.2,"--seedPerReadNmax",200,1000)
maybe(.2,"--seedPerWindowNmax",10,30,50)
maybe(.2,"--seedSplitMin",8,12,20)
maybe(.2,"--seedMapMin",3,5,10)
maybe(.3,"--winAnchorMultimapNmax",20,50,200)
maybe(.2,"--winBinNbits",12,14,16)
maybe(.2,"--alignIntronMax",0,5000,100000,1000000)
maybe(.2,"--alignMatesGapMax",0,1000,100000)
maybe(.2,"--alignIntronMin",10,21,50)
maybe(.2,"--alignSJoverhangMin",3,5,8)
maybe(.2,"--alignSJDBoverhangMin",1,3,5)
maybe(.2,"--alignSplicedMateMapLmin",0,20)
maybe(.2,"--alignSplicedMateMapLminOverLmate",0.3,0.66,0.9)
maybe(.15,"--alignWindowsPerReadNmax",100,10000)
maybe(.15,"--alignTranscriptsPerWindowNmax",10,100)
maybe(.15,"--alignTranscriptsPerReadNmax",1000,10000)
maybe(.3,"--alignEndsType","Local","EndToEnd","Extend5pOfRead1","Extend5pOfReads12")
maybe(.15,"--alignSoftClipAtReferenceEnds","Yes","No")
maybe(.15,"--alignInsertionFlush","None","Right")
maybe(.15,"--alignEndsProtrude",["5","ConcordantPair"],["10","DiscordantPair"],["0","ConcordantPair"])
maybe(.2,"--scoreGap",0,-2,-4)
maybe(.2,"--scoreGapNoncan",-8,-4,-12)
maybe(.15,"--scoreGapGCAG",-4,-2)
maybe(.15,"--scoreGapATAC",-8,-4)
maybe(.2,"--scoreGenomicLengthLog2scale",0,0.25,0.5,1)
maybe(.2,"--scoreDelOpen",-2,-1,-4)
maybe(.2,"--scoreDelBase",-2,-1)
maybe(.2,"--scoreInsOpen",-2,-1,-4)
maybe(.2,"--scoreInsBase",-2,-1)
maybe(.2,"--scoreStitchSJshift",0,1,2)
maybe(.2,"--sjdbScore",0,1,2,4)
maybe(.3,"--outFilterMismatchNmax",3,10,999)
maybe(.3,"--outFilterMismatchNoverLmax",0.05,0.1,0.3,1)
maybe(.2,"--outFilterMismatchNoverReadLmax",0.04,0.1,1)
maybe(.3,"--outFilterMultimapNmax",1,3,10,20)
maybe(.2,"--outFilterMultimapScoreRange",0,1,3)
maybe(.2,"--outFilterScoreMin",0,50)
maybe(.2,"--outFilterScoreMinOverLread",0.3,0.66,0.9)
maybe(.2,"--outFilterMatchNmin",0,50)
maybe(.2,"--outFilterMatchNminOverLread",0.3,0.66,0.9)
maybe(.2,"--outFilterIntronMotifs","None","RemoveNoncanonical","RemoveNoncanonicalUnannotated")
maybe(.2,"--outFilterIntronStrands","RemoveInconsistentStrands","None")
maybe(.2,"--outSAMstrandField","None","intronMotif")
maybe(.2,"--outSAMprimaryFlag","OneBestScore","AllBestScore")
maybe(.2,"--outSAMmultNmax",-1,1,2)
maybe(.2,"--outFilterType","Normal","BySJout")
maybe(.3,"--outSAMunmapped",["Within"],["Within","KeepPairs"],["None"])
maybe(.3,"--outSAMattributes",["NH","HI","AS","nM","NM","MD","jM","jI","MC"],["All"],["Standard"],["NH","HI","XS"])
maybe(.15,"--twopassMode","Basic")
maybe(.2,"--outSJfilterReads","All","Unique")
maybe(.25,"--outReadsUnmapped","Fastx")
qm=random.random()<0.35
ds=ch(("std",["std_1.fq","std_2.fq"]),("hard",["hard_1.fq","hard_2.fq"]),("se",["se_1.fq"]),("hard1",["hard_1.fq"]))
idx=ch("idx","idx0")
if qm:
    idx="idx"; opts.extend(["--quantMode"]+ch(["GeneCounts"],["TranscriptomeSAM"],["TranscriptomeSAM","GeneCounts"]))
    if random.random()<0.4: opts.extend(["--quantTranscriptomeSAMoutput",ch("BanSingleEnd","BanSingleEnd_ExtendSoftclip","BanSingleEnd_BanIndels_ExtendSoftclip")])
if random.random()<0.5:
    nm=len(ds[1])
    if random.random()<0.6: opts.extend(["--clip5pNbases"]+[str(ch(0,3,10,40)) for _ in range(nm)])
    if random.random()<0.6: opts.extend(["--clip3pNbases"]+[str(ch(0,5,20,90,200)) for _ in range(nm)])
    if random.random()<0.5:
        opts.extend(["--clip3pAdapterSeq"]+[ch("AGGTC","GATC","polyA","TTTTTTTT","-") for _ in range(nm)])
        opts.extend(["--clip3pAdapterMMp"]+[str(ch(0,0.1,0.3)) for _ in range(nm)])
        if random.random()<0.4: opts.extend(["--clip3pAfterAdapterNbases"]+[str(ch(0,1,4)) for _ in range(nm)])
args=["--genomeDir",idx,"--readFilesIn"]+ds[1]+opts
res=[]
for tag,b,nt in (("fm_ref","/root/repo/oracle/_ref/STAR","1"),("fm_our","/root/repo/oracle/_build/star_cli_oracle","3")):
    d="%s%d"%(tag,seed); shutil.rmtree(d,ignore_errors=True); os.makedirs(d)
    p=subprocess.run([b]+args+["--outFileNamePrefix",d+"/","--runThreadN",nt],stdout=subprocess.DEVNULL,stderr=subprocess.PIPE,text=True)
    res.append((p.returncode,p.stderr[:200]))
def body(fn):
    try: return [l for l in open(fn) if not l.startswith("@")]
    except Exception: return None
def logc(fn):
    try: return [l for l in open(fn) if "|" in l and not any(k in l for k in ("Started","Finished","speed"))]
    except Exception: return None
r="fm_ref%d/"%seed; o="fm_our%d/"%seed
import gzip
def extra():
    for f in ("ReadsPerGene.out.tab","Unmapped.out.mate1","Unmapped.out.mate2"):
        if os.path.exists(r+f) != os.path.exists(o+f): return False
        if os.path.exists(r+f) and open(r+f,"rb").read()!=open(o+f,"rb").read(): return False
    f="Aligned.toTranscriptome.out.bam"
    if os.path.exists(r+f) != os.path.exists(o+f): return False
    if os.path.exists(r+f) and gzip.decompress(open(r+f,"rb").read())!=gzip.decompress(open(o+f,"rb").read()): return False
    return True
ok = res[0][0]==res[1][0] and (res[0][0]!=0 or (extra() and body(r+"Aligned.out.sam")==body(o+"Aligned.out.sam") and open(r+"SJ.out.tab").read()==open(o+"SJ.out.tab").read() and logc(r+"Log.final.out")==logc(o+"Log.final.out")))
print("seed",seed,"OK" if ok else "MISMATCH",res[0][0],res[1][0],ds[0],idx," ".join(opts), "" if ok else res[1][1], flush=True)
if ok: shutil.rmtree(r,ignore_errors=True); shutil.rmtree(o,ignore_errors=True)

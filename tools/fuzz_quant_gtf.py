#!/usr/bin/env python3
"""Differential fuzz of --quantMode TranscriptomeSAM GeneCounts with random GTFs (kept / altered / random transcripts, overlapping genes, strands) given at the
mapping stage, against the LIVE reference: transcriptome BAM, gene counts, annotation tables, junction database, SAM.  Round 1: seeds 1-100 identical."""
import random, subprocess, os, sys, shutil, gzip
seed=int(sys.argv[1]); random.seed(seed)
os.chdir("/tmp/tp/tiny")
chrs={}
for l in open("idx0/chrNameLength.txt"): n,L=l.split(); chrs[n]=int(L)
orig=[l.rstrip("\n").split("\t") for l in open("annot.gtf") if "\texon\t" in l]
lines=[]
# keep a random subset of original transcripts (reads come from them), sometimes with strand flipped or '.'
keep=set(random.sample(sorted(set(f[8].split('transcript_id "')[1].split('"')[0] for f in orig)), k=random.randint(3,10)))
for f in orig:
    tid=f[8].split('transcript_id "')[1].split('"')[0]
    if tid in keep:
        g=list(f)
        if random.random()<0.15: g[6]=random.choice("+-.")
        lines.append("\t".join(g))
# alternative transcripts: drop an exon / shift a boundary
for tid in list(keep)[:random.randint(0,5)]:
    ex=[f for f in orig if 'transcript_id "%s"'%tid in f[8]]
    if len(ex)<3: continue
    k=random.randrange(len(ex))
    for i,f in enumerate(ex):
        if i==k and random.random()<0.7: continue
        g=list(f); g[8]=g[8].replace(tid,tid+"_alt")
        if random.random()<0.3: g[3]=str(max(1,int(g[3])+random.randint(-20,20)))
        if random.random()<0.3: g[4]=str(int(g[4])+random.randint(-20,20))
        if int(g[4])<=int(g[3]): continue
        lines.append("\t".join(g))
# random genes, some overlapping existing ones, some sharing a gene id (nested / overlapping genes)
for gi in range(random.randint(2,15)):
    c=random.choice(list(chrs)); st=random.choice("+-")
    pos=random.randint(1,chrs[c]-6000)
    gid="R%03d"%(gi if random.random()<0.8 else max(0,gi-1))
    for ti in range(random.randint(1,3)):
        p=pos+random.randint(0,300); 
        for e in range(random.randint(1,6)):
            L=random.randint(20,400); 
            if p+L>=chrs[c]: break
            lines.append("%s\trnd\texon\t%d\t%d\t.\t%s\t.\tgene_id \"%s\"; transcript_id \"%s.t%d\";%s"%(c,p,p+L,st,gid,gid,ti,(' gene_name "N%s"; gene_biotype "x";'%gid if random.random()<0.5 else "")))
            p+=L+random.choice([0,1,2,50,300,1500])
random.shuffle(lines)
open("fq.gtf","w").write("\n".join(lines)+"\n")
opts=["--sjdbGTFfile","fq.gtf","--sjdbOverhang",str(random.choice([99,50])),"--quantMode","TranscriptomeSAM","GeneCounts"]
if random.random()<0.4: opts+=["--quantTranscriptomeSAMoutput",random.choice(["BanSingleEnd","BanSingleEnd_ExtendSoftclip"])]
if random.random()<0.3: opts+=["--outFilterType","BySJout"]
if random.random()<0.3: opts+=["--outSAMunmapped","Within"]
if random.random()<0.2: opts+=["--twopassMode","Basic"]
ds=random.choice((["std_1.fq","std_2.fq"],["hard_1.fq","hard_2.fq"],["se_1.fq"]))
args=["--genomeDir","idx0","--readFilesIn"]+ds+opts
rc=[]
for tag,b,nt in (("fqr","/root/repo/oracle/_ref/STAR","1"),("fqo","/root/repo/oracle/_build/star_cli_oracle","3")):
    shutil.rmtree(tag,ignore_errors=True); os.makedirs(tag)
    rc.append(subprocess.call([b]+args+["--outFileNamePrefix",tag+"/","--runThreadN",nt],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL))
def same(f,bam=False,sam=False):
    a,b="fqr/"+f,"fqo/"+f
    if os.path.exists(a)!=os.path.exists(b): return False
    if not os.path.exists(a): return True
    x,y=open(a,"rb").read(),open(b,"rb").read()
    if bam: x,y=gzip.decompress(x),gzip.decompress(y)
    if sam: x=b"\n".join(l for l in x.split(b"\n") if not l.startswith(b"@")); y=b"\n".join(l for l in y.split(b"\n") if not l.startswith(b"@"))
    return x==y
bad=[f for f in ("ReadsPerGene.out.tab","SJ.out.tab","_STARgenome/exonInfo.tab","_STARgenome/transcriptInfo.tab","_STARgenome/geneInfo.tab","_STARgenome/exonGeTrInfo.tab","_STARgenome/sjdbList.out.tab","_STARgenome/sjdbInfo.txt") if not same(f)]
if not same("Aligned.toTranscriptome.out.bam",bam=True): bad.append("trBAM")
if not same("Aligned.out.sam",sam=True): bad.append("SAM")
print("seed",seed,"OK" if rc[0]==rc[1] and (rc[0]!=0 or not bad) else "MISMATCH",rc,len(lines),"exons"," ".join(opts[4:]),bad,flush=True)

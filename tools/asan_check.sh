#!/bin/bash
# Sanitizer pass over the host code, the oracle and the emulated kernels (no GPU needed).  usage: tools/asan_check.sh <unpacked tiny/ dir with idx0/ from twopass.tar.gz>
# Build first: make -f oracle/Makefile asan
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/oracle/_build/asan/star_cli_asan
T=${1:?directory with the unpacked golden inputs}
export ASAN_OPTIONS=detect_leaks=1
cd "$T" || exit 1
run() { tag=$1; shift; rm -rf AS_$tag; mkdir AS_$tag; timeout 900 "$B" "$@" --outFileNamePrefix AS_$tag/ --runThreadN 3 > /dev/null 2> AS_$tag.err; echo "$tag rc=$? findings=$(grep -c 'ERROR: AddressSanitizer\|runtime error\|LeakSanitizer' AS_$tag.err)"; }
run encode2pass --genomeDir idx --readFilesIn std_1.fq std_2.fq --outFilterType BySJout --outSAMattributes NH HI AS NM MD --outSAMtype BAM Unsorted SortedByCoordinate --quantMode TranscriptomeSAM GeneCounts --outSAMunmapped Within --twopassMode Basic --outReadsUnmapped Fastx
run clip --genomeDir idx --readFilesIn hard_1.fq hard_2.fq --clip3pNbases 0 200 --clip3pAdapterSeq GATC polyA --clip3pAdapterMMp 0.1 0.3 --clip5pNbases 3 0 --outSAMunmapped Within --outSAMattributes NH HI AS nM NM MD MC --quantMode TranscriptomeSAM
run gtfinsert --genomeDir idx0 --readFilesIn se_1.fq --sjdbGTFfile annot.gtf --sjdbOverhang 75 --sjdbInsertSave All --quantMode GeneCounts
export STAR_CLI_SJDB_EMUL=$ROOT/oracle/_build/asan/libengine_emul_asan.so   # the junction-insertion and suffix-sort kernels as emulated CTAs
run twopass_emulated_kernels --genomeDir idx --readFilesIn std_1.fq std_2.fq --twopassMode Basic --sjdbInsertSave All
for cap in 0 1500000; do   # 0: 32-bit path; > 0: the batched 64-bit path (the padding bin of this genome needs > 1.3 M slots)
    [ $cap != 0 ] && export STAR_B200_SA_LARGE_CAP=$cap
    rm -rf AS_gen$cap; mkdir AS_gen$cap
    timeout 900 "$B" --runMode genomeGenerate --genomeDir AS_gen$cap --genomeFastaFiles genome.fa --sjdbGTFfile annot.gtf --sjdbOverhang 99 --genomeSAindexNbases 7 --outFileNamePrefix AS_gen${cap}_ > /dev/null 2> AS_gen$cap.err
    echo "generate cap=$cap rc=$? findings=$(grep -c 'ERROR: AddressSanitizer\|runtime error' AS_gen$cap.err) SA=$(cmp AS_gen$cap/SA idx/SA > /dev/null && echo reference-bytes || echo DIFFERS)"
done
# ThreadSanitizer over the three-stage host pipeline (reader thread, engine thread, formatting threads; many small chunks)
unset STAR_CLI_SJDB_EMUL STAR_B200_SA_LARGE_CAP
rm -rf TS_run; mkdir TS_run
TSAN_OPTIONS=halt_on_error=0 timeout 900 "$ROOT/oracle/_build/asan/star_cli_tsan" --genomeDir idx --readFilesIn std_1.fq std_2.fq --gpuChunkReads 300 --runThreadN 4 --outFilterType BySJout \
    --quantMode TranscriptomeSAM GeneCounts --outSAMtype BAM Unsorted SortedByCoordinate --outSAMunmapped Within --outReadsUnmapped Fastx --outFileNamePrefix TS_run/ > /dev/null 2> TS_run.err
echo "tsan pipeline rc=$? races=$(grep -c 'WARNING: ThreadSanitizer' TS_run.err)"

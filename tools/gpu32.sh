mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "engine" 2>&1 | tail -12 > gpurun_out/t32.log; tail -2 gpurun_out/t32.log
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/an32_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/an32_$tag.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/stitch \1 total \2 ms_heavy \3/') | $(grep 'heavy kernel warp' gpurun_out/an32_$tag.log | tail -1 | cut -c1-90)"; }
run dflt A=1
run na1 STAR_B200_HEAVY_NA=1
run w5 STAR_B200_FLAT_DFS_CTAS_PER_SM=5
timeout 600 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --clock-control none -k regex:flat_dfs_warp_kernel -c 1 --csv --page raw --log-file gpurun_out/k2_quick.csv python tools/analyze_chunk.py 262144 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/k2_quick.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; d=dict(zip(hdr,rows[hi+2]))
print('K2 time', d.get('gpu__time_duration.sum'))
for k in hdr:
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        v=float(d[k] or 0)
        if v>0.2: print('  stall',k.split('issue_stalled_')[1].split('_per_')[0],round(v,2))
for k in ['smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum']:
    if k in d: print('  ',k,d[k])
PY

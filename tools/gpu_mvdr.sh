#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mvdr_times.py 32 501 129 > gpurun_out/mvdr_times_ref.json 2> gpurun_out/mvdr_times.err; cat gpurun_out/mvdr_times_ref.json; tail -2 gpurun_out/mvdr_times.err
timeout 300 python tools/mvdr_times.py 32 500 257 > gpurun_out/mvdr_times_paper.json 2>> gpurun_out/mvdr_times.err; cat gpurun_out/mvdr_times_paper.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_mvdr.csv python tools/mvdr_times.py 32 500 257 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_mvdr.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; h=rows[hi]
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ni=h.index('Metric Name'); ii=h.index('ID')
d={}
for r in rows[hi+1:]:
    if len(r)>vi: d.setdefault((int(r[ii]), r[ki].split('(')[0][-30:]),{})[r[ni]]=r[vi]
for k in sorted(d)[-12:]: print(k, d[k])
PY

O=gpurun_out
for v in 64 512; do
CMD="python bench.py --views $v --steps 4 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/ll_plain_$v.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/ll_$v.csv $CMD > $O/ll_$v.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("$O/ll_$v.csv")))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d=collections.defaultdict(dict)
for r in rows[hdr+1:]:
    if len(r)>iv: d[(r[iid],r[ik].split('(')[0])][r[im]]=float(r[iv].replace(',',''))
for (i,k),m in list(d.items())[:40]:
    if 'raster' in k or 'resolve' in k: print("$v", i, k, {a:round(b,1) for a,b in m.items()})
PY
done

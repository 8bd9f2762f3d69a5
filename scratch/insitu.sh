#!/bin/bash
# in-situ DRAM traffic per kernel of one training step (no cache flush between kernels, application replay)
out=$1; shift
ncu --cache-control none --clock-control none --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"k_march|k_scan_counts|k_emit|k_rgbnet|k_composite|k_ray_bwd|k_density|k_update|k_wgrad|k_prep" -s 39 -c 13 --csv --log-file $out python scratch/one_step.py 4 > /dev/null 2>&1
python - $out <<'PY'
import csv,collections,sys
rows=[l for l in open(sys.argv[1]) if not l.startswith('==')]
agg=collections.OrderedDict()
for x in csv.DictReader(rows):
    k=x['Kernel Name'].split('(')[0][-36:]; m=x['Metric Name']; v=float(x['Metric Value'].replace(',','')); u=x['Metric Unit']
    mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1,'us':1e3}.get(u,1)
    agg.setdefault(k,collections.defaultdict(list))[m].append(v*mult)
tot=0
for k,d in agg.items():
    n=len(d['gpu__time_duration.sum']); rd=sum(d['dram__bytes_read.sum'])/n; wr=sum(d['dram__bytes_write.sum'])/n; t=sum(d['gpu__time_duration.sum'])/n
    print("%-38s n=%d read %7.2f MB write %7.2f MB %6.1f us"%(k,n,rd/1e6,wr/1e6,t/1e3)); tot+=(rd+wr)*n
print("total MB", round(tot/1e6,1))
PY

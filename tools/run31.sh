export NBLK=16384 VARIANTS=13:0:4:0
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:lz4c4 -s 3 -c 3 --csv --log-file gpurun_out/k4.csv python tools/probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k4.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    print(r[ix["Kernel Name"]][:40], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]])
PY

#!/bin/bash
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.draw.instant,temperature.gpu,temperature.memory,clocks_event_reasons.active --format=csv,noheader -lms 20 > /tmp/smi.csv &
SMI=$!
sleep 1
timeout 200 python tools/r2_series.py 1e9 weighted 400 | cut -c1-2400
sleep 0.5
kill $SMI
awk -F, '{print $2","$3","$4","$5","$6","$7","$8}' /tmp/smi.csv | sort | uniq -c | sort -k1 -n -r | head -30
echo; sed -n '40,400p' /tmp/smi.csv | awk 'NR%6==0' | head -70

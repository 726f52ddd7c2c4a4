#!/bin/bash
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory --format=csv
timeout 200 python tools/r2_series.py 1e9 weighted 60
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active --format=csv
timeout 200 python tools/r2_series.py 1e9 weighted 30 50
timeout 200 python tools/r2_series.py 1e9 counts 40

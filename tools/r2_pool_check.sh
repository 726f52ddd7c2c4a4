#!/bin/bash
timeout 50 python -m pytest tests/test_round2_gpu.py -x -q -p no:cacheprovider -k "alias or plan or weighted_mean or packed_counts_take or async or out_" 2>&1 | tail -4

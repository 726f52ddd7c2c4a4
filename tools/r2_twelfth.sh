#!/bin/bash
XH_DEBUG_VERDICT=1 timeout 200 python tools/r2_ab.py 2.5e8 uniform_counts 6 2>&1 | tail -8
XH_DEBUG_VERDICT=1 timeout 200 python tools/r2_ab.py 2.5e8 uniform 6 2>&1 | tail -4
XH_DEBUG_VERDICT=1 timeout 200 python tools/r2_ab.py 2.5e8 counts 6 2>&1 | tail -4

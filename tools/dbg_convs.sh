#!/bin/bash
cd "$(dirname "$0")/.."
HRNET_B200_DBG=1 timeout 300 python tools/profile_convs.py 64 2 2>&1 | tee gpurun_out/dbg_convs.log

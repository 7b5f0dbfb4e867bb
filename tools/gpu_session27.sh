#!/bin/bash
cd "$(dirname "$0")/.."
echo "EPI=direct"; HRNET_B200_EPI=direct timeout 120 python tools/dbg_checksum.py 5 0 2>&1 | tail -3 | cut -c1-300
for f in 1 2 4; do echo "EPIFIX=$f"; HRNET_B200_EPIFIX=$f timeout 120 python tools/dbg_checksum.py 5 0 2>&1 | tail -3 | cut -c1-300; done

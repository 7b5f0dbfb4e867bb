#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/dbg_epi_race.py 0,1,2,4 2>&1 | tail -20

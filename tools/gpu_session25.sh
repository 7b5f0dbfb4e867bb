#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/dbg_checksum.py 5 0 2>&1 | tail -6
timeout 300 python tools/dbg_checksum.py 5 8 2>&1 | tail -6
timeout 300 python tools/dbg_checksum.py 16 8 2>&1 | tail -6

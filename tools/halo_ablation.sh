#!/bin/bash
# ablation timing of the halo kernel roles (results are garbage for dbg != 0; timing only)
for d in 0 1 2 4 8 3 7 15; do
  echo "dbg=$d"; CFR_HALO_DBG=$d timeout 300 python tools/profile_program.py --chunk 32 2>/dev/null | grep -E "halo" | awk -F'\t' '{printf "   %s  %s\n",$3,$5}'
done

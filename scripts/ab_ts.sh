export TS_MODES=0,0
for a in "16 16 1" "32 32 2" "64 64 3" "128 128 4"; do python scripts/debug_modes_ts.py $a 2>&1 | grep -v "^$" | head -3; done

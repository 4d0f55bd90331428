#!/bin/bash
# round 2, call G: split team kernel with one role per warp (A/B), parity tests, 1024-bit check
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2g_pytest.log 2>&1
tail -4 $O/r2g_pytest.log
timeout 900 python tools/split_ab.py > $O/r2g_split_ab.json 2> $O/r2g_split_ab.err
cat $O/r2g_split_ab.err | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.strip()); continue
    print(r['count'], 'team %.1f ms (%.2f w, f=%.3f) | split %.1f ms (%.2f w, f=%.3f) | auto %.1f ms (%d split launches) eq=%s' % (r['team_ms'], r['team_waves'], r['team_frac'], r['split_ms'], r['split_waves'], r['split_frac_of_team_work'], r['auto_ms'], r['auto_split_launches'], r['bytes_equal']))
"
rm -f tools/_ab/*.so
bash tools/gpu_ab1024.sh 4736 > $O/r2g_ab1024.txt 2>&1
cat $O/r2g_ab1024.txt

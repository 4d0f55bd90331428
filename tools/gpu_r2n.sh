#!/bin/bash
# round 2, call N: synccheck again (two-warp pairing with one barrier call site), parity tests, general A/B
O=gpurun_out
mkdir -p $O
( timeout 1500 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "(128 or 64) and (two_warps or lane_pair or split_team or wide_team)" ) > $O/r2n_synccheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" $O/r2n_synccheck.log | tail -3
( timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "(128 or 64) and (two_warps)" ) > $O/r2n_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" $O/r2n_racecheck.log | tail -3
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2n_pytest.log 2>&1
grep -E "passed|failed" $O/r2n_pytest.log
timeout 600 python tools/mapping_ab.py --which general --key-bits 512 --max-log2 15 > $O/r2n_pairduo_512.json 2> $O/r2n_ab.err
python - <<PY
import json
dd=json.load(open("$O/r2n_pairduo_512.json"))
for r in dd["sizes"]:
    print(r["count"], "1thr %.3f | duo %.3f ms f=%.3f eq=%s" % (r["one_thread"]["kernel_ms"], r["two_warps"]["kernel_ms"], r["two_warps"]["imad_frac"], r["bytes_equal"]))
PY

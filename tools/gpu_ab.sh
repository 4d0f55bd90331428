#!/bin/bash
# A/B bench of the default library and every tools/_ab/lib*.so (developer tool).  Usage: tools/gpu_ab.sh <tag> [bench args]
TAG=${1:-ab}; shift
O=gpurun_out; mkdir -p $O
cat > /tmp/ab_fmt.py <<'PY'
import sys, json, os
for ln in sys.stdin:
    d = json.loads(ln); r = d['roofline']
    print(os.environ.get('AB_NAME'), 'pairings/s=%.0f' % d['value'], 'frac=%.4f' % r['frac'], 'kernel_ms=%.2f' % r['kernel_ms'],
          'e2e=%.0f' % d['e2e']['value'], 'match', d['e2e']['bytes_match_device_path'], 'pairs', d['config']['pairs_per_gpu'])
PY
run() { local name=$1 lib=$2; shift 2
  BGN_B200_LIB=$lib timeout 600 python bench.py --no-cpu "$@" 2>>$O/${TAG}_err.txt | AB_NAME=$name python /tmp/ab_fmt.py | tee -a $O/${TAG}_ab.txt
}
run default "" "$@"
# variants/ and build/ are gpurun-ignored: copy the variant libraries to tools/_ab/
for f in tools/_ab/lib*.so; do [ -f "$f" ] && run $(basename $f .so) $PWD/$f "$@"; done

#!/bin/bash
# GPU call T: routing / tuning A/B of the tensor-core kernels (tools/ab_tc_route.py), the parity tests of every
# variant, and the whole GPU suite on the refactored host routing (pytest-xdist: the suite is bound by its CPU oracles)
set -u
mkdir -p gpurun_out
timeout 240 python tools/ab_tc_route.py > gpurun_out/ab_tc_route.jsonl 2> gpurun_out/ab_tc_route.err; echo "ab rc=$?"
tail -3 gpurun_out/ab_tc_route.err
python - <<'PY'
import json
for l in open('gpurun_out/ab_tc_route.jsonl'):
    r = json.loads(l)
    if r['part'] == 'A':
        print('A', r['op'], 'D', r['D'], 'N', r['N'], 'simt', r['simt_Tpairs_s'], 'tc', r['tc_Tpairs_s'], 'x', r['tc_speedup'], 'diff %.1e' % r['tc_vs_simt_relmax'])
    elif r['part'] == 'B':
        print('B', r['op'], 'D', r['D'], r['combo'], r['ms'], 'ms', r['Tpairs_s'], 'vsdef %.1e' % r['vs_default_relmax'], 'fp64 %.1e' % r.get('vs_fp64_relmax_64rows', -1))
    else:
        print(json.dumps(r))
PY
timeout 240 python -m pytest tests/test_gpu_tc_variants.py -q -m gpu -n 4 --tb=short --no-header -p no:cacheprovider > gpurun_out/pytest_tc_variants.log 2>&1; echo "pytest variants rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_tc_variants.log | tail -40
timeout 330 python -m pytest tests -q -m gpu -n 6 --ignore=tests/test_gpu_tc_variants.py --tb=short --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest full rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_full.log | tail -20

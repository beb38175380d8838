#!/bin/bash
# A/B of spmm_fused_kernel builds: fused tests on the default build, then config 4 and the mixed
# config-3 program for the default and every lib under lib/variants/.
set -u
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_filters.py tests/test_gpu_recipes.py -m gpu -x -q -k "fus" > gpurun_out/fused_pytest.log 2>&1; tail -2 gpurun_out/fused_pytest.log
run() {
  python benchmarks/run_configs.py --only 4 --out gpurun_out/cfg4_$1.json > /dev/null 2>&1
  python -c "import json;d=json.load(open('gpurun_out/cfg4_$1.json'));r=[x for x in d if 'fused_ms' in x][0] if isinstance(d,list) else d;print('$1 config4', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('plain_ms','fused_ms','unfused_spmm_plus_pointwise_ms','fused_equals_unfused_bitwise')})" 2>&1 | tail -1
  python benchmarks/ncu_kernels.py 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1 mixed', {k:round(v,3) for k,v in d['ms'].items() if k in ('spmm_f32_kernel','pointwise_kernel<float>','spmm_fused_kernel')})"
}
run default
for v in anemoi-transform_b200/anemoi_transform_b200/lib/variants/*.so; do
  AT_B200_LIBRARY=$PWD/$v run $(basename $v .so)
done

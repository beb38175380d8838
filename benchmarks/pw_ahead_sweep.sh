#!/bin/bash
# Ring depth of the transcendental pointwise path (AT_PW_AHEAD builds under lib/variants/):
# parity tests on the default build, then the per-kind table for every depth.
set -u
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_filters.py tests/test_gpu_filters_more.py tests/test_gpu_recipes.py -m gpu -x -q > gpurun_out/pw_pytest.log 2>&1; tail -3 gpurun_out/pw_pytest.log
echo "== default"; python benchmarks/epi_kinds.py --json gpurun_out/epi_kinds_default.json 2>&1 | grep -v "^$"
for v in anemoi-transform_b200/anemoi_transform_b200/lib/variants/*.so; do
  echo "== $v"; AT_B200_LIBRARY=$PWD/$v python benchmarks/epi_kinds.py --json gpurun_out/epi_kinds_$(basename $v .so).json 2>&1 | grep -v "^$"
done

#!/bin/bash
# Pointwise parity tests, then the per-kind table (default build).
set -u
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_filters.py tests/test_gpu_filters_more.py tests/test_gpu_recipes.py -m gpu -x -q > gpurun_out/pw_pytest.log 2>&1; tail -3 gpurun_out/pw_pytest.log
python benchmarks/epi_kinds.py --json gpurun_out/epi_kinds_default.json 2>&1 | grep -v "^$"

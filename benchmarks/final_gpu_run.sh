#!/bin/bash
# One GPU box, end of round: the full -m gpu suite, smoke, the bench line (no profiler), then
# the ncu launch list of the bench command and --set full captures of the kernels that changed.
# Everything lands in gpurun_out/ and is copied into profiles/ by hand afterwards.
set -u
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; tail -2 gpurun_out/r02_final_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err
python benchmarks/grib_unpack_bench.py > gpurun_out/r02_grib_unpack.json 2> gpurun_out/r02_grib_unpack.err
python benchmarks/grib_e2e.py > gpurun_out/r02_grib_e2e.json 2> gpurun_out/r02_grib_e2e.err
python benchmarks/knn_cell_sweep.py > gpurun_out/r02_knn_cell_sweep.json 2> /dev/null
# profiler passes (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --skip-pipeline > gpurun_out/r02_launch_run.log 2>&1
AT_UNDER_NCU=1 ncu --set full --clock-control none --import-source on -k regex:grib_unpack -f -o gpurun_out/r02_grib_unpack \
    python benchmarks/grib_unpack_bench.py > gpurun_out/r02_grib_ncu_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn1_thread -s 1 -c 1 -f -o gpurun_out/r02_knn1 \
    python benchmarks/ncu_knn.py > gpurun_out/r02_knn1_ncu_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_f32 -c 2 -f -o gpurun_out/r02_spmm \
    python bench.py --value-only --steps 2 --warmup 3 > gpurun_out/r02_ncu_run.log 2>&1
ls -la gpurun_out/*.ncu-rep

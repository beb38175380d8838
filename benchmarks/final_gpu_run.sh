#!/bin/bash
# One GPU box, end of round: the full -m gpu suite, smoke, the bench line (no profiler), the
# per-kernel tables, then the ncu launch list of the bench command and a --set full capture of
# the kernels that changed.  Everything lands in gpurun_out/ and is copied into profiles/ by
# hand afterwards.
set -u
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; tail -2 gpurun_out/r02_final_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err
python benchmarks/epi_kinds.py --json gpurun_out/r02_epi_kinds.json > gpurun_out/r02_epi_kinds.log 2>&1
python benchmarks/ncu_kernels.py > gpurun_out/r02_kernels_ms.json 2> gpurun_out/r02_kernels_ms.err
python benchmarks/run_configs.py --only 1,2,3d,3f64,4 --out gpurun_out/r02_configs.json > gpurun_out/r02_configs.log 2>&1
# profiler passes (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --skip-pipeline --skip-grib > gpurun_out/r02_launch_run.log 2>&1
AT_UNDER_NCU=1 ncu --set full --clock-control none --import-source on -k regex:pointwise_kernel -f -o gpurun_out/r02_pw_kinds \
    python benchmarks/epi_kinds.py > gpurun_out/r02_pw_ncu_run.log 2>&1
AT_UNDER_NCU=1 ncu --set full --clock-control none --import-source on -k regex:'spmm_fused|pointwise_kernel|spmm_f32' -f -o gpurun_out/r02_kernels \
    python benchmarks/ncu_kernels.py > gpurun_out/r02_kernels_ncu_run.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-page exports, drop the reports
for r in r02_pw_kinds r02_kernels; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2> /dev/null
  rm -f gpurun_out/$r.ncu-rep
done
du -sh gpurun_out

#!/bin/bash
# round 2, call D: MVDR (single-pass covariances, L2-resident mixture chunks): tests, stage time, per-kernel DRAM traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mvdr or pipeline or dropin or beamforming" > gpurun_out/r2d_pytest_mvdr.log 2>&1; tail -5 gpurun_out/r2d_pytest_mvdr.log
python tools/mvdr_times.py 32 500 257 20 > gpurun_out/r2d_mvdr_stage_paper.json 2>&1; cat gpurun_out/r2d_mvdr_stage_paper.json
python tools/mvdr_times.py 32 501 129 20 > gpurun_out/r2d_mvdr_stage_ref.json 2>&1; cat gpurun_out/r2d_mvdr_stage_ref.json
MISO_MVDR_L2_MB=100000 python tools/mvdr_times.py 32 500 257 20 > gpurun_out/r2d_mvdr_stage_paper_nochunk.json 2>&1; cat gpurun_out/r2d_mvdr_stage_paper_nochunk.json
MISO_MVDR_L2_MB=80 python tools/mvdr_times.py 32 500 257 20 > gpurun_out/r2d_mvdr_stage_paper_80mb.json 2>&1; cat gpurun_out/r2d_mvdr_stage_paper_80mb.json
MISO_MVDR_L2_MB=20 python tools/mvdr_times.py 32 500 257 20 > gpurun_out/r2d_mvdr_stage_paper_20mb.json 2>&1; cat gpurun_out/r2d_mvdr_stage_paper_20mb.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"scm_kernel|eig6_kernel|solve_kernel|apply_kernel" --csv --log-file gpurun_out/r2d_ncu_mvdr_raw.csv python tools/mvdr_times.py 32 500 257 1 > gpurun_out/r2d_ncu_mvdr.log 2>&1; tail -2 gpurun_out/r2d_ncu_mvdr.log

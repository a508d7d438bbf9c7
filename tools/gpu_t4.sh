#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stft or recording or pipeline or dropin" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log; tail -5 gpurun_out/pytest_f.log
timeout 300 python tools/stage_times.py 32 > gpurun_out/r1_c_small_stages.json 2> gpurun_out/stage.err; cat gpurun_out/r1_c_small_stages.json | python -c "import json,sys; d=json.load(sys.stdin); [print(r['stage'], round(r['ms_median'],4), round(r['achieved_GBs']), round(r['frac_of_hbm_peak'],3)) for r in d['stages']]"; tail -2 gpurun_out/stage.err

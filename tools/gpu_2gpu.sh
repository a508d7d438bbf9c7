#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/long_recording.py 2 > gpurun_out/r1_c_long_recording_1gpu.json 2> gpurun_out/long1.err; cat gpurun_out/r1_c_long_recording_1gpu.json; tail -2 gpurun_out/long1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/long_recording.py 2 > gpurun_out/r1_c_long_recording_2gpu.json 2> gpurun_out/long2.err; cat gpurun_out/r1_c_long_recording_2gpu.json; tail -2 gpurun_out/long2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r1_c_bench_miso1_paper_bf16x3_2gpu.json 2> gpurun_out/bench2.err; cut -c1-260 gpurun_out/r1_c_bench_miso1_paper_bf16x3_2gpu.json; tail -2 gpurun_out/bench2.err

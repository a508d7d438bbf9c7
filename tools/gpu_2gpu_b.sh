#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/long_recording.py 2 > gpurun_out/r1_c_long_recording_2gpu.json 2> gpurun_out/long2.err; cat gpurun_out/r1_c_long_recording_2gpu.json; grep -i "error" gpurun_out/long2.err | head -3

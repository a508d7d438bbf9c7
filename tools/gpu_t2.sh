#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mvdr" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log; tail -12 gpurun_out/pytest_f.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/utterance_mvdr_ranks.py 2> gpurun_out/umvdr.err | grep "^{" > gpurun_out/r1_c_utterance_mvdr_2gpu.json; cat gpurun_out/r1_c_utterance_mvdr_2gpu.json; grep -i "error" gpurun_out/umvdr.err | head -5

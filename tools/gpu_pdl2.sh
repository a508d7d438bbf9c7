#!/bin/bash
mkdir -p gpurun_out
for L in 0 2 0 2; do
MISO_PDL=$L timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_pdl$L.json 2> gpurun_out/bench_pdl$L.err; echo "PDL=$L $(cut -c80-200 gpurun_out/bench_pdl$L.json)"
done

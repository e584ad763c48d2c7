#!/bin/bash
# A very short call (1.6 GPU-minutes were left): id-space fold A/B at RMAT-26, then its parity test.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 55 python profiles/run_ids_ab.py 26 > gpurun_out/fold_ids_ab.jsonl 2> gpurun_out/fold_ids_ab.err; echo "ab rc=$?" > gpurun_out/summary_j.txt
cat gpurun_out/fold_ids_ab.jsonl >> gpurun_out/summary_j.txt
timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "id_space or fold_sampler_bit_equal" > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary_j.txt
tail -3 gpurun_out/pytest_gpu_j.log >> gpurun_out/summary_j.txt
cat gpurun_out/summary_j.txt

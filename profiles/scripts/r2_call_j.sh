#!/bin/bash
# Round 2, GPU call J (2 GPUs): the whole -m gpu suite where the two-GPU tests (torchrun NCCL check, one process driving two GPUs
# through the ABI / CLI) are not skipped, then smoke().
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2j_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2j_pytest.log 2>&1; stage pytest $?
tail -25 gpurun_out/r2j_pytest.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; stage smoke $?
tail -2 gpurun_out/r2j_smoke.log >> $S
cat $S

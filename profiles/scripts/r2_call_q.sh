#!/bin/bash
# Round 2, GPU call Q (2 GPUs): (1) the default N = 2 bench line (hub rows 0.5, one-sort build) with the e2e breakdown;
# (2) RMAT-27 -- 2^32 adjacency entries, beyond what one handle can hold -- built and walked SHARDED over two GPUs with two different
# shardings (hub rows 0 and 0.1): the full-round path checksums must agree.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2q_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
run() { # name, port, args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 $3 > gpurun_out/r2q_$1.json 2> gpurun_out/r2q_$1.err; stage $1 $?
  grep "bench " gpurun_out/r2q_$1.err | tail -4 >> $S
  tail -2 gpurun_out/r2q_$1.err >> $S
}
run bench_2gpu 29561 "--steps 10 --warmup 3"
run rmat27_hub00 29562 "--scale 27 --steps 2 --warmup 1 --batch-rounds 1 --hub-fraction 0 --no-e2e"
run rmat27_hub10 29563 "--scale 27 --steps 2 --warmup 1 --batch-rounds 1 --hub-fraction 0.1 --no-e2e"
python - >> $S 2>&1 <<'PY'
import json
for k in ("bench_2gpu", "rmat27_hub00", "rmat27_hub10"):
    try:
        d = json.load(open("gpurun_out/r2q_%s.json" % k))
        c = d["config"]
        print(k, "value %.3e" % d["value"], "vertices", c["vertices_present"], "shard GB", c["shard_bytes_hbm_rank0"] / 1e9, "build_s", c["build_s"],
              "checksum", d.get("checksum_sharded_last_round"), "parity", (d.get("parity_at_scale") or {}).get("equal"), "replicas", (d.get("replicas") or {}).get("value"),
              "e2e", d.get("e2e") and {x: d["e2e"][x] for x in ("value", "seconds", "h2d_allgather_build_s", "exchange_block_setup_s", "walk_and_d2h_s")})
    except Exception as ex:
        print(k, "no line:", ex)
PY
cat $S

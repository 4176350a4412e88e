#!/bin/bash
# One box, N GPUs: bench.py at N (device-timed value, e2e through btb_dist_matrix_host_comm, snps.csv wall),
# BASELINE config 3 (20,000 x 4.35 Mb) through snp_sites() -> build_snp_matrix() with the checks of
# tools/bench_northstar.py, and -- when the box has the memory -- the north-star target itself (50,000 x 4.35 Mb).
#   gpurun --gpus 8 -- bash tools/run_multi_gpu.sh 8
N=${1:-8}
mkdir -p gpurun_out
{ echo "== cores"; nproc; echo "== memory"; free -g; echo "== filesystems"; df -h /dev/shm .; echo "== topology"; nvidia-smi topo -m; } > gpurun_out/box_${N}gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench rc=$?"
timeout 1200 python tools/bench_northstar.py --samples ${C3_SAMPLES:-20000} --gpus $N --files 100 \
    --out gpurun_out/northstar_c3_${N}gpu.json > gpurun_out/northstar_c3_${N}gpu.log 2>&1
echo "c3 rc=$?"
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
echo "MemAvailable ${avail} GB"
if [ "${avail}" -gt 330 ] && [ -z "$SKIP_50K" ]; then
    timeout 1500 python tools/bench_northstar.py --samples 50000 --gpus $N --files 250 --no-one-gpu --repeat 1 \
        --out gpurun_out/northstar_50k_${N}gpu.json > gpurun_out/northstar_50k_${N}gpu.log 2>&1
    echo "50k rc=$?"
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/northstar_*_${N}gpu.json")):
    d = json.load(open(f))
    print(f, {k: d.get(k) for k in ("fasta_bytes", "generation_s", "expected_number_of_snps", "snps_csv_wall_s", "ingest_GBps", "all_checks_ok")})
    for r in d.get("runs", []) + ([d["one_gpu"]] if "one_gpu" in d else []):
        print("   ", r["label"], "gpus", r["gpus"], "snp_sites %.2f s" % r["snp_sites_s"], "build_snp_matrix %.2f s" % r["build_snp_matrix_s"])
PY
tail -c 1500 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_${N}gpu.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "csv_wall_s", "csv_wall_s_disk")}, d["e2e"]["seconds_per_step"], d["e2e"].get("rank0_phase_ms"), d["cpu_baseline"]["rows_match_gpu"])
except Exception as e:
    print("bench line unreadable:", e)
PY

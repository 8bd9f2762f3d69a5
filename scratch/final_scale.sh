#!/bin/bash
# Scaling run on one box, back to back: N = 1, 2, 4, 8 (training + render; the S512 sub-benchmark and the CPU legs are left to the 1-GPU run)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02c}
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus=$NG"
timeout 300 python bench.py --no-s512 --no-cpu-render > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  if [ "$NG" -ge $n ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-s512 --no-cpu-render \
      > gpurun_out/bench_${TAG}_n$n.json 2> gpurun_out/bench_${TAG}_n$n.err; echo "n$n rc=$?"
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get("render",{})
        print(f, "value=%.4g ms=%.4f e2e=%.4g sync=%.4g render=%.1f e2e_fps=%.1f parity=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"].get("synchronous_value",0),r.get("value",0),r.get("e2e_fps",0),d.get("parity",{}).get("ok")))
    except Exception as e:
        print(f, "unreadable", e)
PY

#!/bin/bash
# One-call GPU check: full -m gpu suite, smoke, bench at N=1 and (when 2 GPUs are visible) N=2 with both frame-assembly modes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01e}
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus=$NG"
timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
if [ "$NG" -ge 2 ]; then
  for mode in peer nccl; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --render-gather $mode \
      > gpurun_out/bench_${TAG}_n2_${mode}.json 2> gpurun_out/bench_${TAG}_n2_${mode}.err; echo "bench n2 $mode rc=$?"
  done
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("render",{})
        print(f, "value=%.3g ms=%.4f e2e=%.3g render_fps=%.1f e2e_fps=%.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r.get("value",0),r.get("e2e_fps",0)), r.get("sharding"))
    except Exception as e:
        print(f, "unreadable", e)
PY

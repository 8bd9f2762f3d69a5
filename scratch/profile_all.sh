#!/bin/bash
# One call on the GPU box: launch list of the bench, --set full of one training step and of one frame, in-situ DRAM traffic.
# usage: bash scratch/profile_all.sh r02b
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02b}
K="k_march|k_scan_counts|k_emit|k_rgbnet|k_composite|k_ray_bwd|k_density|k_update|k_wgrad|k_prep|k_ray_pe"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --frames 3 --no-s512 --no-cpu-render > gpurun_out/launches_${TAG}.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:"$K" -s 13 -c 13 -f -o gpurun_out/prof_step_${TAG} python scratch/one_step.py 2 > gpurun_out/prof_step_${TAG}.log 2>&1; echo "step full rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:"k_render_gather|k_render_mlp_tc|k_render_composite|k_render_probe|k_render_march_lanes|k_render_pass1|k_render_pass2|k_render_emit" -s 8 -c 8 -f -o gpurun_out/prof_render_${TAG} python scratch/render_one.py 1 2 > gpurun_out/prof_render_${TAG}.log 2>&1; echo "render full rc=$?"
timeout 500 bash scratch/insitu.sh gpurun_out/insitu_${TAG}.csv > gpurun_out/insitu_${TAG}.txt 2>&1; echo "insitu rc=$?"; cat gpurun_out/insitu_${TAG}.txt
ls -la gpurun_out/*${TAG}*

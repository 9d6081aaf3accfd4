#!/bin/bash
# 8-GPU box: C5 strong-scaling sweep (3840x2160, the frame's subframes split over N GPUs, NCCL reduce every frame), the full C5 frame
# (8192 spp) through bench.py and through the C++ driver, and a C3 dataset run with the dynamic scene queue.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale8.sh <tag> [c3_scenes]'
tag=${1:-r02}; scenes=${2:-48}
mkdir -p gpurun_out
run() { # n, port, args...
  local n=$1 port=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; fi
}
for n in 1 2 4 8; do
  run $n $((29600 + n)) --config C5 --scaling strong --spp 512 --steps 3 --warmup 3 --no-secondary --no-cpu-baseline 2> gpurun_out/${tag}_c5_strong_n$n.err | grep '^{' > gpurun_out/${tag}_c5_strong_n$n.json
  python -c "import json,sys; r=json.load(open('gpurun_out/${tag}_c5_strong_n$n.json')); print('C5 strong 512 spp N=$n', round(r['value'],1), 'Mpaths/s  e2e', round(r['e2e']['value'],1), ' ms/frame', round(r['ms_per_step'],1), ' kernel share', round(r['roofline']['kernel_share_of_step'],4))"
done
run 8 29650 --config C5 --scaling strong --spp 8192 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline 2> gpurun_out/${tag}_c5_full_n8.err | grep '^{' > gpurun_out/${tag}_c5_full_n8.json
python -c "import json; r=json.load(open('gpurun_out/${tag}_c5_full_n8.json')); print('C5 FULL 8192 spp N=8', round(r['value'],1), 'Mpaths/s  e2e', round(r['e2e']['value'],1), ' s/frame', round(r['ms_per_step']/1e3,2))"
./deepestscatter_b200/datagen render synth:512 --gpus 8 --spp 8192 --width 3840 --height 2160 --mode all --light Front --out gpurun_out > gpurun_out/${tag}_datagen_c5_n8.log 2>&1; tail -2 gpurun_out/${tag}_datagen_c5_n8.log
rm -f gpurun_out/multigpu.Front.PathTracing.exr
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 tools/bench_dataset.py --scenes $scenes 2> gpurun_out/${tag}_c3_n8.err | grep '^{' > gpurun_out/${tag}_c3_n8.json
python -c "import json; r=json.load(open('gpurun_out/${tag}_c3_n8.json')); print('C3 N=8', r['scenes'], 'scenes', r['samples'], 'samples in', round(r['seconds'],1), 's =', round(r['value'],1), 'samples/s; stages', {k: round(v,1) for k,v in r['seconds_per_stage_per_rank'].items()})"

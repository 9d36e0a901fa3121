# usage: bash tools/evidence_multigpu.sh N [with_image]   (under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
$TR bench.py --gpus $N --workload infer --scene-scale 1.0 --steps 2 --warmup 1 > gpurun_out/r02_infer_scale1_${N}gpu.json 2> gpurun_out/r02_infer_scale1_${N}gpu.err
if [ -n "$2" ]; then
$TR bench.py --gpus $N --workload train_image --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_train_image_${N}gpu.json 2> gpurun_out/r02_train_image_${N}gpu.err
fi
true

set -x
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --tiles 4 --micro-batch 4 --no-cpu-baseline --no-graph --no-extras --no-cudnn-benchmark > /dev/null 2>&1
python tools/bench_ops.py --full > gpurun_out/r02_ops_sweep_clustered.jsonl 2> gpurun_out/r02_ops_sweep.err
python tools/bench_ops.py --full --uniform > gpurun_out/r02_ops_sweep_uniform.jsonl 2>> gpurun_out/r02_ops_sweep.err
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/prof_ops python tools/profile_ops.py --all > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_ops.ncu-rep > gpurun_out/r02_ncu_ops_summary.txt
ncu --set full --clock-control none -k regex:"linear_x3_persistent|wgrad_x3" -c 4 -f -o /tmp/prof_gemm python tools/bench_linear.py 524288 1024 512 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > gpurun_out/r02_ncu_gemm_summary.txt
python bench.py --workload infer --scene-scale 1.0 --steps 2 --warmup 1 > gpurun_out/r02_infer_scale1.json 2> gpurun_out/r02_infer_scale1.err
python bench.py --workload train_image --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_train_image.json 2> gpurun_out/r02_train_image.err
ls -la gpurun_out

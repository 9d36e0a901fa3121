set -x
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --tiles 4 --micro-batch 4 --no-cpu-baseline --no-graph --no-extras --no-cudnn-benchmark > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"linear_x3_persistent" -c 2 -f -o /tmp/prof_gemm python tools/bench_linear.py 524288 1024 512 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > gpurun_out/r02_ncu_gemm_summary.txt
ncu --set full --clock-control none -k regex:"wgrad_f16_pair" -c 2 -f -o /tmp/prof_wg python tools/bench_linear.py 524288 1024 512 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_wg.ncu-rep >> gpurun_out/r02_ncu_gemm_summary.txt
python tools/shape_table.py > gpurun_out/r02_shape_table.txt 2>&1
python tools/gemm_probe.py > gpurun_out/r02_gemm_probe.txt 2>&1
ls -la gpurun_out | tail -8

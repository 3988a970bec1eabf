set -x
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches_v2.csv python tools/profile_step.py 256 > gpurun_out/profile_step.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --config B --no-eager-baseline > gpurun_out/bench_n1_cfgB.json 2> gpurun_out/bench_cfgB.err
timeout 300 python bench.py --steps 20 --warmup 3 --pairs 55 --no-eager-baseline > gpurun_out/bench_n1_p55.json 2> gpurun_out/bench_p55.err
timeout 100 python tools/bench_attention.py > gpurun_out/bench_attention.txt 2>&1
timeout 100 python tools/bench_elementwise.py > gpurun_out/bench_elementwise.txt 2>&1
tail -c 600 gpurun_out/bench_n1.json

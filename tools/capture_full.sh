# ncu --set full of representative launches of the final build (one eager step between cudaProfilerStart/Stop)
set -x
mkdir -p gpurun_out
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 44 --launch-count 8 -o gpurun_out/r02_gemm_fwd_full -f python tools/profile_step.py 256 > gpurun_out/cap1.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 150 --launch-count 10 -o gpurun_out/r02_gemm_bwd_full -f python tools/profile_step.py 256 > gpurun_out/cap2.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gelu_bwd_kernel|gelu_fwd_kernel|ln_bwd_kernel|dropout_grad_vec|bn_bwd_apply_gsum" --launch-skip 20 --launch-count 8 -o gpurun_out/r02_stream_full -f python tools/profile_step.py 256 > gpurun_out/cap3.log 2>&1
for f in r02_gemm_fwd_full r02_gemm_bwd_full r02_stream_full; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv

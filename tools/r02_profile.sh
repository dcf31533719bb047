# Round-2 profile pass (one gpurun call): launch lists of one cfg2 step and of one batched call through both pipelines,
# then `ncu --set full` of the dominant kernel (the fused variance sweep of gemm_nt_kernel).
set -x
python tools/profile_step.py > gpurun_out/r02_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg2_step.csv python tools/profile_step.py > gpurun_out/r02_ncu_step.log 2>&1
python tools/profile_batched.py 1 > gpurun_out/r02_plain_batched.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_batched.csv python tools/profile_batched.py 1 > gpurun_out/r02_ncu_batched.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 25 -c 25 -o gpurun_out/r02_gemm_step python tools/profile_step.py > gpurun_out/r02_ncu_gemm.log 2>&1
tail -2 gpurun_out/r02_plain_step.log gpurun_out/r02_plain_batched.log gpurun_out/r02_ncu_gemm.log

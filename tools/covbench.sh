set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph or cov or ragged or separable or short or two_streams" 2>&1 | tail -5 > gpurun_out/r2c_tests.log
python tools/bench_cov.py J16 > gpurun_out/r2c_cov_J16.json 2> gpurun_out/r2c_cov.err
for v in 3 6; do
  GPM_EXTRA_NVCC_FLAGS="-DGPM_EXP_LOG2J=$v" python -m gaussianprocesspathmodelling_b200.build --force > /dev/null 2>> gpurun_out/r2c_cov.err
  python tools/bench_cov.py J$v > gpurun_out/r2c_cov_J$v.json 2>> gpurun_out/r2c_cov.err
done
GPM_EXTRA_NVCC_FLAGS="-DGPM_EXP_LIBDEVICE" python -m gaussianprocesspathmodelling_b200.build --force > /dev/null 2>> gpurun_out/r2c_cov.err
python tools/bench_cov.py libdevice > gpurun_out/r2c_cov_libdevice.json 2>> gpurun_out/r2c_cov.err
cat gpurun_out/r2c_tests.log gpurun_out/r2c_cov_*.json; tail -5 gpurun_out/r2c_cov.err

GPU_TEST_FILES="test_gpu_optim test_gpu_api_misc" bash tools/gpu_trip_r2.sh tests bench noncu

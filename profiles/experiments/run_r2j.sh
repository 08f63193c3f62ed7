set -x
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py tests/test_train_gpu.py tests/test_bn_algebra_gpu.py -x -q -m gpu > gpurun_out/t_r2j.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2j.log
tail -6 gpurun_out/t_r2j.log
for m in "" "ARGUS_BN_REDUCE_FUSED=0" "ARGUS_B_RESIDENT=1"; do
  tag=$(echo $m | tr '= ' '__')
  env $m ARGUS_PROFILE_DETAIL=1 timeout 300 python profiles/profile_detail.py > gpurun_out/detail_r2j_$tag.log 2>&1
  echo "== $m"; head -1 gpurun_out/detail_r2j_$tag.log
done

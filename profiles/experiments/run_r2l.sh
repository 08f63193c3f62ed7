set -x
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py tests/test_train_gpu.py tests/test_bn_algebra_gpu.py tests/test_optin_modes_gpu.py -x -q -m gpu > gpurun_out/t_r2l.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2l.log
tail -6 gpurun_out/t_r2l.log
for m in "" "ARGUS_FUSED_TAIL=1" "ARGUS_BN_REDUCE_FUSED=0"; do
  tag=$(echo $m | tr '= ' '__')
  env $m ARGUS_PROFILE_DETAIL=1 timeout 300 python profiles/profile_detail.py > gpurun_out/detail_r2l_$tag.log 2>&1
  echo "== $m"; head -1 gpurun_out/detail_r2l_$tag.log
done

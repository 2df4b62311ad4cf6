set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "flash_attention" > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
tail -15 gpurun_out/t_attn.log
timeout 200 python scripts/bench_kernels.py attn > gpurun_out/kb_attn2.log 2>&1; cat gpurun_out/kb_attn2.log
SMTL_FATTN_V1=1 timeout 200 python scripts/bench_kernels.py attn > gpurun_out/kb_attn1.log 2>&1; cat gpurun_out/kb_attn1.log
